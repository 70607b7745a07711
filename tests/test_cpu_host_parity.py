"""Host-side halves of the product (octree builder, flatten inputs, matrices, tables) against the oracle -- no GPU.

The octree is built on the host by libcrt_b200's own C++ builder (Octtree_Model.h:33-63,180-358 restated a second
time, independently of the oracle); node order, bounds, child ids and per-leaf (mesh, tri) order must agree exactly
because closest-hit ids depend on them."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from computational_ray_tracer_b200 import api, scenes
from computational_ray_tracer_b200._capi import f32p


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


SCENES = {
    "heightfield": lambda: scenes.heightfield(48),
    "heightfield_fat_leaves": lambda: scenes.heightfield(200),      # leaves the reference cannot split (up to 321 triangles): memoised aborts
    "soup": lambda: scenes.random_soup(1500, seed=3),
    "cornell": scenes.cornell_box,
    "axis_grid": lambda: scenes.axis_grid(20, layers=2),
    "negative_octant": lambda: [dict(m, positions=m["positions"] - np.float32([900, 900, 1500])) for m in scenes.random_soup(300, seed=5)],
}


@pytest.mark.parametrize("name", sorted(SCENES))
def test_octree_build_matches_oracle(oracle, crt_lib, name):
    meshes = SCENES[name]()
    orc = O.OracleScene(); orc.set_model(meshes); orc.build_octree()
    ms = api.MeshSet(meshes); oc = api.Octtree_Model(ms)
    so, sc = orc.octree_stats(), oc.stats()
    assert so == pytest.approx(sc)
    do, dc = orc.octree_dump(), oc.dump()
    assert np.array_equal(_bits(do["bounds"]), _bits(dc["bounds"]))
    assert np.array_equal(do["leaf"], dc["leaf"])
    assert np.array_equal(do["child"], dc["child"])
    assert np.array_equal(do["list_off"], dc["list_off"])
    assert np.array_equal(do["pairs"], dc["pairs"])
    assert np.array_equal(_bits(orc.model_bounds()), _bits(oc.model_bounds()))
    if name == "negative_octant":          # Shapes.h:1292: max is initialised with FLT_MIN, not -FLT_MAX
        assert (oc.model_bounds()[3:] > 0).all()
    oc.close(); orc.close()


def test_octree_with_rigid_transform_and_backface(oracle, crt_lib):
    meshes = scenes.random_soup(800, seed=9, center=(0, 0, 0))
    rigid = np.eye(4, dtype=np.float32)
    rigid[3, :3] = (10, -20, 600)                       # column-major: translation lives in the 4th column
    ang = np.float32(0.3)
    rigid[0, 0], rigid[0, 1], rigid[1, 0], rigid[1, 1] = np.cos(ang), np.sin(ang), -np.sin(ang), np.cos(ang)
    orc = O.OracleScene(); orc.set_model(meshes, rigid=rigid, precomputed_world=False, cull_backface=True, look_dir=(0, 0, 1)); orc.build_octree()
    ms = api.MeshSet(meshes); oc = api.Octtree_Model(ms, rigid=rigid, precomputed_world=False)
    do, dc = orc.octree_dump(), oc.dump()
    assert np.array_equal(_bits(do["bounds"]), _bits(dc["bounds"]))
    assert np.array_equal(do["pairs"], dc["pairs"])
    bits = oc.compute_backface((0, 0, 1))
    for m in range(len(meshes)):
        ob = orc.backfacing(m, len(meshes[m]["indices"]))
        assert np.array_equal(ob, bits[m])
        assert 0 < ob.mean() < 1
    oc.close(); orc.close()


def test_empty_and_tiny_models(oracle, crt_lib):
    one = [dict(positions=np.float32([[0, 0, 500], [10, 0, 500], [0, 10, 500]]), normals=None, indices=np.uint32([[0, 1, 2]]))]
    orc = O.OracleScene(); orc.set_model(one); orc.build_octree()
    oc = api.Octtree_Model(api.MeshSet(one))
    assert orc.octree_stats()["nodes"] == oc.stats()["nodes"] == 1
    assert np.array_equal(orc.octree_dump()["pairs"], oc.dump()["pairs"])
    oc.close(); orc.close()


@pytest.mark.parametrize("kind", [0, 1])
def test_camera_matrices(oracle, crt_lib, kind):
    for (pos, look, fov, res) in [((0, 0, 0), (0, 0, 1), 45.0, (1920, 1080)), ((5, -3, 2), (0.3, 0.1, 1), 60.0, (500, 500)), ((0, 10, 0), (1, 0, 0.2), 30.0, (256, 256))]:
        a = api.camera_matrices(kind, 1.0, 1000.0, fov, pos, look, (0, 1, 0), res[0], res[1], sensor_w=2.0, sensor_h=1.5)
        b = O.camera_matrices(kind, 1.0, 1000.0, 2.0, 1.5, fov, pos, look, (1, 0, 0), (0, 1, 0), res[0], res[1])
        assert np.array_equal(_bits(a[0]), _bits(b[0]))
        assert np.array_equal(_bits(a[1]), _bits(b[1]))


def test_shape_matrices(oracle, crt_lib):
    rs = np.random.RandomState(2)
    for _ in range(5):
        rigid = np.eye(4, dtype=np.float32)
        q, _r = np.linalg.qr(rs.randn(3, 3))
        rigid[:3, :3] = q.astype(np.float32)
        rigid[3, :3] = rs.uniform(-100, 100, 3)
        a = api.shape_matrices(rigid)
        o2r = np.zeros(16, np.float32); r2o = np.zeros(16, np.float32)
        oracle.lib().orc_shape_matrices(O.fp(O.f32(rigid).reshape(-1)), O.fp(o2r), O.fp(r2o))
        assert np.array_equal(_bits(a[0]), _bits(o2r))
        assert np.array_equal(_bits(a[1]), _bits(r2o))


def test_spectral_tables_and_colour_constants(oracle, crt_lib):
    for which in range(4):
        a = np.zeros(471, np.float32); b = np.zeros(471, np.float32)
        oracle.lib().orc_dense_table(which, O.fp(a))
        assert crt_lib.crt_dense_table(which, b.ctypes.data_as(f32p)) == 0
        assert np.array_equal(_bits(a), _bits(b))
        assert a.max() > 0
    oa = [np.zeros(9, np.float32), np.zeros(9, np.float32), np.zeros(9, np.float32), np.zeros(2, np.float32)]
    ca = [np.zeros(9, np.float32), np.zeros(9, np.float32), np.zeros(9, np.float32), np.zeros(2, np.float32)]
    oracle.lib().orc_color_constants(*[O.fp(x) for x in oa])
    assert crt_lib.crt_color_constants(*[x.ctypes.data_as(f32p) for x in ca]) == 0
    for x, y in zip(oa, ca):
        assert np.array_equal(_bits(x), _bits(y))
    # CIE_Y_integral (spectrum.h:21) is the integral of the Y matching curve over 1 nm bins
    y = np.zeros(471, np.float32); oracle.lib().orc_dense_table(1, O.fp(y))
    assert abs(float(y.astype(np.float64).sum()) - 106.856895) < 1e-3
    # sRGB primaries: XYZFromRGB * RGBFromXYZ = I
    m = oa[2].reshape(3, 3).T.astype(np.float64) @ oa[1].reshape(3, 3).T.astype(np.float64)
    assert np.allclose(m, np.eye(3), atol=1e-5)


def test_sigmoid_and_wavelength_sampling(oracle):
    L = oracle.lib()
    assert L.orc_sigmoid_eval(0, 0, 0, 500.0) == 0.5                     # color.h:394-399: s(0) = 0.5
    c = np.zeros(3, np.float32)
    assert L.orc_grey_sigmoid(0.5, O.fp(c)) == 0 and c[0] == 0 and c[1] == 0 and c[2] == 0
    lam = np.zeros(8, np.float32); pdf = np.zeros(8, np.float32)
    for u in [0.0, 0.25, 0.5, 0.999]:
        L.orc_sample_visible(u, O.fp(lam), O.fp(pdf))
        assert (lam >= 360).all() and (lam <= 830).all() and (pdf > 0).all()
        # pdf is the analytic density of the sampling map (Sampling.h:63-71)
        want = 0.0039398042 / np.cosh(0.0072 * (lam.astype(np.float64) - 538)) ** 2
        assert np.allclose(pdf, want, rtol=1e-5)
