"""Triangle::CalculateLocalSurface in full (RayTracer/Shapes.h:982-1083) -- hitp, uv, du, dv, n, wo of the
LocalSurfaceInfo record (Shapes.h:144-170).

The expected values are outputs of the REFERENCE'S OWN compiled code (tests/golden/ref_pin.npz, group `local_surface`, produced by
tools/make_ref_golden.py through oracle/ref_harness.cpp); tests/test_cpu_ref_pin.py shows the restated oracle reproduces them bit for bit.
Here the product is held to the same arrays: its host restatement of the function (no GPU) and the device kernels behind
crt_traverse_local_surface / crt_kat_local_surface."""
import os

import numpy as np
import pytest

import ref_pin_cases as P
from common import bits
from computational_ray_tracer_b200 import api, scenes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_pin.npz")
FIELDS = ("hitp", "uv", "du", "dv", "n", "wo")


@pytest.fixture(scope="module")
def golden():
    return {k.split("/", 1)[1]: v for k, v in np.load(GOLD).items() if k.startswith("local_surface/")}


def _object_to_render(tri):
    """ObjectToRender of an identity rigid transform is the y <-> z permutation (Shapes.h:175-182), which CalculateLocalSurface applies to
    every vertex (:995-997): exact, so it is applied here and crt_kat_local_surface receives the p*_w of the reference."""
    return np.ascontiguousarray(tri[..., [0, 2, 1]])


def _cases():
    plain = [dict(scenes.random_soup(600, 8)[0], normals=None)]
    return {"plain": (plain, {}), "normals_only_hf": (scenes.heightfield(32), {}), "attributes": (P.attribute_model(), {}),
            "attributes_rigid": (P.attribute_model(500, 6), dict(rigid=P._rigid(12, -7, 25, 0.35), precomputed_world=False))}


def test_direct_call_on_the_host_matches_the_compiled_reference(crt_lib, golden):
    """crt_kat_local_surface(on_device=0): the product's restatement compiled for the host, on the degenerate-triangle cases
    (collinear vertices reach the Frisvad-frame fallback of Shapes.h:1016-1029; collapsed ones give the reference's NaNs)."""
    tri, bary, rayd = P.degenerate_triangles()
    got = P._canon(api.kat_local_surface(_object_to_render(tri).reshape(-1, 9), bary, rayd, on_device=False))
    want = golden["direct.info"]
    assert np.array_equal(bits(got), bits(want))
    du = want[40:48, 5:8]
    unit = np.isfinite(du).all(1) & (np.abs(np.linalg.norm(du, axis=1) - 1) < 1e-5)
    assert unit.any() and np.isnan(want[48:52, 5:11]).all()       # the fallback frame really ran (collinear) / really failed (collapsed)


@pytest.mark.gpu
def test_direct_call_on_the_device_matches_the_compiled_reference(gpu_ctx, golden):
    tri, bary, rayd = P.degenerate_triangles()
    got = P._canon(api.kat_local_surface(_object_to_render(tri).reshape(-1, 9), bary, rayd, on_device=True))
    assert np.array_equal(bits(got), bits(golden["direct.info"]))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, api.DEFAULT_TRACE_MODE])
@pytest.mark.parametrize("name", ["plain", "normals_only_hf", "attributes", "attributes_rigid"])
def test_traverse_returns_the_reference_record(gpu_ctx, golden, name, mode):
    meshes, kw = _cases()[name]
    ms = api.MeshSet(meshes)
    oc = api.Octtree_Model(ms, rigid=kw.get("rigid"), precomputed_world=kw.get("precomputed_world", True))
    sc = api.Scene(gpu_ctx); sc.set_model(oc); sc.commit()
    b = oc.model_bounds()
    rays = P._rays(1200, 41, center=tuple((b[:3] + b[3:]) / 2), spread=220.0)
    g = sc.traverse_local_surface(rays, mode=mode)
    assert np.array_equal(g["found"], golden[f"{name}.found"])
    f = g["found"] > 0
    assert 0.3 < f.mean()
    for key in FIELDS:
        assert np.array_equal(bits(np.ascontiguousarray(g[key][f])), bits(golden[f"{name}.{key}"])), (name, key)
    sc.close(); oc.close()


@pytest.mark.gpu
def test_traverse_record_against_the_oracle_at_scale(gpu_ctx, oracle):
    """A larger live sweep against the oracle (which is pinned above): 40 k rays on a 51 k-triangle height field with attributes."""
    import oracle_lib as O
    m = scenes.heightfield(160, with_light=False)[0]
    rs = np.random.RandomState(3)
    nv = len(m["positions"])
    meshes = [dict(m, texcoords=rs.rand(nv, 2).astype(np.float32), tangents=rs.normal(size=(nv, 3)).astype(np.float32),
                   bitangents=rs.normal(size=(nv, 3)).astype(np.float32))]
    orc = O.OracleScene(); orc.set_model(meshes); orc.build_octree()
    oc = api.Octtree_Model(api.MeshSet(meshes))
    sc = api.Scene(gpu_ctx); sc.set_model(oc); sc.commit()
    rays = P._rays(40000, 5, center=(0, 0, 800), spread=380.0)
    g = sc.traverse_local_surface(rays); o = orc.traverse_local_surface(rays)
    assert np.array_equal(g["found"], o["found"])
    f = o["found"] > 0
    assert f.mean() > 0.5
    for key in FIELDS:
        assert np.array_equal(bits(np.ascontiguousarray(g[key][f])), bits(np.ascontiguousarray(o[key][f]))), key
    sc.close(); oc.close(); orc.close()
