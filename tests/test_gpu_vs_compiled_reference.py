"""The CUDA path against the REFERENCE'S OWN compiled code (oracle/_ref/libcrt_ref.so, see tests/ref_lib.py) -- no
restatement in between.  The .so is built in the build container from /root/reference and travels with the repo snapshot;
where it is absent these tests skip (tests/test_cpu_ref_pin.py then still ties the oracle to the committed reference
outputs, and every other GPU test ties the CUDA path to the oracle).

* hit ids: the reference's own closest hit that exposes ids is TriModel::BasicIntersect (brute force, Shapes.h:1414-1471);
  Octtree_Model::Traverse returns only the surface record, so the GPU's octree hit id is checked by rebuilding that record
  with the reference's Triangle(id).BasicIntersect -> CalculateLocalSurface and comparing it with what Traverse returned.
* film: the reference's evaluate_pixel + Li (RayTracerTestApp.h:218-345) with the same counter-based sampler streams.
"""
import numpy as np
import pytest

import common
import ref_lib as R
import ref_pin_cases as P
from common import bits
from computational_ray_tracer_b200 import api, scenes

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="compiled reference (oracle/_ref/libcrt_ref.so) did not travel")]


def _gpu_scene(ctx, meshes, cull=False, look=(0, 0, 1)):
    ms = api.MeshSet(meshes)
    oct_ = api.Octtree_Model(ms)
    sc = api.Scene(ctx)
    sc.set_model(oct_, cull_bits=oct_.compute_backface(look) if cull else None)
    sc.commit()
    return sc, oct_


@pytest.mark.parametrize("name,make,cull", [("heightfield", lambda: scenes.heightfield(96, with_light=False), False),
                                            ("soup_culled", lambda: scenes.random_soup(3000, seed=13), True)])
@pytest.mark.parametrize("mode", [0, api.DEFAULT_TRACE_MODE])
def test_closest_hit_ids_against_the_reference_traversal(gpu_ctx, name, make, cull, mode):
    meshes = make()
    look = (0.1, -0.2, 1)
    r = R.RefScene(); r.set_model(meshes, cull_backface=cull, look_dir=look); r.build_octree()
    g, oct_ = _gpu_scene(gpu_ctx, meshes, cull, look)
    w, h = 160, 90
    r2c, c2w = common.camera_1080p_like(w, h)
    rays = np.concatenate([common.pixel_center_rays(w, h, r2c, c2w), common.random_rays(6000, 31), common.random_rays(1500, 32, origin_box=5.0)])
    got = g.trace_closest(rays, mode=mode)
    ref = r.traverse_surface(rays, nthreads=8)
    hit = ref["found"] > 0
    assert np.array_equal(got["tri"] >= 0, hit) and 0.3 < hit.mean()
    # the record the reference returned == the record of the GPU's id, rebuilt by the reference's own Triangle code
    rec = r.surface_of(got["mesh"], got["tri"], rays)
    assert (rec["found"][hit] == 1).all()
    for k in ("n", "hitp", "uv"):
        assert np.array_equal(bits(rec[k][hit]), bits(ref[k][hit])), k
    # t and barycentrics of that id, computed by the reference's BasicIntersect, are the GPU's bit for bit
    ti = r.triangle_intersect(np.maximum(got["mesh"], 0), np.maximum(got["tri"], 0), rays, np.full(len(rays), common.FLT_MAX, np.float32))
    assert np.array_equal(bits(ti["t"][hit]), bits(got["t"][hit]))
    assert np.array_equal(bits(ti["bary"][hit]), bits(got["bary"][hit]))
    g.close(); oct_.close(); r.close()


def test_film_against_the_reference_renderer(gpu_ctx):
    """Tier A film, 192x108 @ 8 spp, against evaluate_pixel/Li compiled from the reference.  Weights are exact; radiance differs
    only through libm (atanh/cosh of the wavelength sampling: glibc on the host, libdevice on the GPU), tolerance as in
    tests/test_gpu_render_tier_a.py."""
    meshes = scenes.heightfield(128, with_light=False)
    w, h, spp = 192, 108, 8
    r = R.RefScene(); r.set_model(meshes); r.build_octree()
    g, oct_ = _gpu_scene(gpu_ctx, meshes)
    r2c, c2w = common.camera_1080p_like(w, h)
    rr2c, rc2w = R.camera_matrices(0, 1.0, 1000.0, 0.0, 0.0, 45.0, (0, 0, 0), (0, 0, 1), (1, 0, 0), (0, 1, 0), w, h)
    assert np.array_equal(bits(r2c), bits(rr2c)) and np.array_equal(bits(c2w), bits(rc2w))       # product camera == reference camera
    film = api.Film(gpu_ctx, w, h)
    g.render(film, api.make_config(w, h, r2c, c2w, sampler_kind=1, xs=4, ys=2, jitter=1, spp_begin=0, spp_end=spp))
    gf = film.download()
    rf = r.render_tier_a(R.make_params(w, h, sampler_kind=1, xs=4, ys=2, jitter=1, spp_begin=0, spp_end=spp, nthreads=8))
    assert np.array_equal(gf[:, 3], rf[:, 3])
    rmse = float(np.sqrt(np.mean((gf[:, :3] - rf[:, :3]) ** 2)))
    assert rmse < 2e-5 * spp, rmse
    diff = np.abs(gf[:, :3] - rf[:, :3])
    assert (diff > 1e-4).mean() < 2e-3 and diff.max() < 2e-2
    g8, _ = film.resolve()
    r8, _ = R.resolve(rf)
    assert np.abs(g8.astype(int) - r8.astype(int)).max() <= 1
    # per-sample: camera rays bit exact, wavelengths within 3 ulp
    rs = np.random.RandomState(0)
    pid = rs.randint(0, w * h, 4000).astype(np.int32); idx = rs.randint(0, 8, 4000).astype(np.int32)
    gs = g.eval_samples(api.make_config(w, h, r2c, c2w, sampler_kind=1, xs=4, ys=2, jitter=1), pid, idx)
    rs_ = r.eval_samples(R.make_params(w, h, sampler_kind=1, xs=4, ys=2, jitter=1), pid, idx)
    assert np.array_equal(bits(gs["ray"]), bits(rs_["ray"]))
    assert np.array_equal(bits(gs["weight"]), bits(rs_["weight"]))
    np.testing.assert_allclose(gs["lam"], rs_["lam"], rtol=4e-7, atol=0)
    film.close(); g.close(); oct_.close(); r.close()


@pytest.mark.parametrize("k", range(len(P.SHAPES)))
def test_analytic_shapes_against_the_reference(gpu_ctx, k):
    kind, params = P.SHAPES[k]
    rigid = P._rigid(10, -5, 500, ang=0.4 + 0.1 * k)
    g = api.Scene(gpu_ctx); gid = g.add_shape(kind, rigid, params); g.commit()
    r = R.RefScene(); rid = r.add_shape(kind, rigid, params)
    rays = common.random_rays(20000, 20 + k, center=(10, -5, 500), spread=90.0, origin_box=120.0)
    a = g.shape_intersect(gid, rays, common.FLT_MAX); b = r.shape_intersect(rid, rays, common.FLT_MAX)
    assert np.array_equal(a["found"], b["found"])
    f = b["found"] > 0
    partial = kind < 3 and params[3] < 360.0
    if not partial:
        for key in ("t", "hitp", "n"):
            assert np.array_equal(bits(a[key][f]), bits(b[key][f])), key
    else:
        np.testing.assert_allclose(a["t"][f], b["t"][f], rtol=1e-6)
        np.testing.assert_allclose(a["hitp"][f], b["hitp"][f], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(a["n"][f], b["n"][f], atol=1e-5)
    np.testing.assert_allclose(a["uv"][f], b["uv"][f], atol=2e-6)
    g.close(); r.close()
