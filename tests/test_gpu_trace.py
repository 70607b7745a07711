"""GPU parity: closest-hit ids of the CUDA traversal vs the oracle's BFS octree traversal -- bit exact."""
import numpy as np
import pytest

import common
from common import ScenePair, bits
from computational_ray_tracer_b200 import scenes

pytestmark = pytest.mark.gpu


def _compare_hits(pair, rays, nthreads=8, mode=0):
    g = pair.gpu.trace_closest(rays, mode=mode)
    o = pair.orc.trace(rays, 0, nthreads=nthreads)
    assert np.array_equal(g["mesh"], o["mesh"]), f"mesh id mismatches: {(g['mesh'] != o['mesh']).sum()} of {len(rays)}"
    assert np.array_equal(g["tri"], o["tri"]), f"tri id mismatches: {(g['tri'] != o['tri']).sum()} of {len(rays)}"
    hit = o["tri"] >= 0
    assert np.array_equal(bits(g["t"][hit]), bits(o["t"][hit]))
    assert np.array_equal(bits(g["bary"][hit]), bits(o["bary"][hit]))
    return hit.mean()


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("name", ["heightfield", "soup", "cornell", "axis_grid"])
def test_closest_hit_ids_bit_exact(gpu_ctx, name, mode):
    meshes = {"heightfield": lambda: scenes.heightfield(160), "soup": lambda: scenes.random_soup(4000),
              "cornell": scenes.cornell_box, "axis_grid": lambda: scenes.axis_grid(32, layers=3)}[name]()
    pair = ScenePair(gpu_ctx, meshes)
    r2c, c2w = common.camera_1080p_like(480, 270)
    rays = np.concatenate([common.pixel_center_rays(480, 270, r2c, c2w), common.random_rays(20000, 3)])
    frac = _compare_hits(pair, rays, mode=mode)
    assert frac > 0.05
    pair.close()


def test_closest_hit_with_backface_culling(gpu_ctx):
    pair = ScenePair(gpu_ctx, scenes.random_soup(3000, seed=11), cull=True)
    rays = common.random_rays(30000, 5)
    _compare_hits(pair, rays)
    _compare_hits(pair, rays, mode=3)
    pair.close()


def test_degenerate_axis_parallel_and_empty(gpu_ctx):
    pair = ScenePair(gpu_ctx, scenes.axis_grid(16, layers=2))
    # axis-parallel directions (1/d = inf), rays starting on geometry, rays that miss everything
    rays = np.array([[0, 0, 0, 0, 0, 1], [0.5, 0.25, 0, 0, 0, 1], [10, 10, 500, 0, 0, 1], [0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 1, 0],
                     [-300, 0, 500, 1, 0, 0], [0, 0, 1000, 0, 0, -1], [0, 0, 0, 0, 0, -1], [15, 15, 0, 0, 0, 1], [30, -30, 100, 0, 0, 1]], np.float32)
    _compare_hits(pair, rays, nthreads=1)
    _compare_hits(pair, rays, nthreads=1, mode=3)
    assert pair.gpu.trace_closest(np.zeros((0, 6), np.float32))["tri"].shape == (0,)
    pair.close()


def test_any_hit_matches_oracle(gpu_ctx):
    pair = ScenePair(gpu_ctx, scenes.random_soup(3000, seed=2))
    rays = common.random_rays(20000, 9)
    tmax = np.random.RandomState(1).uniform(50, 900, len(rays)).astype(np.float32)
    g = pair.gpu.trace_any(rays, tmax)
    o = pair.orc.trace(rays, 2, tmax=tmax, nthreads=8)["mesh"]
    assert np.array_equal(g, o)
    assert np.array_equal(pair.gpu.trace_any(rays, tmax, mode=3), o)
    assert 0.05 < g.mean() < 0.95
    pair.close()


def test_traverse_surface_normal(gpu_ctx):
    for meshes in (scenes.heightfield(64, with_light=False), [dict(m, normals=None) for m in scenes.random_soup(500)]):
        pair = ScenePair(gpu_ctx, meshes)
        rays = common.random_rays(5000, 4, center=(0, 0, 700), spread=150)
        g = pair.gpu.traverse_surface(rays)
        o = pair.orc.traverse_surface(rays)
        assert np.array_equal(g["found"], o["found"])
        f = o["found"] > 0
        assert f.any()
        assert np.array_equal(bits(g["n"][f]), bits(o["n"][f]))
        pair.close()


def test_ordered_traversal_equals_exact_bfs_at_scale(gpu_ctx):
    """trace_mode 3 (ordered traversal + exact re-trace of order-sensitive rays) must return what the exact BFS kernel
    returns for every ray -- ids, t and barycentrics -- including rays aimed at shared edges and vertices."""
    from computational_ray_tracer_b200 import api
    meshes = scenes.heightfield(300)
    ms = api.MeshSet(meshes); oc = api.Octtree_Model(ms)
    sc = api.Scene(gpu_ctx); sc.set_model(oc); sc.commit()
    r2c, c2w = common.camera_1080p_like(960, 540)
    pos = meshes[0]["positions"]
    rs = np.random.RandomState(0)
    vi = rs.randint(0, len(pos), 60000)
    vdir = pos[vi] / np.linalg.norm(pos[vi], axis=1, keepdims=True)                    # rays through mesh vertices: 6-way ties
    ei = rs.randint(0, len(pos) - 1, 60000)
    mid = 0.5 * (pos[ei] + pos[ei + 1]); edir = mid / np.linalg.norm(mid, axis=1, keepdims=True)   # through edge midpoints
    extra = np.concatenate([np.zeros((120000, 3), np.float32), np.concatenate([vdir, edir]).astype(np.float32)], 1)
    rays = np.concatenate([common.pixel_center_rays(960, 540, r2c, c2w), common.random_rays(200000, 3, center=(0, 0, 800), spread=400), extra])
    a = sc.trace_closest(rays, mode=0)
    for mode in (3,):
        b = sc.trace_closest(rays, mode=mode)
        for k in ("mesh", "tri"):
            assert np.array_equal(a[k], b[k]), (mode, k, int((a[k] != b[k]).sum()))
        assert np.array_equal(bits(a["t"]), bits(b["t"])) and np.array_equal(bits(a["bary"]), bits(b["bary"]))
    assert (a["tri"] >= 0).mean() > 0.3
    sc.close(); oc.close()


def test_exact_ties_everywhere_are_resolved_like_the_reference(gpu_ctx):
    """Every triangle duplicated (coincident copies in one mesh and in a second mesh) plus sheets meeting along shared
    edges: every hit is an exact tie, i.e. every ray is order-sensitive and must come out of the exact re-trace pass with
    the id the reference's BFS order picks."""
    g = scenes.axis_grid(24, layers=2)[0]
    dup = scenes.merge_meshes([g, g])                                  # coincident copies inside one mesh
    meshes = [dup, dict(g, positions=g["positions"].copy())]            # and a third copy as a second mesh
    pair = ScenePair(gpu_ctx, meshes)
    r2c, c2w = common.camera_1080p_like(320, 180)
    rays = np.concatenate([common.pixel_center_rays(320, 180, r2c, c2w), common.random_rays(30000, 17, center=(0, 0, 520), spread=260)])
    for mode in (0, 3):
        frac = _compare_hits(pair, rays, mode=mode)
    assert frac > 0.3
    # the film is identical too, and the statistics show the hand-over actually happened
    w, h = 160, 90
    r2c, c2w = common.camera_1080p_like(w, h)
    films = []
    for tm in (0, 3):
        film = api_film(gpu_ctx, w, h)
        st = pair.gpu.render(film, api_cfg(w, h, r2c, c2w, mode=0, xs=2, ys=2, spp_begin=0, spp_end=4, trace_mode=tm))
        films.append(film.download()); film.close()
        if tm == 3:
            assert st["exact_retraced_rays"] > 0.3 * st["closest_rays"]
    assert np.array_equal(bits(films[0]), bits(films[1]))
    pair.close()


def api_film(ctx, w, h):
    from computational_ray_tracer_b200 import api
    return api.Film(ctx, w, h)


def api_cfg(*a, **kw):
    from computational_ray_tracer_b200 import api
    return api.make_config(*a, **kw)
