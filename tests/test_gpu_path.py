"""GPU parity of the wavefront path integrator (Tier B) against the oracle's Renderer::LiPath with identical RNG streams.

Tolerances (DESIGN.md "floating-point tolerance"): surface records of triangle hits are bit-exact (IEEE-only arithmetic);
per-sample radiance agrees to 2e-4 relative for all but a small fraction of samples, because bounce directions pass through
cosf/sinf (libdevice vs glibc differ in the last ulp), which can move a later hit across a triangle edge or a wavelength
across a 1 nm table bin; converged films agree to a stated RMSE."""
import numpy as np
import pytest

import common
import oracle_lib as O
from common import ScenePair, bits
from computational_ray_tracer_b200 import api, scenes

pytestmark = pytest.mark.gpu


def _cornell(ctx, **kw):
    return ScenePair(ctx, scenes.cornell_box(), materials=lambda sc: scenes.cornell_materials(sc, **kw))


def _cfgs(w, h, r2c, c2w, **kw):
    return api.make_config(w, h, r2c, c2w, **kw), O.make_params(w, h, r2c, c2w, **kw)


def test_scene_closest_surface_record(gpu_ctx):
    pair = _cornell(gpu_ctx)
    r2c, c2w = common.camera_1080p_like(160, 160)
    rays = np.concatenate([common.pixel_center_rays(160, 160, r2c, c2w), common.random_rays(8000, 12, center=(0, 0, 650), spread=260, origin_box=200)])
    g = pair.gpu.scene_closest(rays); o = pair.orc.scene_closest(rays)
    assert np.array_equal(g["kind"], o["kind"])
    assert (o["kind"] == 0).any() and (o["kind"] == 1).any() and (o["kind"] == -1).any()
    hit = o["kind"] >= 0
    assert np.array_equal(g["id0"][hit], o["id0"][hit])
    tri = o["kind"] == 0
    assert np.array_equal(g["id1"][tri], o["id1"][tri])
    assert np.array_equal(g["backside"][hit], o["backside"][hit])
    for k in ("t", "p", "ns", "ng"):
        assert np.array_equal(bits(g[k][tri]), bits(o[k][tri])), k
    sph = o["kind"] == 1           # full spheres: IEEE-only arithmetic as well
    for k in ("t", "p", "ns", "ng"):
        assert np.array_equal(bits(g[k][sph]), bits(o[k][sph])), k
    pair.close()


def _sample_parity(pair, w, h, n, frac_ok, **kw):
    r2c, c2w = common.camera_1080p_like(w, h)
    gc, oc = _cfgs(w, h, r2c, c2w, mode=1, **kw)
    rs = np.random.RandomState(3)
    pix = rs.randint(0, w * h, n).astype(np.int32); idx = rs.randint(0, kw.get("xs", 4) * kw.get("ys", 4), n).astype(np.int32)
    g = pair.gpu.eval_samples(gc, pix, idx); o = pair.orc.eval_samples(oc, pix, idx)
    assert np.array_equal(bits(g["ray"]), bits(o["ray"]))
    np.testing.assert_allclose(g["lam"], o["lam"], rtol=1e-5)
    ok = np.isclose(g["L"], o["L"], rtol=2e-4, atol=1e-5).all(axis=1)
    assert ok.mean() >= frac_ok, ok.mean()
    assert (o["L"] > 0).any()
    return g, o


def test_per_sample_radiance_cornell(gpu_ctx):
    pair = _cornell(gpu_ctx)
    _sample_parity(pair, 96, 96, 6000, 0.99, xs=4, ys=4, max_depth=5)
    pair.close()


def test_per_sample_radiance_glass_dispersion_and_rr(gpu_ctx):
    pair = _cornell(gpu_ctx, glass=True)
    g, o = _sample_parity(pair, 96, 96, 6000, 0.98, xs=4, ys=4, max_depth=8, rr_depth=3)
    # TerminateSecondary (spectrum.h:302-310) happened for the same samples
    assert np.array_equal(g["pdf"][:, 1] == 0, o["pdf"][:, 1] == 0)
    assert (o["pdf"][:, 1] == 0).any()
    pair.close()


def test_per_sample_radiance_conductors(gpu_ctx):
    pair = ScenePair(gpu_ctx, scenes.spheres_lattice_meshes(), materials=lambda sc: scenes.spheres_lattice_materials(sc, n=4))
    _sample_parity(pair, 96, 54, 5000, 0.98, xs=4, ys=4, max_depth=8)
    pair.close()


@pytest.mark.parametrize("scene", ["cornell", "heightfield", "many_lights"])
def test_film_rmse_and_counters(gpu_ctx, scene):
    if scene == "cornell":
        pair = _cornell(gpu_ctx); w, h = 96, 96
    elif scene == "heightfield":
        pair = ScenePair(gpu_ctx, scenes.heightfield(96), materials=scenes.c2_materials); w, h = 128, 72
    else:
        pair = ScenePair(gpu_ctx, scenes.many_light_scene(64, 200), materials=scenes.many_light_materials); w, h = 128, 72
    r2c, c2w = common.camera_1080p_like(w, h)
    spp = 16
    gc, oc = _cfgs(w, h, r2c, c2w, mode=1, xs=4, ys=4, spp_begin=0, spp_end=spp, max_depth=5)
    oc.nthreads = 8
    film = api.Film(gpu_ctx, w, h)
    st = pair.gpu.render(film, gc)
    gf = film.download()
    orr = pair.orc.render(oc, counters=True)
    of, c = orr["film"], orr["counters"]
    assert st["paths"] == c["paths"] == w * h * spp
    assert np.array_equal(gf[:, 3], of[:, 3])
    for k in ("closest_rays", "shadow_rays", "depth_sum"):
        assert abs(st[k] - c[k]) <= 2e-3 * c[k] + 2, (k, st[k], c[k])
    mean_g, mean_o = gf[:, :3].mean(0) / spp, of[:, :3].mean(0) / spp
    np.testing.assert_allclose(mean_g, mean_o, rtol=2e-3)
    rmse = float(np.sqrt(np.mean(((gf[:, :3] - of[:, :3]) / spp) ** 2)))
    assert rmse < 5e-3, rmse                              # per-pixel mean sensor RGB in [0,1]
    assert of[:, :3].max() > 0
    film.close(); pair.close()


def test_partitions_give_the_single_gpu_film(gpu_ctx):
    pair = _cornell(gpu_ctx)
    w, h, spp = 64, 64, 8
    r2c, c2w = common.camera_1080p_like(w, h)
    kw = dict(mode=1, xs=4, ys=2, spp_begin=0, spp_end=spp, max_depth=4)
    film = api.Film(gpu_ctx, w, h)
    pair.gpu.render(film, api.make_config(w, h, r2c, c2w, **kw))
    full = film.download()
    for partition in (0, 1):
        film.clear()
        for rank in range(3):                                     # three "GPUs" accumulating into one film = reduce(sum)
            pair.gpu.render(film, api.make_config(w, h, r2c, c2w, rank=rank, world=3, partition=partition, tile=(16, 8), **kw))
        got = film.download()
        assert np.array_equal(bits(got), bits(full))              # same stream order per pixel -> identical sums
    film.close(); pair.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_ordered_traversal_renders_the_identical_film(gpu_ctx, mode):
    """The ordered traversal changes no hit, so films are bit-identical to the exact-BFS render (Tier A and path integrator)."""
    pair = ScenePair(gpu_ctx, scenes.heightfield(128), materials=scenes.c2_materials)
    w, h = 160, 90
    r2c, c2w = common.camera_1080p_like(w, h)
    films = []
    stats = []
    for tm in (0, 3):
        film = api.Film(gpu_ctx, w, h)
        stats.append(pair.gpu.render(film, api.make_config(w, h, r2c, c2w, mode=mode, xs=4, ys=2, spp_begin=0, spp_end=8, max_depth=4, trace_mode=tm, collect_stats=1)))
        films.append(film.download()); film.close()
    for k in (1,):
        assert np.array_equal(bits(films[0]), bits(films[k]))
        assert stats[0]["closest_rays"] == stats[k]["closest_rays"] and stats[0]["shadow_rays"] == stats[k]["shadow_rays"]
        assert stats[k]["tris_tested"] < stats[0]["tris_tested"]            # and it does less work
    pair.close()


def test_render_is_deterministic_and_resumable(gpu_ctx):
    """Queue compaction order varies run to run (atomics) but no result may depend on it: two renders give identical bits.
    Resume (checkpoint/resume): film saved after sample index k + render of [k, n) == render of [0, n)."""
    pair = _cornell(gpu_ctx, glass=True)
    w, h = 80, 80
    r2c, c2w = common.camera_1080p_like(w, h)
    kw = dict(mode=1, xs=4, ys=4, max_depth=6, rr_depth=3, trace_mode=3)
    film = api.Film(gpu_ctx, w, h)
    pair.gpu.render(film, api.make_config(w, h, r2c, c2w, spp_begin=0, spp_end=12, **kw))
    a = film.download()
    film.clear()
    pair.gpu.render(film, api.make_config(w, h, r2c, c2w, spp_begin=0, spp_end=12, **kw))
    assert np.array_equal(bits(a), bits(film.download()))
    film.clear()
    pair.gpu.render(film, api.make_config(w, h, r2c, c2w, spp_begin=0, spp_end=5, **kw))
    saved = film.download().copy()
    film2 = api.Film(gpu_ctx, w, h)
    film2.upload(saved)                                   # "restart the process": new film object, restored state
    pair.gpu.render(film2, api.make_config(w, h, r2c, c2w, spp_begin=5, spp_end=12, **kw))
    assert np.array_equal(bits(a), bits(film2.download()))
    film.close(); film2.close(); pair.close()


@pytest.mark.parametrize("rho,depth", [(0.5, 4), (0.8, 3)])
def test_white_furnace_on_the_device(gpu_ctx, rho, depth):
    """Closed box, every wall emits L_e and reflects rho: radiance = L_e (1 + rho + ... + rho^D) (tests/test_cpu_furnace.py states the
    argument for the oracle); here for the wavefront integrator, with the same sampler streams."""
    from test_cpu_furnace import _closed_box
    le = 0.25

    def mats(sc):
        refl = sc.add_spectrum(0, c=rho); emit = sc.add_spectrum(0, c=1.0)
        m = sc.add_material(type=0, refl=refl, emit=emit, emit_scale=le)
        return [m] * 6
    pair = ScenePair(gpu_ctx, _closed_box(), materials=mats)
    w = h = 24
    r2c, c2w = api.camera_matrices(0, 1.0, 1000.0, 45.0, (3, -2, 5), (0.2, 0.1, 1), (0, 1, 0), w, h)
    kw = dict(mode=1, xs=8, ys=8, jitter=1, max_depth=depth, rr_depth=0)
    pid = np.repeat(np.arange(w * h, dtype=np.int32), 64); idx = np.tile(np.arange(64, dtype=np.int32), w * h)
    g = pair.gpu.eval_samples(api.make_config(w, h, r2c, c2w, **kw), pid, idx)["L"]
    want = le * sum(rho ** k for k in range(depth + 1))
    got = float(g.astype(np.float64).mean())
    assert abs(got - want) / want < 0.01, (got, want)
    o = pair.orc.eval_samples(O.make_params(w, h, r2c, c2w, **kw), pid, idx)["L"]
    assert np.isclose(g, o, rtol=2e-4, atol=1e-6).all(axis=1).mean() > 0.98
    pair.close()


def test_shape_hierarchy_returns_the_list_order_answer(gpu_ctx):
    """The analytic shapes are visited through a threaded BVH instead of the oracle's list-order loop: the closest shape, its surface record
    and occlusion must be the oracle's on the C3 lattice (64 spheres), on overlapping shapes of all four kinds, and -- the one case where
    the visiting order could matter -- on exactly coincident shapes, where the lowest list index must win."""
    def both(build):
        orc = O.OracleScene(); gpu = api.Scene(gpu_ctx)
        for sc in (orc, gpu):
            build(sc)
        gpu.commit()
        return orc, gpu

    def lattice(sc):
        scenes.spheres_lattice_materials(sc)

    def mixed(sc):
        g = sc.add_spectrum(0, c=0.5); m = sc.add_material(type=0, refl=g)
        rs = np.random.RandomState(5)
        for i in range(40):
            t = scenes.translation(*rs.uniform(-150, 150, 2), rs.uniform(450, 750))
            kind = i % 4
            params = [[40.0, -40.0, 40.0, 360.0], [25.0, -30.0, 30.0, 360.0], [0.0, 0.0, 45.0, 360.0], [-40, -30, 0, 45, -20, 5, 0, 50, -8]][kind]
            sc.add_shape(kind, t, params, material=m)

    def coincident(sc):
        g = sc.add_spectrum(0, c=0.5); m = sc.add_material(type=0, refl=g)
        for i in range(6):                       # the same sphere six times, then the same triangle three times in front of half of it
            sc.add_shape(0, scenes.translation(0, 0, 600), [80.0, -80.0, 80.0, 360.0], material=m)
        for i in range(3):
            sc.add_shape(3, scenes.translation(0, 0, 480), [-200, -200, 0, 200, -200, 0, 0, 200, 0], material=m)

    for build, center, spread in ((lattice, (0, 0, 650), 260), (mixed, (0, 0, 600), 220), (coincident, (0, 0, 600), 120)):
        orc, gpu = both(build)
        rays = common.random_rays(30000, 21, center=center, spread=spread, origin_box=150)
        g = gpu.scene_closest(rays); o = orc.scene_closest(rays)
        assert np.array_equal(g["kind"], o["kind"]) and np.array_equal(g["id0"], o["id0"]), build.__name__
        hit = o["kind"] == 1
        assert 0.1 < hit.mean()
        assert np.array_equal(bits(g["t"][hit]), bits(o["t"][hit])) and np.array_equal(g["backside"][hit], o["backside"][hit])
        np.testing.assert_allclose(g["p"][hit], o["p"][hit], rtol=0, atol=2e-3)          # partial sweeps / cylinders pass through atan2
        if build is coincident:
            assert set(np.unique(o["id0"][hit])) <= {0, 6}                                  # ties go to the first of the coincident copies
        gpu.close(); orc.close()


def test_small_frames_replayed_from_a_cuda_graph_give_the_same_film(gpu_ctx):
    """C1-sized frames are launch bound: their full waves are captured once into a CUDA graph and replayed (crt_render_stats.graph_launches).
    The film must be bit-identical to the ordinary launch sequence (time_kernels = 1 keeps the render off the graph path), also when the
    same graph is reused for another sample range and after the scene is committed again."""
    pair = _cornell(gpu_ctx, glass=True)
    w, h = 96, 96
    r2c, c2w = common.camera_1080p_like(w, h)
    kw = dict(mode=1, xs=16, ys=16, max_depth=6, rr_depth=3, trace_mode=3)
    film = api.Film(gpu_ctx, w, h)
    films = {}
    for name, extra in (("graph", {}), ("plain", dict(time_kernels=1))):
        film.clear()
        st = pair.gpu.render(film, api.make_config(w, h, r2c, c2w, spp_begin=0, spp_end=150, **kw, **extra))     # 2 full waves of 64 + a tail of 22
        films[name] = film.download()
        assert (st["graph_launches"] == 2) == (name == "graph"), st
        assert st["paths"] == 150 * w * h
    assert np.array_equal(bits(films["graph"]), bits(films["plain"])) and films["plain"][:, :3].max() > 0
    # the cached graph serves another range; a re-commit invalidates it
    film.clear()
    pair.gpu.render(film, api.make_config(w, h, r2c, c2w, spp_begin=0, spp_end=11, **kw))
    st = pair.gpu.render(film, api.make_config(w, h, r2c, c2w, spp_begin=11, spp_end=150, **kw))
    assert st["graph_launches"] == 2
    assert np.array_equal(bits(film.download()), bits(films["plain"]))
    pair.gpu.commit()
    film.clear()
    pair.gpu.render(film, api.make_config(w, h, r2c, c2w, spp_begin=0, spp_end=150, **kw))
    assert np.array_equal(bits(film.download()), bits(films["plain"]))
    film.close(); pair.close()


def _lattice_with_lights(sc, n):
    mats = scenes.spheres_lattice_materials(sc, n=n)
    white = sc.add_spectrum(0, c=1.0)
    sc.add_light(0, (40.0, -30.0, 420.0), white, 3.0e5)            # a point light: one additional next-event slot per bounce
    return mats


@pytest.mark.parametrize("scene", ["lattice", "lattice+point_light+each_light", "cornell_glass"])
def test_staged_shading_renders_the_identical_film(gpu_ctx, scene):
    """crt_render_config.shade_mode: the staged bounce (surface-record kernel, then one kernel per material type over that type's queue;
    the default when the scene has analytic shapes) performs, per path, the fused kernel's operations in the fused kernel's order -- only
    the order of the queues differs.  Films, ray counts and per-sample radiance must be identical bit for bit, with all three material
    types, with additional next-event slots (which read the staged surface record), through Russian roulette and from a CUDA graph."""
    if scene == "cornell_glass":
        pair = _cornell(gpu_ctx, glass=True)
        extra = {}
    elif scene == "lattice":
        pair = ScenePair(gpu_ctx, scenes.spheres_lattice_meshes(), materials=lambda sc: scenes.spheres_lattice_materials(sc, n=4))
        extra = {}
    else:
        pair = ScenePair(gpu_ctx, scenes.spheres_lattice_meshes(), materials=lambda sc: _lattice_with_lights(sc, 4))
        extra = dict(light_strategy=1)
    w, h = 128, 72
    r2c, c2w = common.camera_1080p_like(w, h)
    kw = dict(mode=1, xs=4, ys=4, max_depth=10, rr_depth=3, spp_begin=0, spp_end=16, **extra)
    films, stats = {}, {}
    for name, sm, more in (("fused", 1, dict(time_kernels=1)), ("staged", 2, dict(time_kernels=1)), ("auto", 0, {})):
        film = api.Film(gpu_ctx, w, h)
        stats[name] = pair.gpu.render(film, api.make_config(w, h, r2c, c2w, shade_mode=sm, **kw, **more))
        films[name] = film.download(); film.close()
    assert films["fused"][:, :3].max() > 0
    for name in ("staged", "auto"):
        assert np.array_equal(bits(films["fused"]), bits(films[name])), name
        for k in ("paths", "closest_rays", "shadow_rays", "depth_sum"):
            assert stats["fused"][k] == stats[name][k], (name, k)
    assert stats["staged"]["kernel_launches"] > stats["fused"]["kernel_launches"]          # it really took the other route
    # the per-sample probe goes through the same wave code
    pix = np.arange(0, w * h, 5, dtype=np.int32); idx = (pix % 16).astype(np.int32)
    a = pair.gpu.eval_samples(api.make_config(w, h, r2c, c2w, shade_mode=1, **kw), pix, idx)
    b = pair.gpu.eval_samples(api.make_config(w, h, r2c, c2w, shade_mode=2, **kw), pix, idx)
    assert np.array_equal(bits(a["L"]), bits(b["L"])) and np.array_equal(bits(a["pdf"]), bits(b["pdf"]))
    pair.close()


@pytest.mark.parametrize("scene,shade_mode", [("cornell_glass", 1), ("lattice", 1), ("lattice", 2)])
def test_root_leaf_models_are_traversed_inside_the_shading_kernels(gpu_ctx, scene, shade_mode):
    """An octree that is one root leaf (Cornell box: 12 triangles, the floor and light under the sphere lattice: 4) needs no traversal launch:
    under the production trace mode the shading kernels run the reference loop for that leaf themselves (trace_root_leaf).  trace_mode 0 keeps
    the exact BFS kernel's launches, so the two renders cross-check each other: identical films, ray counts and per-sample radiance; shadow
    rays included (the occlusion test moves into k_shadow_resolve)."""
    if scene == "cornell_glass":
        pair = _cornell(gpu_ctx, glass=True)
    else:
        pair = ScenePair(gpu_ctx, scenes.spheres_lattice_meshes(), materials=lambda sc: scenes.spheres_lattice_materials(sc, n=4))
    assert pair.oct.stats()["nodes"] == 1
    w, h = 128, 72
    r2c, c2w = common.camera_1080p_like(w, h)
    kw = dict(mode=1, xs=4, ys=4, max_depth=8, rr_depth=3, spp_begin=0, spp_end=16, shade_mode=shade_mode, time_kernels=1)
    films, stats = {}, {}
    for tm in (0, 3):
        film = api.Film(gpu_ctx, w, h)
        stats[tm] = pair.gpu.render(film, api.make_config(w, h, r2c, c2w, trace_mode=tm, **kw))
        films[tm] = film.download(); film.close()
    assert films[0][:, :3].max() > 0 and np.array_equal(bits(films[0]), bits(films[3]))
    for k in ("paths", "closest_rays", "shadow_rays", "depth_sum"):
        assert stats[0][k] == stats[3][k], k
    assert stats[3]["kernel_launches"] < stats[0]["kernel_launches"] and stats[3]["trace_ms"] > 0
    pix = np.arange(0, w * h, 3, dtype=np.int32); idx = (pix % 16).astype(np.int32)
    a = pair.gpu.eval_samples(api.make_config(w, h, r2c, c2w, trace_mode=0, **kw), pix, idx)
    b = pair.gpu.eval_samples(api.make_config(w, h, r2c, c2w, trace_mode=3, **kw), pix, idx)
    assert np.array_equal(bits(a["L"]), bits(b["L"]))
    pair.close()
