"""Edge cases of the render entry points, each against the oracle: shapes-only scenes, empty models, tiny and ragged
frames, depth 0, sample ranges that do not start at 0, thin lens with the path integrator, back-face culling, error paths."""
import ctypes as C

import numpy as np
import pytest

import common
import oracle_lib as O
from common import ScenePair, bits
from computational_ray_tracer_b200 import _capi, api, scenes

pytestmark = pytest.mark.gpu


def _films(pair_gpu, pair_orc, ctx, w, h, r2c, c2w, **kw):
    film = api.Film(ctx, w, h)
    st = pair_gpu.render(film, api.make_config(w, h, r2c, c2w, **kw))
    gf = film.download(); film.close()
    of = pair_orc.render(O.make_params(w, h, r2c, c2w, nthreads=4, **kw))["film"]
    return gf, of, st


def _close(gf, of, spp, tol=5e-3):
    assert np.array_equal(gf[:, 3], of[:, 3])
    rmse = float(np.sqrt(np.mean(((gf[:, :3] - of[:, :3]) / max(spp, 1)) ** 2)))
    assert rmse < tol, rmse


def test_shapes_only_scene(gpu_ctx):
    """No triangle model at all: spheres + a TriangleSimple floor + an emissive TriangleSimple cannot be lights (only mesh
    triangles are sampled), so radiance comes from hitting the emitter directly through specular chains."""
    def build(sc):
        grey = sc.add_spectrum(3, n=scenes.SWATCH["grey"]); d65 = sc.add_spectrum(4, n=2); bk7 = sc.add_spectrum(2, name="glass_bk7")
        m_floor = sc.add_material(type=0, refl=grey)
        m_glass = sc.add_material(type=1, eta=bk7, eta_constant=1)
        m_emit = sc.add_material(type=0, refl=-1, emit=d65, emit_scale=5.0, two_sided=1)
        ident = scenes.COLUMN_MAJOR_IDENTITY
        sc.add_shape(3, ident, [-300, -200, 900, 300, -200, 900, 0, 300, 900], material=m_emit)
        sc.add_shape(0, scenes.translation(0, 0, 500), [70.0, -70.0, 70.0, 360.0], material=m_glass)
        sc.add_shape(2, scenes.translation(0, -90, 500), [0.0, 0.0, 200.0, 360.0], material=m_floor)
    g = api.Scene(gpu_ctx); build(g); g.commit()
    o = O.OracleScene(); build(o)
    w, h = 64, 48
    r2c, c2w = common.camera_1080p_like(w, h)
    gf, of, st = _films(g, o, gpu_ctx, w, h, r2c, c2w, mode=1, xs=2, ys=2, spp_begin=0, spp_end=4, max_depth=6)
    _close(gf, of, 4)
    assert of[:, :3].max() > 0 and st["shadow_rays"] == 0
    g.close(); o.close()


def test_empty_model_and_all_miss(gpu_ctx):
    tri = [dict(positions=np.float32([[0, 0, -500], [10, 0, -500], [0, 10, -500]]), normals=np.float32([[0, 0, 1]] * 3), indices=np.uint32([[0, 1, 2]]))]
    pair = ScenePair(gpu_ctx, tri, materials=lambda sc: [sc.add_material(type=0, refl=sc.add_spectrum(0, c=0.5))])
    w, h = 33, 17                                         # ragged frame: not a multiple of any tile or block size
    r2c, c2w = common.camera_1080p_like(w, h)
    for mode in (0, 1):
        gf, of, st = _films(pair.gpu, pair.orc, gpu_ctx, w, h, r2c, c2w, mode=mode, xs=2, ys=2, spp_begin=0, spp_end=3)
        assert np.array_equal(bits(gf), bits(of))         # everything misses (the triangle is behind the camera): exact zeros + weights
        assert gf[:, :3].max() == 0 and st["paths"] == w * h * 3
    pair.close()


@pytest.mark.parametrize("case", ["one_pixel", "depth0", "spp_offset", "lens", "culling", "independent_sampler"])
def test_path_integrator_variants(gpu_ctx, case):
    cull = case == "culling"
    pair = ScenePair(gpu_ctx, scenes.cornell_box(), cull=cull, look=(0, 0, 1), materials=lambda sc: scenes.cornell_materials(sc))
    w, h = (1, 1) if case == "one_pixel" else (72, 56)
    r2c, c2w = common.camera_1080p_like(w, h)
    kw = dict(mode=1, xs=4, ys=2, spp_begin=0, spp_end=8, max_depth=5)
    if case == "depth0":
        kw["max_depth"] = 0
    if case == "spp_offset":
        kw.update(spp_begin=3, spp_end=8)
    if case == "lens":
        kw.update(lens_radius=4.0, focal_distance=650.0)
    if case == "independent_sampler":
        kw.update(sampler_kind=0)
    gf, of, st = _films(pair.gpu, pair.orc, gpu_ctx, w, h, r2c, c2w, **kw)
    _close(gf, of, kw["spp_end"] - kw["spp_begin"], tol=2e-2 if case == "one_pixel" else 5e-3)
    if case == "depth0":
        assert st["shadow_rays"] == 0 and st["closest_rays"] == st["paths"]
    pair.close()


def test_error_paths(gpu_ctx, crt_lib):
    pair = ScenePair(gpu_ctx, scenes.cornell_box())                   # no materials
    w, h = 16, 16
    r2c, c2w = common.camera_1080p_like(w, h)
    film = api.Film(gpu_ctx, w, h)
    with pytest.raises(_capi.CrtError, match="materials"):
        pair.gpu.render(film, api.make_config(w, h, r2c, c2w, mode=1))
    with pytest.raises(_capi.CrtError, match="film size"):
        pair.gpu.render(film, api.make_config(w + 1, h, r2c, c2w))
    with pytest.raises(_capi.CrtError, match="non-grey"):
        pair.gpu.render(film, api.make_config(w, h, r2c, c2w, albedo=(0.2, 0.5, 0.8)))
    with pytest.raises(_capi.CrtError, match="jitter"):
        pair.gpu.render(film, api.make_config(w, h, r2c, c2w, xs=2, ys=2, jitter=0, spp_end=5))
    with pytest.raises(_capi.CrtError, match="trace_mode"):
        pair.gpu.render(film, api.make_config(w, h, r2c, c2w, trace_mode=7))
    g = api.Scene(gpu_ctx)
    with pytest.raises(_capi.CrtError, match="not committed"):
        g.trace_closest(np.zeros((1, 6), np.float32))
    g.close(); film.close(); pair.close()


def test_the_library_runs_on_the_stream_it_is_given(gpu_ctx):
    """crt_context_set_stream: work must be ordered on the caller's stream (bench.py relies on it for its events, fills and the reduce).
    A film clear queued behind a 0.5 s spin on that stream must not have happened when the call returns, and must have after the stream
    drains; with the context's own stream restored (NULL) the clear does not wait for the spin."""
    import time
    import torch
    dev = torch.device("cuda", 0)
    w, h = 64, 32
    film_t = torch.ones(w * h * 4, dtype=torch.float32, device=dev)
    film = api.Film(gpu_ctx, w, h)
    film.attach(film_t.data_ptr())
    side = torch.cuda.Stream(dev); probe = torch.cuda.Stream(dev)
    spin = int(0.5 * 1.9e9)
    # Everything the measurement uses runs once beforehand: the first launch of a kernel loads its module lazily, and a module load
    # waits for running kernels - which would make the probe below wait for the spin and read the film after the clear.
    torch.cuda._sleep(1000)
    with torch.cuda.stream(probe):
        float(film_t.sum().item())
    with torch.cuda.stream(side):
        float(film_t.sum().item())
    film.clear()
    gpu_ctx.synchronize()
    film_t.fill_(1.0)
    torch.cuda.synchronize(dev)
    try:
        gpu_ctx.set_stream(side.cuda_stream)
        with torch.cuda.stream(side):
            torch.cuda._sleep(spin)
        film.clear()                                            # asynchronous: queued on `side`, behind the spin
        with torch.cuda.stream(probe):
            early = float(film_t.sum().item())
        side.synchronize()
        late = float(film_t.sum().item())
        assert (early, late) == (w * h * 4.0, 0.0), f"clear not ordered on the caller's stream: sum {early} before the stream drained, {late} after"
        # back on the context's own stream the same call is independent of `side`
        film_t.fill_(1.0)
        torch.cuda.synchronize(dev)
        gpu_ctx.set_stream(0)
        with torch.cuda.stream(side):
            torch.cuda._sleep(spin)
        t0 = time.perf_counter()
        film.clear()
        gpu_ctx.synchronize()
        dt = time.perf_counter() - t0
        with torch.cuda.stream(probe):
            own = float(film_t.sum().item())
        assert own == 0.0 and dt < 0.25, f"on its own stream the clear took {dt:.3f} s (sum {own}): it waited for the other stream's 0.5 s spin"
    finally:
        gpu_ctx.set_stream(0)
        torch.cuda.synchronize(dev)
        film.close()
