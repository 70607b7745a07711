"""SURVEY 8 Tier-B row TB5: point and sun lights (RayTracer/Lights.h:5-8: "points light: position, color, and r^2 falloff",
"sunlight: direction, color") and the "1 sample from each light source" rule of RayTracer/Shading.h:4.

The reference only names these in comments, so -- like the rest of Tier B -- they are DEFINED by the oracle (oracle/oracle_render.cpp,
DeltaLight / light_strategy; parity unpinned by the reference) and checked two ways: physically (closed forms for a point light and a sun
over a Lambert plane; the one-sample-from-each estimator against the power-CDF estimator) and GPU against oracle per sample."""
import numpy as np
import pytest

import common
import oracle_lib as O
from computational_ray_tracer_b200 import api, scenes


def _plane(z=600.0, half=2000.0):
    return [scenes.quad_mesh((-half, -half, z), (half, -half, z), (-half, half, z), (half, half, z), (0, 0, -1))]


def _lit_plane(sc, rho, lights):
    refl = sc.add_spectrum(0, c=rho); white = sc.add_spectrum(0, c=1.0)
    m = sc.add_material(type=0, refl=refl)
    for kind, v, scale in lights:
        sc.add_light(kind, v, white, scale)
    return [m]


def test_point_and_sun_irradiance_closed_form(oracle):
    """One bounce (max_depth 1), a constant-spectrum Lambert plane, no occluders: L = (rho / pi) * (I cos / r^2 + E cos_sun) at every
    wavelength, exactly what Lights.h's "r^2 falloff" and Shading.h's "(r/pi)*lightcolor*cos(theta)" say."""
    rho, I, E = 0.6, 4.0e5, 0.8
    P = np.array([150.0, -80.0, 250.0]); sun = np.array([0.3, 0.2, -1.0]); sun /= np.linalg.norm(sun)
    orc = O.OracleScene(); orc.set_model(_plane()); orc.build_octree()
    orc.set_mesh_materials(_lit_plane(orc, rho, [(0, P, I), (1, sun, E)]))
    w, h = 64, 36
    r2c, c2w = common.camera_1080p_like(w, h)
    p = O.make_params(w, h, r2c, c2w, mode=1, xs=1, ys=1, jitter=0, max_depth=1)
    pix = np.arange(0, w * h, 7, dtype=np.int32)
    s = orc.eval_samples(p, pix, np.zeros_like(pix))
    ray = s["ray"].astype(np.float64)
    t = (600.0 - ray[:, 2]) / ray[:, 5]
    x = ray[:, :3] + t[:, None] * ray[:, 3:]
    n = np.array([0, 0, -1.0])
    d = P - x; r2 = (d * d).sum(1); cos_p = (d @ n) / np.sqrt(r2)
    want = rho / np.pi * (I * cos_p / r2 + E * (sun @ n))
    assert np.allclose(s["L"], want[:, None], rtol=3e-4), float(np.abs(s["L"] / want[:, None] - 1).max())
    orc.close()


def _cornell_with_lights(sc, extra):
    mats = scenes.cornell_materials(sc)
    white = sc.add_spectrum(0, c=1.0)
    for kind, v, scale in extra:
        sc.add_light(kind, v, white, scale)
    return mats


def test_one_sample_from_each_light_is_the_same_estimator_in_expectation(oracle):
    """Strategy 1 (Shading.h:4, every emissive triangle once) and strategy 0 (one triangle by the power CDF) estimate the same integral:
    converged Cornell-box films agree; strategy 1 traces about twice the shadow rays (two emissive triangles)."""
    orc = O.OracleScene(); orc.set_model(scenes.cornell_box()); orc.build_octree()
    orc.set_mesh_materials(scenes.cornell_materials(orc))
    w, h, spp = 48, 48, 64
    r2c, c2w = common.camera_1080p_like(w, h)
    films, shadow = [], []
    for strat in (0, 1):
        r = orc.render(O.make_params(w, h, r2c, c2w, mode=1, xs=8, ys=8, spp_end=spp, max_depth=3, nthreads=8, light_strategy=strat), counters=True)
        films.append(r["film"][:, :3].mean(0) / spp); shadow.append(r["counters"]["shadow_rays"])
    np.testing.assert_allclose(films[0], films[1], rtol=0.02)
    assert 1.7 < shadow[1] / shadow[0] < 2.1
    orc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("strategy", [0, 1])
def test_gpu_matches_the_oracle_with_point_and_sun_lights(gpu_ctx, strategy):
    """Cornell box + spheres, its area light, one point light inside the box and a sun through the open front: per-sample radiance and the
    film against the oracle, for both light strategies (bounce 0..5, identical RNG streams)."""
    extra = [(0, (120.0, 100.0, 520.0), 2.5e5), (1, (0.2, 0.5, -1.0), 0.6)]
    pair = common.ScenePair(gpu_ctx, scenes.cornell_box(), materials=lambda sc: _cornell_with_lights(sc, extra))
    w, h = 96, 96
    r2c, c2w = common.camera_1080p_like(w, h)
    kw = dict(mode=1, xs=4, ys=4, max_depth=5, light_strategy=strategy)
    rs = np.random.RandomState(2)
    pix = rs.randint(0, w * h, 5000).astype(np.int32); idx = rs.randint(0, 16, 5000).astype(np.int32)
    g = pair.gpu.eval_samples(api.make_config(w, h, r2c, c2w, trace_mode=3, **kw), pix, idx)
    o = pair.orc.eval_samples(O.make_params(w, h, r2c, c2w, **kw), pix, idx)
    ok = np.isclose(g["L"], o["L"], rtol=2e-4, atol=1e-5).all(axis=1)
    assert ok.mean() >= 0.98, ok.mean()
    assert (o["L"].max(1) > 0).mean() > 0.5
    film = api.Film(gpu_ctx, w, h)
    spp = 16
    st = pair.gpu.render(film, api.make_config(w, h, r2c, c2w, spp_begin=0, spp_end=spp, trace_mode=3, **kw))
    orr = pair.orc.render(O.make_params(w, h, r2c, c2w, spp_begin=0, spp_end=spp, nthreads=8, **kw), counters=True)
    gf, of, k = film.download(), orr["film"], orr["counters"]
    assert np.array_equal(gf[:, 3], of[:, 3])
    for key in ("closest_rays", "shadow_rays", "depth_sum"):
        assert abs(st[key] - k[key]) <= 2e-3 * k[key] + 2, (key, st[key], k[key])
    assert float(np.sqrt(np.mean(((gf[:, :3] - of[:, :3]) / spp) ** 2))) < 5e-3
    np.testing.assert_allclose(gf[:, :3].mean(0), of[:, :3].mean(0), rtol=2e-3)
    # the lights do something: the film is brighter than without them
    pair0 = common.ScenePair(gpu_ctx, scenes.cornell_box(), materials=scenes.cornell_materials)
    film.clear()
    pair0.gpu.render(film, api.make_config(w, h, r2c, c2w, spp_begin=0, spp_end=spp, trace_mode=3, **kw))
    assert gf[:, :3].mean() > 1.2 * film.download()[:, :3].mean()
    film.close(); pair.close(); pair0.close()


@pytest.mark.gpu
def test_gpu_point_light_closed_form_and_limits(gpu_ctx):
    rho, I = 0.5, 3.0e5
    P = (-100.0, 60.0, 300.0)
    ms = api.MeshSet(_plane()); oc = api.Octtree_Model(ms)
    sc = api.Scene(gpu_ctx); mm = _lit_plane(sc, rho, [(0, P, I)]); sc.set_model(oc, mesh_materials=mm); sc.commit()
    w, h = 64, 36
    r2c, c2w = common.camera_1080p_like(w, h)
    pix = np.arange(0, w * h, 5, dtype=np.int32)
    s = sc.eval_samples(api.make_config(w, h, r2c, c2w, mode=1, xs=1, ys=1, jitter=0, max_depth=1, trace_mode=3), pix, np.zeros_like(pix))
    ray = s["ray"].astype(np.float64)
    x = ray[:, :3] + ((600.0 - ray[:, 2]) / ray[:, 5])[:, None] * ray[:, 3:]
    d = np.asarray(P) - x; r2 = (d * d).sum(1)
    want = rho / np.pi * I * (-d[:, 2] / np.sqrt(r2)) / r2
    assert np.allclose(s["L"], want[:, None], rtol=3e-4)
    # more next-event slots than the wavefront keeps queues for is an explicit error, not a silent truncation
    from computational_ray_tracer_b200 import _capi
    white = sc.add_spectrum(0, c=1.0)
    for i in range(17):
        sc.add_light(0, (float(i), 0.0, 100.0), white, 1.0)
    sc.commit()
    film = api.Film(gpu_ctx, w, h)
    with pytest.raises(_capi.CrtError, match="next-event slots"):
        sc.render(film, api.make_config(w, h, r2c, c2w, mode=1, max_depth=1))
    film.close(); sc.close(); oc.close()
