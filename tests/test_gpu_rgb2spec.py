"""RGB -> spectrum on the GPU: the table generator kernel (csrc/crt_rgb2spec.cuh) and rendering with non-grey RGB.

* generator: every cell of the 3 x 64^3 table, looked up at its own grid point and pushed through the renderer's spectrum -> XYZ
  -> RGB path, reproduces the cell's RGB; random colours (trilinear interpolation between cells) round-trip within a stated error;
  grid cells agree with the host's single-cell solver;
* Tier A (`colors`, RayTracerTestApp.h:208,254) and Tier B (RGBAlbedoSpectrum reflectances) with coloured RGB against the oracle
  fed the SAME table -- and, where it travelled, against the reference's own compiled code fed the same table."""
import numpy as np
import pytest

import common
import oracle_lib as O
import ref_lib as R
from common import ScenePair, bits
from computational_ray_tracer_b200 import _capi, api, scenes
from test_cpu_rgb2spec import _roundtrip_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def table(gpu_ctx):
    scale, data, ms = gpu_ctx.generate_rgb2spec()
    print(f"rgb2spec table generated on the GPU in {ms:.1f} ms")
    yield scale, data, ms
    O.lib().orc_set_rgb_table(None, None)


def test_generated_table_round_trips(crt_lib, table):
    scale, data, ms = table
    assert np.isfinite(data).all() and ms < 5000
    k = np.arange(64) / 63.0
    sm = lambda x: x * x * (3 - 2 * x)
    np.testing.assert_allclose(scale, sm(sm(k)), rtol=1e-6)
    rgb_of = _roundtrip_model(crt_lib)
    rs = np.random.RandomState(1)
    # (1) grid cells: stored coefficients reproduce the cell colour
    errs = []
    for _ in range(400):
        l, kk, j, i = rs.randint(0, 3), rs.randint(1, 64), rs.randint(0, 64), rs.randint(0, 64)
        b = float(scale[kk]); rgb = np.zeros(3); rgb[l] = b; rgb[(l + 1) % 3] = b * i / 63.0; rgb[(l + 2) % 3] = b * j / 63.0
        errs.append(np.abs(rgb_of(data[l, kk, j, i]) - rgb).max())
    errs = np.array(errs)
    assert np.median(errs) < 1e-5 and (errs < 1e-3).mean() > 0.97 and errs.max() < 2e-2, (np.median(errs), errs.max())
    # (2) random colours through the trilinear lookup
    cols = rs.uniform(0.02, 0.98, (600, 3)).astype(np.float32)
    e2 = np.array([np.abs(rgb_of(api.rgb2spec_lookup(scale, data, c)) - c).max() for c in cols])
    assert np.median(e2) < 1e-3 and e2.max() < 2e-2, (np.median(e2), e2.max())
    # (3) the host's single-cell solver lands on the same spectra for moderately saturated cells
    lam = np.arange(360, 831, dtype=np.float64)
    for l, kk, j, i in [(0, 40, 20, 30), (1, 50, 40, 10), (2, 30, 32, 32), (0, 60, 50, 50)]:
        b = float(scale[kk]); rgb = np.zeros(3, np.float32); rgb[l] = b; rgb[(l + 1) % 3] = b * i / 63.0; rgb[(l + 2) % 3] = b * j / 63.0
        ch = api.rgb2spec_fit(rgb).astype(np.float64); cg = data[l, kk, j, i].astype(np.float64)
        sig = lambda c: 0.5 + (c[0] * lam * lam + c[1] * lam + c[2]) / (2 * np.sqrt(1 + (c[0] * lam * lam + c[1] * lam + c[2]) ** 2))
        assert np.abs(sig(ch) - sig(cg)).max() < 2e-3, (l, kk, j, i)


def test_non_grey_is_refused_without_a_table(crt_lib):
    ctx = api.Context(0)
    sc = api.Scene(ctx)
    with pytest.raises(_capi.CrtError, match="spectrum table"):
        sc.add_spectrum(7, interleaved=np.float32([0.2, 0.5, 0.8]))
    assert sc.add_spectrum(7, interleaved=np.float32([0.4, 0.4, 0.4])) >= 0           # grey needs none
    sc.close(); ctx.close()


def test_tier_a_coloured_albedo(gpu_ctx, table):
    scale, data, _ = table
    O.lib().orc_set_rgb_table(O.fp(scale), O.fp(data))
    meshes = scenes.heightfield(96, with_light=False)
    pair = ScenePair(gpu_ctx, meshes)
    w, h, spp = 160, 90, 4
    r2c, c2w = common.camera_1080p_like(w, h)
    albedo = (0.8, 0.3, 0.1)
    kw = dict(sampler_kind=1, xs=2, ys=2, jitter=1, spp_begin=0, spp_end=spp, albedo=albedo)
    film = api.Film(gpu_ctx, w, h)
    pair.gpu.render(film, api.make_config(w, h, r2c, c2w, **kw))
    gf = film.download()
    of = pair.orc.render(O.make_params(w, h, r2c, c2w, nthreads=8, **kw))["film"]
    assert np.array_equal(gf[:, 3], of[:, 3])
    assert float(np.sqrt(np.mean((gf[:, :3] - of[:, :3]) ** 2))) < 2e-5 * spp
    g8, _ = film.resolve()
    m = g8.reshape(-1, 3).astype(float).mean(0)
    assert m[0] > m[1] > m[2], m                   # an orange surface renders orange
    if R.available():
        L = R.lib(); L.ref_set_rgb_table(R.fp(scale), R.fp(data))
        r = R.RefScene(); r.set_model(meshes); r.build_octree()
        rf = r.render_tier_a(R.make_params(w, h, sampler_kind=1, xs=2, ys=2, jitter=1, spp_begin=0, spp_end=spp, albedo=albedo, nthreads=8))
        assert np.array_equal(bits(of), bits(rf)), "oracle and compiled reference must agree bit for bit on a coloured film"
        assert float(np.sqrt(np.mean((gf[:, :3] - rf[:, :3]) ** 2))) < 2e-5 * spp
        r.close()
    film.close(); pair.close()


def test_tier_b_rgb_reflectances(gpu_ctx, table):
    """Cornell box whose walls are RGBAlbedoSpectrum colours and whose light is an RGBIlluminantSpectrum."""
    scale, data, _ = table
    O.lib().orc_set_rgb_table(O.fp(scale), O.fp(data))

    def mats(sc):
        rgb = lambda c: np.float32(c)
        w_ = sc.add_spectrum(7, interleaved=rgb([0.73, 0.73, 0.70])); r_ = sc.add_spectrum(7, interleaved=rgb([0.65, 0.06, 0.05]))
        g_ = sc.add_spectrum(7, interleaved=rgb([0.12, 0.45, 0.15])); li = sc.add_spectrum(8, interleaved=rgb([1.0, 0.85, 0.6]))
        un = sc.add_spectrum(9, interleaved=rgb([0.2, 0.3, 0.9]))
        mw = sc.add_material(type=0, refl=w_); mr = sc.add_material(type=0, refl=r_); mg = sc.add_material(type=0, refl=g_)
        ml = sc.add_material(type=0, refl=-1, emit=li, emit_scale=12.0)
        mu = sc.add_material(type=0, refl=un)
        sc.add_shape(0, scenes.translation(-110, -170, 720), [80.0, -80.0, 80.0, 360.0], material=mu)
        return [mw, mr, mg, ml]
    pair = ScenePair(gpu_ctx, scenes.cornell_box(), materials=mats)
    w, h = 96, 96
    r2c, c2w = common.camera_1080p_like(w, h)
    kw = dict(mode=1, sampler_kind=1, xs=4, ys=4, jitter=1, max_depth=5)
    rs = np.random.RandomState(3)
    pix = rs.randint(0, w * h, 4000).astype(np.int32); idx = rs.randint(0, 16, 4000).astype(np.int32)
    g = pair.gpu.eval_samples(api.make_config(w, h, r2c, c2w, **kw), pix, idx)
    o = pair.orc.eval_samples(O.make_params(w, h, r2c, c2w, **kw), pix, idx)
    assert np.array_equal(bits(g["ray"]), bits(o["ray"]))
    ok = np.isclose(g["L"], o["L"], rtol=2e-4, atol=1e-5).all(axis=1)
    assert ok.mean() >= 0.98, ok.mean()
    assert (o["L"] > 0).any()
    pair.close()
