"""Committed golden fixtures (tests/golden/, written by tools/make_golden.py from the oracle).
CPU: the oracle still reproduces them bit for bit (guards the checker itself against drift).
GPU: the CUDA path reproduces them -- ids, t and barycentrics exactly; radiance within the stated tolerance."""
import os

import numpy as np
import pytest

import golden_cases as G
import oracle_lib as O
from common import bits

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_reproduces_golden(oracle, name):
    want = _load(name)
    got = G.CASES[name](oracle)
    assert sorted(got) == sorted(want)
    for k in want:
        a, b = np.asarray(got[k]), want[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        assert np.array_equal(bits(a), bits(b)), k


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_kat_matches_golden(crt_lib, gpu_ctx):
    import ctypes as C
    from computational_ray_tracer_b200._capi import f32p, u32p
    want = _load("kat")
    k = G.kat_inputs()
    for i, (mode, seq, off, adv) in enumerate(k["pcg"]):
        out = np.zeros(8, np.uint32)
        assert crt_lib.crt_kat_pcg32(mode, seq, off, adv, 8, 1, out.ctypes.data_as(u32p), None) == 0
        assert np.array_equal(out, want["pcg"][i])
    for i, a in enumerate(k["samplers"]):
        out = np.zeros(7, np.float32)
        assert crt_lib.crt_kat_sampler(*a, b"12221", 1, out.ctypes.data_as(f32p)) == 0
        assert np.array_equal(bits(out), bits(want["samplers"][i]))
    for i, (x, y, s) in enumerate(k["hashes"]):
        out = C.c_uint64()
        key = np.array([x, y, s], np.int32).tobytes()
        assert crt_lib.crt_kat_hash(key, 12, 0, 1, C.byref(out)) == 0
        assert out.value == int(want["hashes"][i])


@pytest.mark.gpu
def test_gpu_trace_matches_golden(gpu_ctx):
    from computational_ray_tracer_b200 import api
    want = _load("trace_soup")
    meshes, rays = G.trace_inputs()
    ms = api.MeshSet(meshes); oc = api.Octtree_Model(ms)
    st = oc.stats()
    assert [st["nodes"], st["leaves"], st["max_leaf"], st["depth"], st["refs"]] == want["octree"].tolist()
    sc = api.Scene(gpu_ctx); sc.set_model(oc); sc.commit()
    for mode in (0, 3):
        g = sc.trace_closest(rays, mode=mode)
        assert np.array_equal(g["mesh"], want["mesh"]) and np.array_equal(g["tri"], want["tri"])
        hit = want["tri"] >= 0
        assert np.array_equal(bits(g["t"][hit]), bits(want["t"][hit]))
        assert np.array_equal(bits(g["bary"][hit]), bits(want["bary"][hit]))
    tmax = np.linspace(100, 900, len(rays)).astype(np.float32)
    for mode in (0, 3):
        assert np.array_equal(sc.trace_any(rays, tmax, mode=mode).astype(np.int8), want["occluded"])
    sc.close(); oc.close()


@pytest.mark.gpu
def test_gpu_tier_a_film_matches_golden(gpu_ctx):
    from computational_ray_tracer_b200 import api
    want = _load("tier_a_film")
    meshes, r2c, c2w, kw, pix, idx = G.film_inputs()
    ms = api.MeshSet(meshes); oc = api.Octtree_Model(ms)
    sc = api.Scene(gpu_ctx); sc.set_model(oc); sc.commit()
    cfg = api.make_config(G.FILM_W, G.FILM_H, r2c, c2w, **kw)
    s = sc.eval_samples(cfg, pix, idx)
    assert np.array_equal(bits(s["ray"]), bits(want["ray"]))              # IEEE-only arithmetic: exact
    assert np.array_equal(bits(s["weight"]), bits(want["weight"]))
    # wavelengths come from atanh/cosh (libm vs libdevice): 1e-5 relative; radiance follows through 1 nm table bins
    np.testing.assert_allclose(s["lam"], want["lam"], rtol=1e-5)
    np.testing.assert_allclose(s["pdf"], want["pdf"], rtol=1e-4)
    close = np.isclose(s["L"], want["L"], rtol=1e-4, atol=1e-6)
    assert close.mean() > 0.995                                             # a bin-edge crossing moves one lambda's table value
    film = api.Film(gpu_ctx, G.FILM_W, G.FILM_H)
    sc.render(film, cfg)
    gf = film.download()
    assert np.array_equal(gf[:, 3], want["film"][:, 3])
    rmse = float(np.sqrt(np.mean((gf[:, :3] - want["film"][:, :3]) ** 2)))
    assert rmse < 1e-4, rmse
    g8, _ = film.resolve()
    assert np.abs(g8.astype(int) - want["rgb8"].astype(int)).max() <= 1
    film.close(); sc.close(); oc.close()
