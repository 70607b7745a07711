"""White-furnace check of the Tier-B DEFINITION (the oracle's LiPath, which the CUDA path follows expression by expression and which has
no reference implementation to be pinned to): inside a closed box whose walls all emit L_e and reflect a Lambertian rho, the radiance
along any ray is the truncated geometric series L_e (1 + rho + ... + rho^D), D = max_depth -- emission is counted at the camera vertex,
every later term comes from next-event estimation, and an unbiased NEE estimator has to reproduce each of them in the mean."""
import numpy as np
import pytest

import oracle_lib as O
from computational_ray_tracer_b200 import scenes


def _closed_box(h=100.0):
    f = []
    # (p00, p10, p01, p11, inward normal) for the six faces of [-h, h]^3
    f.append(scenes.quad_mesh((-h, -h, h), (h, -h, h), (-h, h, h), (h, h, h), (0, 0, -1)))
    f.append(scenes.quad_mesh((-h, -h, -h), (h, -h, -h), (-h, h, -h), (h, h, -h), (0, 0, 1)))
    f.append(scenes.quad_mesh((-h, -h, -h), (-h, -h, h), (-h, h, -h), (-h, h, h), (1, 0, 0)))
    f.append(scenes.quad_mesh((h, -h, -h), (h, -h, h), (h, h, -h), (h, h, h), (-1, 0, 0)))
    f.append(scenes.quad_mesh((-h, -h, -h), (h, -h, -h), (-h, -h, h), (h, -h, h), (0, 1, 0)))
    f.append(scenes.quad_mesh((-h, h, -h), (h, h, -h), (-h, h, h), (h, h, h), (0, -1, 0)))
    return f


@pytest.mark.parametrize("rho,depth,rr", [(0.5, 1, 0), (0.5, 4, 0), (0.8, 3, 0), (0.5, 6, 1), (0.8, 8, 2)])
def test_white_furnace(oracle, rho, depth, rr):
    le = 0.25
    sc = O.OracleScene(); sc.set_model(_closed_box()); sc.build_octree()
    refl = sc.add_spectrum(0, c=rho); emit = sc.add_spectrum(0, c=1.0)
    m = sc.add_material(type=0, refl=refl, emit=emit, emit_scale=le)
    sc.set_mesh_materials([m] * 6)
    w = h = 24
    r2c, c2w = O.camera_matrices(0, 1.0, 1000.0, 0.0, 0.0, 45.0, (3, -2, 5), (0.2, 0.1, 1), (1, 0, 0), (0, 1, 0), w, h)
    p = O.make_params(w, h, r2c, c2w, mode=1, xs=8, ys=8, jitter=1, max_depth=depth, rr_depth=rr, ray_eps=1e-2, shadow_eps=1e-3)      # rr > 0: Russian roulette from that depth on must not change the mean
    pid = np.repeat(np.arange(w * h, dtype=np.int32), 64); idx = np.tile(np.arange(64, dtype=np.int32), w * h)
    L = sc.eval_samples(p, pid, idx)["L"]
    want = le * sum(rho ** k for k in range(depth + 1))
    assert L.shape == (w * h * 64, 8) and (L >= le * 0.999).all()           # every path sees the emitting wall it hits first
    got = float(L.astype(np.float64).mean())
    assert abs(got - want) / want < (0.01 if rr == 0 else 0.02), (got, want)
    # all eight wavelengths carry the same (constant-spectrum) radiance
    assert np.abs(L.mean(0) / got - 1).max() < 1e-5
    sc.close()


@pytest.mark.parametrize("glass,dispersive", [("glass_bk7", True), ("glass_sf11", True), ("glass_bk7", False)])
def test_glass_is_invisible_in_a_furnace(oracle, glass, dispersive):
    """A non-absorbing dielectric inside a uniformly emitting enclosure cannot be seen: every path that refracts in picks up 1/eta^2, loses
    it again on the way out, and splits between reflection and refraction with probabilities that sum to one -- so the radiance along every
    camera ray is still L_e, for the hero wavelength and (same path, same factors) for its companions.  Checks the smooth-dielectric branch
    of LiPath (Fresnel choice, Snell direction, radiance scaling, TerminateSecondary bookkeeping) without any reference to compare with."""
    le = 0.5
    sc = O.OracleScene(); sc.set_model(_closed_box()); sc.build_octree()
    emit = sc.add_spectrum(0, c=1.0)
    wall = sc.add_material(type=0, refl=-1, emit=emit, emit_scale=le, two_sided=1)
    eta = sc.add_spectrum(2, name=glass)
    g = sc.add_material(type=1, eta=eta, eta_constant=0 if dispersive else 1)
    sc.set_mesh_materials([wall] * 6)
    sc.add_shape(0, scenes.translation(2, -3, 65), [7.0, -7.0, 7.0, 360.0], material=g)
    w = h = 32
    r2c, c2w = O.camera_matrices(0, 1.0, 1000.0, 0.0, 0.0, 45.0, (0, 0, 0), (0, 0, 1), (1, 0, 0), (0, 1, 0), w, h)
    p = O.make_params(w, h, r2c, c2w, mode=1, xs=4, ys=4, jitter=1, max_depth=24, rr_depth=0)
    pid = np.repeat(np.arange(w * h, dtype=np.int32), 16); idx = np.tile(np.arange(16, dtype=np.int32), w * h)
    e = sc.eval_samples(p, pid, idx)
    L = e["L"].astype(np.float64)
    kinds = sc.scene_closest(e["ray"])["kind"]
    through = kinds == 1                                   # primary ray meets the sphere
    assert 0.2 < through.mean() < 0.9
    assert np.allclose(L[~through], le, rtol=1e-6)         # walls seen directly
    # through the glass: L_e again, up to the few paths still bouncing inside after 24 interactions (probability ~ R^23) and fp32 rounding
    assert np.allclose(L[through], le, rtol=2e-4), float(np.abs(L[through] / le - 1).max())
    if dispersive:                                         # refraction terminated the companion wavelengths: pdf[1..7] = 0, pdf[0] /= 8
        pdf = e["pdf"][through]
        assert (pdf[:, 1:] == 0).all() and (pdf[:, 0] > 0).all()
    else:
        assert (e["pdf"][through] > 0).all()
    sc.close()


@pytest.mark.parametrize("metal", ["cu", "au"])
def test_conductor_reflectance_is_the_complex_fresnel_term(oracle, metal):
    """A convex conductor in the furnace reflects the wall exactly once: L / L_e per wavelength must be the unpolarised Fresnel reflectance
    of eta(lambda) + i k(lambda) at the angle of incidence, computed here independently (numpy complex arithmetic in float64)."""
    le = 0.5
    sc = O.OracleScene(); sc.set_model(_closed_box()); sc.build_octree()
    emit = sc.add_spectrum(0, c=1.0)
    wall = sc.add_material(type=0, refl=-1, emit=emit, emit_scale=le, two_sided=1)
    eta = sc.add_spectrum(2, name=metal + "_eta"); k = sc.add_spectrum(2, name=metal + "_k")
    m = sc.add_material(type=2, eta=eta, k=k)
    sc.set_mesh_materials([wall] * 6)
    centre, radius = np.float64([2, -3, 65]), 7.0
    sc.add_shape(0, scenes.translation(*centre), [radius, -radius, radius, 360.0], material=m)
    w = h = 32
    r2c, c2w = O.camera_matrices(0, 1.0, 1000.0, 0.0, 0.0, 45.0, (0, 0, 0), (0, 0, 1), (1, 0, 0), (0, 1, 0), w, h)
    p = O.make_params(w, h, r2c, c2w, mode=1, xs=2, ys=2, jitter=1, max_depth=5, rr_depth=0)
    pid = np.repeat(np.arange(w * h, dtype=np.int32), 4); idx = np.tile(np.arange(4, dtype=np.int32), w * h)
    e = sc.eval_samples(p, pid, idx)
    hit = sc.scene_closest(e["ray"])
    on = hit["kind"] == 1
    assert on.sum() > 500
    # angle of incidence from the shading normal the scene reports (fp32 sphere intersection, pinned to the reference elsewhere): at
    # grazing incidence a float64 re-derivation of the hit point differs by 1e-4 in the normal, which the steep Fresnel curve there amplifies
    d = e["ray"][on, 3:].astype(np.float64)
    cos_i = np.clip(-(hit["ns"][on].astype(np.float64) * d).sum(1), 0, 1)
    assert cos_i.min() < 0.2 and cos_i.max() > 0.95          # from grazing to normal incidence
    lam = e["lam"][on]
    want = np.zeros_like(lam, dtype=np.float64)
    for i in range(len(lam)):
        ce = sc.spectrum_sample(eta, lam[i]).astype(np.float64) + 1j * sc.spectrum_sample(k, lam[i]).astype(np.float64)
        ci = cos_i[i]
        sin2_t = (1 - ci * ci) / (ce * ce)
        cos_t = np.sqrt(1 - sin2_t)
        r_parl = (ce * ci - cos_t) / (ce * ci + cos_t); r_perp = (ci - ce * cos_t) / (ci + ce * cos_t)
        want[i] = (np.abs(r_parl) ** 2 + np.abs(r_perp) ** 2) / 2
    got = e["L"][on].astype(np.float64) / le
    assert np.allclose(got, want, rtol=2e-4, atol=1e-5), float(np.abs(got - want).max())
    assert 0.2 < want.min() and want.max() <= 1.0 and np.ptp(want) > 0.2      # copper / gold: strongly wavelength dependent
    sc.close()
