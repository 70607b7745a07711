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


@pytest.mark.parametrize("rho,depth", [(0.5, 1), (0.5, 4), (0.8, 3)])
def test_white_furnace(oracle, rho, depth):
    le = 0.25
    sc = O.OracleScene(); sc.set_model(_closed_box()); sc.build_octree()
    refl = sc.add_spectrum(0, c=rho); emit = sc.add_spectrum(0, c=1.0)
    m = sc.add_material(type=0, refl=refl, emit=emit, emit_scale=le)
    sc.set_mesh_materials([m] * 6)
    w = h = 24
    r2c, c2w = O.camera_matrices(0, 1.0, 1000.0, 0.0, 0.0, 45.0, (3, -2, 5), (0.2, 0.1, 1), (1, 0, 0), (0, 1, 0), w, h)
    p = O.make_params(w, h, r2c, c2w, mode=1, xs=8, ys=8, jitter=1, max_depth=depth, rr_depth=0, ray_eps=1e-2, shadow_eps=1e-3)
    pid = np.repeat(np.arange(w * h, dtype=np.int32), 64); idx = np.tile(np.arange(64, dtype=np.int32), w * h)
    L = sc.eval_samples(p, pid, idx)["L"]
    want = le * sum(rho ** k for k in range(depth + 1))
    assert L.shape == (w * h * 64, 8) and (L >= le * 0.999).all()           # every path sees the emitting wall it hits first
    got = float(L.astype(np.float64).mean())
    assert abs(got - want) / want < 0.01, (got, want)
    # all eight wavelengths carry the same (constant-spectrum) radiance
    assert np.abs(L.mean(0) / got - 1).max() < 1e-5
    sc.close()
