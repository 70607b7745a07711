"""Definitions of the golden cases: inputs are regenerated deterministically, expected outputs live in tests/golden/.
Shared by tools/make_golden.py (writes them from the oracle) and tests/test_golden.py (checks oracle and CUDA)."""
import numpy as np

from computational_ray_tracer_b200 import api, scenes

import common


def kat_inputs():
    return dict(pcg=[(0, 0, 0, 0), (1, 7, 0, 0), (1, 99, 0, 65536 * 5 + 2), (2, 54, 42, 0)],
                samplers=[(0, 4, 4, 0, 0, 3, 7, 2, 0), (1, 4, 4, 1, 0, 100, 200, 15, 0), (1, 8, 8, 1, 0, 1919, 1080, 63, 0), (1, 2, 3, 0, 5, 9, 9, 4, 2)],
                hashes=[(0, 0, 0), (1, 1, 0), (640, 360, 0), (1919, 1080, 3)])


def case_kat(O):
    L = O.lib()
    k = kat_inputs()
    pcg = np.zeros((len(k["pcg"]), 8), np.uint32)
    for i, (mode, seq, off, adv) in enumerate(k["pcg"]):
        L.orc_pcg32(mode, seq, off, adv, 8, O.up(pcg[i]), None)
    smp = np.zeros((len(k["samplers"]), 7), np.float32)
    for i, a in enumerate(k["samplers"]):
        L.orc_sampler_sequence(*a, b"12221", O.fp(smp[i]))
    hs = np.array([L.orc_hash_pixel_seed(*h) for h in k["hashes"]], np.uint64)
    lam = np.zeros((4, 16), np.float32)
    for i, u in enumerate([0.0, 0.3, 0.71, 0.9999]):
        L.orc_sample_visible(u, O.fp(lam[i, :8]), O.fp(lam[i, 8:]))
    return dict(pcg=pcg, samplers=smp, hashes=hs, lambdas=lam)


def trace_inputs():
    meshes = scenes.random_soup(600, seed=21)
    r2c, c2w = common.camera_1080p_like(96, 54)
    rays = np.concatenate([common.pixel_center_rays(96, 54, r2c, c2w), common.random_rays(3000, 8)])
    return meshes, rays


def case_trace_soup(O):
    meshes, rays = trace_inputs()
    sc = O.OracleScene(); sc.set_model(meshes); sc.build_octree()
    r = sc.trace(rays, 0, nthreads=4, counters=True)
    tmax = np.linspace(100, 900, len(rays)).astype(np.float32)
    occ = sc.trace(rays, 2, tmax=tmax, nthreads=4)["mesh"]
    st = sc.octree_stats()
    out = dict(mesh=r["mesh"], tri=r["tri"], t=r["t"], bary=r["bary"], occluded=occ.astype(np.int8),
               counters=np.array([r["counters"][k] for k in ("rays", "nodes", "tris", "leaves")], np.int64),
               octree=np.array([st["nodes"], st["leaves"], st["max_leaf"], st["depth"], st["refs"]], np.int64),
               vertex_checksum=np.array([np.ascontiguousarray(meshes[0]["positions"]).view(np.uint32).astype(np.uint64).sum()], np.uint64))
    sc.close()
    return out


FILM_W, FILM_H, FILM_SPP = 64, 36, 4


def film_inputs():
    meshes = scenes.heightfield(32, with_light=False)
    r2c, c2w = common.camera_1080p_like(FILM_W, FILM_H)
    kw = dict(sampler_kind=1, xs=2, ys=2, jitter=1, spp_begin=0, spp_end=FILM_SPP)
    rs = np.random.RandomState(4)
    pix = rs.randint(0, FILM_W * FILM_H, 256).astype(np.int32)
    idx = rs.randint(0, FILM_SPP, 256).astype(np.int32)
    return meshes, r2c, c2w, kw, pix, idx


def case_tier_a_film(O):
    meshes, r2c, c2w, kw, pix, idx = film_inputs()
    sc = O.OracleScene(); sc.set_model(meshes); sc.build_octree()
    p = O.make_params(FILM_W, FILM_H, r2c, c2w, nthreads=1, **kw)
    film = sc.render(p)["film"]
    s = sc.eval_samples(p, pix, idx)
    rgb8, rgbf = O.resolve(film)
    sc.close()
    return dict(film=film, rgb8=rgb8, ray=s["ray"], lam=s["lam"], pdf=s["pdf"], L=s["L"], rgb=s["rgb"], weight=s["weight"])


CASES = {"kat": case_kat, "trace_soup": case_trace_soup, "tier_a_film": case_tier_a_film}
