#!/usr/bin/env python
"""Render BASELINE.json's five configs on one GPU, time them, and check each against the oracle on a pixel subsample.

  python tools/run_configs.py [--configs C1,C2,C3,C4,C5] [--cpu-seconds 8] [--out gpurun_out/configs.json]

Per config: GPU Mpaths/s and Mrays/s (CUDA events inside crt_render, scene resident), the oracle's rate on the host cores
(bounded subsample), RMSE of the per-pixel mean sensor RGB between GPU film and oracle on the subsampled pixels, and the
number of closest-hit mismatches between the exact BFS kernel and the ordered traversal on the primary rays."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import common  # noqa: E402
import oracle_lib as O  # noqa: E402
from computational_ray_tracer_b200 import api, scenes  # noqa: E402


STRIDE = {}          # config -> fixed pixel stride of the oracle subsample (--stride)


def cfg_table():
    return {
        "C1": dict(name="Cornell box + 2 spheres, 256x256 @16 spp, depth<=5, NEE", meshes=scenes.cornell_box, materials=scenes.cornell_materials,
                   w=256, h=256, spp=16, xs=4, ys=4, max_depth=5, rr_depth=0, full_oracle=True),
        "C2": dict(name="1.0M-tri height field + emissive quad, 1080p @64 spp, diffuse + NEE", meshes=lambda: scenes.heightfield(708), materials=scenes.c2_materials,
                   w=1920, h=1080, spp=64, xs=8, ys=8, max_depth=5, rr_depth=0),
        "C3": dict(name="8x8 dielectric/conductor sphere lattice (dispersion), 1080p @256 spp, depth<=16, RR after 3", meshes=scenes.spheres_lattice_meshes,
                   materials=scenes.spheres_lattice_materials, w=1920, h=1080, spp=256, xs=16, ys=16, max_depth=16, rr_depth=3),
        "C4": dict(name="250k-tri height field + 1000 emissive triangles, light-sampled NEE, 1080p @64 spp", meshes=scenes.many_light_scene,
                   materials=scenes.many_light_materials, w=1920, h=1080, spp=64, xs=8, ys=8, max_depth=5, rr_depth=0),
        "C5": dict(name="10.0M-tri height field, 4K, 16 of 1024 spp on one GPU", meshes=lambda: scenes.heightfield(2237, seed=5), materials=scenes.c2_materials,
                   w=3840, h=2160, spp=16, xs=32, ys=32, max_depth=5, rr_depth=0, no_oracle=True),
    }


def run(key, c, ctx, cpu_seconds):
    out = dict(config=key, name=c["name"])
    meshes = c["meshes"]()
    ms = api.MeshSet(meshes)
    t0 = time.time()
    oc = api.Octtree_Model(ms, algorithm=api.BUILD_GPU, ctx=ctx)
    out["triangles"] = ms.n_triangles
    out["gpu_octree_build_s"] = round(time.time() - t0, 3)
    out["octree"] = oc.stats()
    sc = api.Scene(ctx)
    mm = c["materials"](sc)
    sc.set_model(oc, mesh_materials=mm)
    sc.commit()
    out["scene_device_MB"] = round(sc.device_bytes() / 1e6, 1)
    w, h, spp = c["w"], c["h"], c["spp"]
    r2c, c2w = common.camera_1080p_like(w, h)
    kw = dict(mode=1, xs=c["xs"], ys=c["ys"], jitter=1, max_depth=c["max_depth"], rr_depth=c["rr_depth"], spp_begin=0, spp_end=spp)
    film = api.Film(ctx, w, h)
    sc.render(film, api.make_config(w, h, r2c, c2w, trace_mode=api.DEFAULT_TRACE_MODE, **dict(kw, spp_end=min(spp, 2))))      # warm-up
    film.clear()
    st = sc.render(film, api.make_config(w, h, r2c, c2w, trace_mode=api.DEFAULT_TRACE_MODE, time_kernels=1, **kw))
    gf = film.download()
    secs = st["total_ms"] / 1e3
    out.update(gpu_mpaths_s=st["paths"] / secs / 1e6, gpu_mrays_s=(st["closest_rays"] + st["shadow_rays"]) / secs / 1e6, gpu_ms=st["total_ms"],
               rays_per_path=(st["closest_rays"] + st["shadow_rays"]) / st["paths"], mean_depth=st["depth_sum"] / st["paths"],
               traversal_share=st["trace_ms"] / st["total_ms"], exact_retraced_rays=st["exact_retraced_rays"], kernel_launches=st["kernel_launches"])
    # exact BFS kernel vs ordered traversal on the primary rays (every 3rd pixel)
    rays = common.pixel_center_rays(w, h, r2c, c2w, step=3)
    a = sc.trace_closest(rays, mode=0); b = sc.trace_closest(rays, mode=api.DEFAULT_TRACE_MODE)
    # whole-path check (bounce and shadow rays included): 2 sample indices rendered with the exact BFS kernel and with the
    # production traversal must give bit-identical films
    f0 = api.Film(ctx, w, h); f1 = api.Film(ctx, w, h)
    s0 = sc.render(f0, api.make_config(w, h, r2c, c2w, trace_mode=0, **dict(kw, spp_end=2)))
    s1 = sc.render(f1, api.make_config(w, h, r2c, c2w, trace_mode=api.DEFAULT_TRACE_MODE, **dict(kw, spp_end=2)))
    out["films_bit_identical_across_trace_modes"] = bool(np.array_equal(f0.download().view(np.uint32), f1.download().view(np.uint32)))
    out["rays_in_that_check"] = int(s0["closest_rays"] + s0["shadow_rays"])
    out["bfs_kernel_mrays_s"] = (s0["closest_rays"] + s0["shadow_rays"]) / (s0["total_ms"] / 1e3) / 1e6
    f0.close(); f1.close()
    out["primary_rays_checked"] = len(rays)
    out["ordered_vs_bfs_mismatches"] = int((a["tri"] != b["tri"]).sum() + (a["mesh"] != b["mesh"]).sum() + (a["t"].view(np.uint32) != b["t"].view(np.uint32)).sum())
    if not c.get("no_oracle"):
        orc = O.OracleScene(); orc.set_model(meshes); orc.build_octree()
        orc.set_mesh_materials(c["materials"](orc))
        nthreads = os.cpu_count() or 1
        okw = dict(kw); okw.pop("spp_end")
        if c.get("full_oracle"):
            stride, ospp = 1, spp
        else:
            # calibrate: thin sample, then size for ~cpu_seconds
            p = O.make_params(w, h, r2c, c2w, nthreads=nthreads, pixel_stride=2039, **dict(okw, spp_end=1))
            r = orc.render(p, counters=True)
            rate = max(r["counters"]["paths"], 1) / max(r["seconds"], 1e-6)
            ospp = spp
            stride = int(max(1, round(w * h * ospp / (rate * cpu_seconds))))
            while stride > 1 and (w % stride == 0 or stride % 2 == 0):
                stride += 1
            stride = STRIDE.get(key, stride)
        p = O.make_params(w, h, r2c, c2w, nthreads=nthreads, pixel_stride=stride, **dict(okw, spp_end=ospp))
        r = orc.render(p, counters=True)
        of, cn = r["film"], r["counters"]
        sel = of[:, 3] > 0
        out.update(cpu_cores=nthreads, cpu_mpaths_s=cn["paths"] / r["seconds"] / 1e6, cpu_mrays_s=(cn["closest_rays"] + cn["shadow_rays"]) / r["seconds"] / 1e6,
                   cpu_sample=f"every {stride}th pixel x {ospp} spp = {cn['paths']} paths in {r['seconds']:.1f} s", oracle_pixels=int(sel.sum()))
        assert np.array_equal(gf[sel, 3], of[sel, 3])
        mg, mo = gf[sel, :3] / gf[sel, 3:4], of[sel, :3] / of[sel, 3:4]
        out["rmse_mean_rgb"] = float(np.sqrt(np.mean((mg - mo) ** 2)))
        out["mean_rgb_gpu"] = [float(x) for x in mg.mean(0)]
        out["mean_rgb_oracle"] = [float(x) for x in mo.mean(0)]
        out["speedup_vs_cpu"] = out["gpu_mpaths_s"] / out["cpu_mpaths_s"]
        orc.close()
    film.close(); sc.close(); oc.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,C3,C4")
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--stride", default="", help="fixed pixel strides of the oracle subsample, e.g. C2=165,C4=99 (default: sized for --cpu-seconds)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    a = ap.parse_args()
    for item in filter(None, a.stride.split(",")):
        STRIDE[item.split("=")[0]] = int(item.split("=")[1])
    O.build()
    ctx = api.Context(0)
    table = cfg_table()
    res = []
    for key in a.configs.split(","):
        r = run(key, table[key], ctx, a.cpu_seconds)
        print(json.dumps(r), flush=True)
        res.append(r)
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
