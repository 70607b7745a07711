#!/usr/bin/env python
"""Render a Wavefront OBJ (or, with no argument, a procedural height field) the way the reference's RayTracerTestApp renders
its dragon: one TriModel in an Octtree_Model, PerspectiveCamera at the origin looking down +z, the reference's own Li
(`--mode 0`) or the path integrator with an overhead emissive quad (`--mode 1`).  Writes a binary PPM.

    python examples/render_obj.py model.obj --out model.ppm --spp 16 --distance 3
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from computational_ray_tracer_b200 import api, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj", nargs="?")
    ap.add_argument("--out", default="render.ppm")
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=640)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--distance", type=float, default=2.5, help="camera distance in units of the model's bounding radius")
    a = ap.parse_args()
    if a.obj:
        meshes = api.load_obj(a.obj)
        allp = np.concatenate([m["positions"] for m in meshes])
        c = 0.5 * (allp.min(0) + allp.max(0)); r = float(np.linalg.norm(allp.max(0) - allp.min(0))) / 2
        for m in meshes:                        # centre the model in front of the camera (world-space positions, like the reference's precomputed path)
            m["positions"] = ((m["positions"] - c) * np.float32([1, 1, -1]) + np.float32([0, 0, a.distance * r])).astype(np.float32)
            m["normals"] = (m["normals"] * np.float32([1, 1, -1])).astype(np.float32)
            m["indices"] = m["indices"][:, ::-1].copy()
    else:
        meshes = scenes.heightfield(256, with_light=False)
        r = 400.0
    ctx = api.Context(0)
    ms = api.MeshSet(meshes); oc = api.Octtree_Model(ms)
    sc = api.Scene(ctx)
    mats = None
    if a.mode == 1:
        z = a.distance * r if a.obj else 650.0
        meshes.append(scenes.quad_mesh((-r, 1.5 * r, z - r), (r, 1.5 * r, z - r), (-r, 1.5 * r, z + r), (r, 1.5 * r, z + r), (0, -1, 0)))
        ms = api.MeshSet(meshes); oc = api.Octtree_Model(ms)
        grey = sc.add_spectrum(0, c=0.6); d65 = sc.add_spectrum(4, n=2)
        surf = sc.add_material(type=0, refl=grey); light = sc.add_material(type=0, refl=-1, emit=d65, emit_scale=6.0, two_sided=1)
        mats = [surf] * (len(meshes) - 1) + [light]
    sc.set_model(oc, mesh_materials=mats)
    sc.commit()
    print("octree:", oc.stats())
    film = api.Film(ctx, a.width, a.height)
    r2c, c2w = api.camera_matrices(0, 1.0, 1000.0, 45.0, (0, 0, 0), (0, 0, 1), (0, 1, 0), a.width, a.height)
    n = int(np.ceil(np.sqrt(a.spp)))
    st = sc.render(film, api.make_config(a.width, a.height, r2c, c2w, mode=a.mode, xs=n, ys=n, spp_begin=0, spp_end=a.spp, trace_mode=api.DEFAULT_TRACE_MODE))
    rgb8, _ = film.resolve(want_float=False)
    with open(a.out, "wb") as f:
        f.write(f"P6 {a.width} {a.height} 255\n".encode())
        f.write(rgb8.reshape(a.height, a.width, 3)[::-1].tobytes())         # film row 0 is the top of the reference's GL texture
    print(f"{st['paths']} paths in {st['total_ms']:.1f} ms -> {a.out}")


if __name__ == "__main__":
    main()
