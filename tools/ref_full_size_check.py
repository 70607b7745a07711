#!/usr/bin/env python
"""Full-size (C2: 1 002 530 triangles) parity between the reference's own compiled code (oracle/_ref), the restated oracle and the
product's host octree builder.  Too slow for the unit suite (the reference's incremental builder needs ~25 s on 1 M triangles), so it is
a tool; its report is committed as profiles/ref_full_size_check.txt.   python tools/ref_full_size_check.py [quads=708]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import oracle_lib as O  # noqa: E402
import ref_lib as R  # noqa: E402
import ref_pin_cases as P  # noqa: E402
from computational_ray_tracer_b200 import api, scenes  # noqa: E402


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def main():
    quads = int(sys.argv[1]) if len(sys.argv) > 1 else 708
    meshes = scenes.heightfield(quads, with_light=True)
    ntri = sum(len(m["indices"]) for m in meshes)
    lines = [f"# tools/ref_full_size_check.py {quads}: height field of {quads}^2 quads + emissive quad = {ntri} triangles (bench.py's C2 scene)"]
    t = time.time(); r = R.RefScene(); r.set_model(meshes); nr = r.build_octree(); tr = time.time() - t
    t = time.time(); o = O.OracleScene(); o.set_model(meshes); no = o.build_octree(); to = time.time() - t
    t = time.time(); oc = api.Octtree_Model(api.MeshSet(meshes)); tp = time.time() - t
    lines.append(f"octree build: compiled reference {tr:.1f} s ({nr} nodes), oracle {to:.1f} s ({no} nodes), product host builder {tp:.1f} s ({oc.stats()['nodes']} nodes)")
    dr, do, dp = r.octree_dump(), o.octree_dump(), oc.dump()
    for name, d in (("oracle", do), ("product host builder", dp)):
        same = all(dr[k].shape == d[k].shape and np.array_equal(bits(dr[k]), bits(d[k])) for k in ("bounds", "leaf", "child", "list_off", "pairs"))
        lines.append(f"  {name} == compiled reference, node for node (bounds bits, leaf flags, child ids, per-leaf (mesh, tri) order; {len(dr['pairs'])} references): {same}")
        assert same
    W, H = 1920, 1080
    r2c, c2w = O.camera_matrices(0, 1.0, 1000.0, 0.0, 0.0, 45.0, (0, 0, 0), (0, 0, 1), (1, 0, 0), (0, 1, 0), W, H)
    rs = np.random.RandomState(0)
    n = 30000
    pid = rs.randint(0, W * H, n); idx = rs.randint(0, 64, n)
    po = O.make_params(W, H, r2c, c2w, sampler_kind=1, xs=8, ys=8, jitter=1, mode=0)
    pr = R.make_params(W, H, sampler_kind=1, xs=8, ys=8, jitter=1)
    t = time.time(); eo = o.eval_samples(po, pid, idx); to = time.time() - t
    t = time.time(); er = r.eval_samples(pr, pid, idx); tr = time.time() - t
    same = {k: bool(np.array_equal(bits(eo[k]), bits(er[k]))) for k in eo}
    lines.append(f"Tier A evaluate_pixel + Li on {n} random (pixel, sample index) pairs of the 1080p / 64 spp frame: oracle {to:.1f} s, compiled reference {tr:.1f} s "
                 f"({1e6 * tr / n:.0f} us per sample on one core); bit-identical per field: {same}")
    assert all(same.values())
    so = o.traverse_surface(eo["ray"]); sr = r.traverse_surface(eo["ray"], nthreads=8)
    hit = sr["found"] > 0
    ok = np.array_equal(so["found"], sr["found"]) and all(np.array_equal(bits(so[k][hit]), bits(sr[k][hit])) for k in ("n", "hitp", "uv"))
    lines.append(f"Octtree_Model::Traverse surface records on those camera rays ({int(hit.sum())} hits): identical: {ok}")
    assert ok
    to_ = o.trace(eo["ray"], mode=0)
    s2 = r.surface_of(to_["mesh"], to_["tri"], eo["ray"])
    ok = bool((s2["found"][hit] == 1).all()) and all(np.array_equal(bits(s2[k][hit]), bits(sr[k][hit])) for k in ("n", "hitp", "uv"))
    lines.append(f"oracle hit ids re-evaluated by the reference's Triangle::BasicIntersect -> CalculateLocalSurface reproduce those records: {ok}")
    assert ok
    out = os.path.join(ROOT, "profiles", "ref_full_size_check.txt")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
