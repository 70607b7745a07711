#!/usr/bin/env python
"""Build tuning variants of libcrt_b200.so (compile-time constants) into variants/<name>.so, in parallel; the in-tree build is untouched.
usage: tools/build_variants.py "name1:-DCRT_WIDE_LEAF_WAIT=8" "name2:-DCRT_SUBPACKET=4 -DCRT_WIDE_STACK=12" ...
run one with:  CRT_B200_LIB=$PWD/variants/name1.so python bench.py --no-cpu-baseline"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from computational_ray_tracer_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, "variants")


def one(spec):
    name, defs = spec.split(":", 1)
    objdir = os.path.join(OUT, "obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src in B.SOURCES:
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [B.NVCC] + [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + defs.split() + ["-x", "cu", "-c", os.path.join(B.CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            return name, "FAILED " + src + "\n" + r.stderr[-2000:]
        objs.append(obj)
    lib = os.path.join(OUT, name + ".so")
    r = subprocess.run([B.NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"], capture_output=True, text=True)
    return name, "ok" if r.returncode == 0 else "LINK FAILED\n" + r.stderr[-2000:]


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    with ThreadPoolExecutor(max_workers=int(os.environ.get("JOBS", "6"))) as ex:
        for name, status in ex.map(one, sys.argv[1:]):
            print(name, status, flush=True)
