"""Build libcrt_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libcrt_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# --fmad=false: no implicit FMA contraction -- hit ids and radiance must reproduce the reference's fp32
# operation sequence bit for bit; fmaf() is still emitted where the reference calls std::fma.
# Division and sqrt stay IEEE (nvcc defaults --prec-div=true --prec-sqrt=true, --ftz=false).
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-mfma,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
]
if os.environ.get("CRT_NVCC_DEFINES"):          # experiments, e.g. CRT_NVCC_DEFINES="-DCRT_MR_MINBLOCKS=4"
    NVCC_FLAGS += os.environ["CRT_NVCC_DEFINES"].split()
SOURCES = ["crt_host.cpp", "crt_spectra.cpp", "crt_obj.cpp", "crt_capi.cu"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    deps += [os.path.join(PKG, "data", "spectral_tables.inc"), os.path.join(PKG, "..", "include", "crt_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    log = []
    for src in SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        cmd = [NVCC] + NVCC_FLAGS + ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on " + src)
        objs.append(obj)
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(CSRC, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
