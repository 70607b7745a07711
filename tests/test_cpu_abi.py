"""The drop-in boundary: libcrt_b200.so loads, exports every symbol include/crt_b200.h declares, its structs have the
layout the ctypes mirror assumes, and without a GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from computational_ray_tracer_b200 import _capi, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "crt_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(crt_lib):
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(crt_lib, n), f"{n} declared in crt_b200.h but not exported by libcrt_b200.so"
    # and the ctypes table binds exactly the header's functions
    assert sorted(_capi.EXPORTS) == names


def test_header_compiles_as_c_and_struct_sizes_match(tmp_path, crt_lib):
    src = tmp_path / "abi.c"
    src.write_text('#include "crt_b200.h"\n#include <stdio.h>\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(crt_mesh_desc), '
                   'sizeof(crt_octree_stats), sizeof(crt_render_config), sizeof(crt_render_stats)); return 0;}\n')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(_capi.MeshDesc), C.sizeof(_capi.OctreeStats), C.sizeof(_capi.RenderConfig), C.sizeof(_capi.RenderStats)]


def test_no_torch_types_in_the_abi():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)         # declarations only, comments stripped
    assert "torch" not in src.lower() and "at::" not in src and "#include <cuda" not in src and "std::" not in src


def test_no_gpu_means_loud_failure(crt_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = crt_lib.crt_context_create(0, C.byref(h))
    assert rc != 0 and not h
    assert b"no CPU fallback" in crt_lib.crt_last_error()
    with pytest.raises(_capi.CrtError):
        api.Context(0)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "computational_ray_tracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                # citations of the oracle's source files in comments are fine; loading, linking or including it is not
                for needle in ("oracle_lib", "liboracle", "oracle/_", "#include \"oracle", "#include \"../oracle", "import oracle"):
                    assert needle not in txt, (f, needle)


def test_argument_errors_are_reported(crt_lib):
    cfg = api.make_config(0, 0, np.eye(4), np.eye(4))
    b = C.c_int32(); e = C.c_int32()
    assert crt_lib.crt_partition_spp_range(C.byref(cfg), C.byref(b), C.byref(e)) != 0
    assert b"empty image" in crt_lib.crt_last_error()
    cfg = api.make_config(8, 8, np.eye(4), np.eye(4), rank=3, world=2)
    assert crt_lib.crt_partition_pixel_count(C.byref(cfg)) < 0


def test_header_is_plain_c_and_links(tmp_path, crt_lib):
    """include/crt_b200.h is a C header (no C++ in the signatures): compile a C99 translation unit against it, link libcrt_b200.so and call
    host-only entry points."""
    import subprocess
    src = tmp_path / "abi_demo.c"
    src.write_text(r'''
#include "crt_b200.h"
#include <stdio.h>
int main(void) {
    float r2c[16], c2w[16], fit[3];
    const float pos[3] = {0, 0, 0}, look[3] = {0, 0, 1}, right[3] = {1, 0, 0}, up[3] = {0, 1, 0}, rgb[3] = {0.2f, 0.6f, 0.3f};
    crt_render_config cfg;
    crt_context* ctx = 0;
    if (crt_camera_matrices(0, 1.0f, 1000.0f, 0.0f, 0.0f, 45.0f, pos, look, right, up, 640.0f, 480.0f, r2c, c2w) != 0) return 1;
    if (crt_rgb2spec_fit(rgb, fit) != 0) return 2;
    cfg.filter_kind = 2; cfg.filter_sigma = 0.5f;
    printf("version %d r2c[0] %g fit %g %g %g sizeof(cfg) %u\n", crt_version(), r2c[0], fit[0], fit[1], fit[2], (unsigned)sizeof cfg);
    if (crt_context_create(0, &ctx) != 0) printf("no gpu: %s\n", crt_last_error()); else crt_context_destroy(ctx);
    return 0;
}
''')
    pkg = os.path.join(ROOT, "computational_ray_tracer_b200")
    exe = tmp_path / "abi_demo"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", pkg, "-l:libcrt_b200.so", f"-Wl,-rpath,{pkg}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "version" in r.stdout, r.stdout + r.stderr


def test_nccl_is_resolved_at_run_time_not_linked(crt_lib):
    """The library has no link-time dependency on libnccl (it must load on a box without it); the entry points bind to the copy in the
    process, $CRT_NCCL_LIB or the system's at first use and report which."""
    import subprocess
    from computational_ray_tracer_b200 import _capi, api
    needed = subprocess.run(["readelf", "-d", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "libnccl" not in needed
    version, origin = api.Context.nccl_version()
    assert version >= 22000 and "libnccl" in origin
    uid = api.Context.nccl_unique_id()
    assert uid.shape == (128,) and uid.any()
