"""Two real GPUs: every rank renders its partition into its own film and the films are summed onto rank 0 INSIDE the library
(crt_film_reduce_nccl -> ncclReduce over NVLink, communicator created here through NCCL's C API).  Skipped with < 2 GPUs
(the single-GPU tests emulate ranks by accumulating partitions into one film; tests/test_cpu_partition_gloo.py covers the
host logic with gloo)."""
import ctypes as C
import glob
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP = 96, 64, 8


def _nccl():
    import nvidia                                          # namespace package: torch's bundled NCCL lives under it
    libs = []
    for base in list(nvidia.__path__):
        libs += glob.glob(os.path.join(base, "nccl", "lib", "libnccl.so*"))
    libs += ["libnccl.so.2"]                               # system NCCL as a last resort
    return C.CDLL(libs[0], mode=C.RTLD_GLOBAL)            # RTLD_GLOBAL: the library resolves ncclReduce with dlsym


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


def _worker(rank, world, port, partition, out_path):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from computational_ray_tracer_b200 import api, scenes
    import common
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)          # only carries the NCCL unique id
    torch.cuda.set_device(rank)
    nccl = _nccl()
    uid = _UniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8)
    dist.broadcast(t, 0)
    C.memmove(C.byref(uid), bytes(t.tolist()), 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    ctx = api.Context(rank)
    meshes = scenes.cornell_box()
    ms = api.MeshSet(meshes); oc = api.Octtree_Model(ms)
    sc = api.Scene(ctx); mm = scenes.cornell_materials(sc); sc.set_model(oc, mesh_materials=mm); sc.commit()
    r2c, c2w = common.camera_1080p_like(W, H)
    kw = dict(mode=1, xs=4, ys=2, spp_begin=0, spp_end=SPP, max_depth=4)
    film = api.Film(ctx, W, H)
    sc.render(film, api.make_config(W, H, r2c, c2w, rank=rank, world=world, partition=partition, tile=(16, 8), **kw))
    assert ctx.L.crt_film_reduce_nccl(film.h, comm, 0) == 0, ctx.L.crt_last_error()
    ctx.synchronize()
    if rank == 0:
        got = film.download()
        film.clear()
        sc.render(film, api.make_config(W, H, r2c, c2w, **kw))
        np.savez(out_path, got=got, want=film.download())
    dist.barrier()
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)
    film.close(); sc.close(); oc.close(); ctx.close()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("partition", [0, 1])
def test_two_gpu_render_with_in_library_nccl_reduce(tmp_path, crt_lib, partition):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = str(tmp_path / f"multi{partition}.npz")
    mp.spawn(_worker, args=(2, _free_port(), partition, out), nprocs=2, join=True)
    d = np.load(out)
    assert np.array_equal(d["got"][:, 3], d["want"][:, 3])
    if partition == 0:
        assert np.array_equal(d["got"].view(np.uint32), d["want"].view(np.uint32))      # tiles: bit identical
    else:
        np.testing.assert_allclose(d["got"], d["want"], rtol=2e-6, atol=1e-6)            # spp ranges: summation order
    assert d["want"][:, :3].max() > 0
