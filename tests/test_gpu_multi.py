"""Two real GPUs: every rank renders its partition into its own film and the films are summed onto rank 0 INSIDE the library
(crt_nccl_comm_create + crt_film_reduce -> ncclReduce over NVLink; torch.distributed/gloo only carries the 128-byte unique id).  Skipped with < 2 GPUs
(the single-GPU tests emulate ranks by accumulating partitions into one film; tests/test_cpu_partition_gloo.py covers the
host logic with gloo)."""
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
    uid = torch.from_numpy(api.Context.nccl_unique_id() if rank == 0 else np.zeros(128, np.uint8))
    dist.broadcast(uid, 0)
    ctx = api.Context(rank)
    ctx.nccl_init(world, rank, uid.numpy())
    meshes = scenes.cornell_box()
    ms = api.MeshSet(meshes); oc = api.Octtree_Model(ms)
    sc = api.Scene(ctx); mm = scenes.cornell_materials(sc); sc.set_model(oc, mesh_materials=mm); sc.commit()
    r2c, c2w = common.camera_1080p_like(W, H)
    kw = dict(mode=1, xs=4, ys=2, spp_begin=0, spp_end=SPP, max_depth=4)
    film = api.Film(ctx, W, H)
    sc.render(film, api.make_config(W, H, r2c, c2w, rank=rank, world=world, partition=partition, tile=(16, 8), **kw))
    film.reduce(0)                                                         # ncclReduce(sum) inside the library, on the context's stream
    ctx.synchronize()
    assert ctx.nccl_async_error() == 0
    if rank == 0:
        got = film.download()
        film.clear()
        sc.render(film, api.make_config(W, H, r2c, c2w, **kw))
        np.savez(out_path, got=got, want=film.download())
    dist.barrier()
    film.close(); sc.close(); oc.close(); ctx.close()                       # the context destroys its communicator
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
