"""Multi-GPU host logic on CPU: two gloo ranks take their partition from the C ABI (crt_partition_*), evaluate their
share of the samples (the oracle stands in for the kernels here), and one reduce(sum) to rank 0 must reproduce the
single-process film -- bit-exactly for interleaved tiles, to fp32 summation-order tolerance for spp ranges."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP = 48, 27, 8


def _film_from_samples(orc, O, params, pixels, idx_begin, idx_end):
    film = np.zeros((W * H, 4), np.float32)
    for idx in range(idx_begin, idx_end):          # the product also accumulates one sample index per wave
        s = orc.eval_samples(params, pixels, np.full(len(pixels), idx, np.int32))
        film[pixels, :3] += s["weight"][:, None] * s["rgb"]
        film[pixels, 3] += s["weight"]
    return film


def _worker(rank, world, port, partition, out_path):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_lib as O
    from computational_ray_tracer_b200 import api, scenes
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    meshes = scenes.heightfield(24, with_light=False)
    orc = O.OracleScene(); orc.set_model(meshes); orc.build_octree()
    r2c, c2w = api.camera_matrices(0, 1.0, 1000.0, 45.0, (0, 0, 0), (0, 0, 1), (0, 1, 0), W, H)
    params = O.make_params(W, H, r2c, c2w, xs=4, ys=2, spp_end=SPP)
    cfg = api.make_config(W, H, r2c, c2w, xs=4, ys=2, spp_begin=0, spp_end=SPP, rank=rank, world=world, partition=partition, tile=(8, 8))
    pixels = api.partition_pixels(cfg)
    b, e = api.partition_spp_range(cfg)
    film = torch.from_numpy(_film_from_samples(orc, O, params, pixels, b, e))
    counts = torch.tensor([len(pixels) * (e - b)], dtype=torch.int64)
    dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)
    dist.reduce(counts, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.savez(out_path, film=film.numpy(), samples=counts.numpy())
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("partition", [0, 1])
def test_two_rank_partition_reduces_to_the_single_process_film(tmp_path, oracle, crt_lib, partition):
    from computational_ray_tracer_b200 import api, scenes
    out = str(tmp_path / f"film{partition}.npz")
    mp.spawn(_worker, args=(2, _free_port(), partition, out), nprocs=2, join=True)
    got = np.load(out)
    assert int(got["samples"][0]) == W * H * SPP                      # every (pixel, sample index) exactly once
    meshes = scenes.heightfield(24, with_light=False)
    orc = oracle.OracleScene(); orc.set_model(meshes); orc.build_octree()
    r2c, c2w = api.camera_matrices(0, 1.0, 1000.0, 45.0, (0, 0, 0), (0, 0, 1), (0, 1, 0), W, H)
    params = oracle.make_params(W, H, r2c, c2w, xs=4, ys=2, spp_end=SPP)
    want = _film_from_samples(orc, oracle, params, np.arange(W * H, dtype=np.int32), 0, SPP)
    assert np.array_equal(got["film"][:, 3], want[:, 3])
    if partition == 0:
        assert np.array_equal(got["film"].view(np.uint32), want.view(np.uint32))      # tiles: x + 0 is exact
    else:
        np.testing.assert_allclose(got["film"], want, rtol=2e-6, atol=1e-6)            # spp ranges: (a+b)+(c+d) vs ((a+b)+c)+d
    assert want[:, :3].max() > 0


def test_partitions_cover_the_image_disjointly(crt_lib):
    from computational_ray_tracer_b200 import api
    eye = np.eye(4, dtype=np.float32)
    for (w, h, world, tile) in [(48, 27, 2, (8, 8)), (1920, 1080, 8, (32, 32)), (50, 30, 3, (16, 4)), (7, 5, 4, (32, 32))]:
        seen = np.zeros(w * h, np.int32)
        for r in range(world):
            cfg = api.make_config(w, h, eye, eye, rank=r, world=world, partition=0, tile=tile)
            px = api.partition_pixels(cfg)
            assert (np.diff(px) > 0).all()
            seen[px] += 1
        assert (seen == 1).all()
    for (spp, world) in [(64, 8), (10, 3), (3, 4), (1, 2)]:
        ranges = [api.partition_spp_range(api.make_config(4, 4, eye, eye, spp_begin=2, spp_end=2 + spp, rank=r, world=world, partition=1)) for r in range(world)]
        assert ranges[0][0] == 2 and ranges[-1][1] == 2 + spp
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        assert max(e - b for b, e in ranges) - min(e - b for b, e in ranges) <= 1
