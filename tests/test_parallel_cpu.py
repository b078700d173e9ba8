"""CPU tests (gloo, world size 2) of the host-side multi-GPU logic: block-cyclic maps, sharding, the unique-id
exchange and slice gathering.  The CUDA side of the same paths is covered by tests/test_gpu_multigpu.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussian_process_b200 import parallel as P


def test_block_cyclic_maps_are_consistent():
    for world in (1, 2, 3, 8):
        nblk = 5 * world
        seen = []
        for r in range(world):
            blocks = P.local_blocks(nblk, r, world)
            for q, j in enumerate(blocks):
                assert P.owner_of_block(j, world) == r and P.local_index_of_block(j, world) == q
                assert P.global_block(q, r, world) == j
            seen += blocks
        assert sorted(seen) == list(range(nblk))
        for r in range(world):
            for j in range(nblk):
                q = P.first_local_block_after(j, r, world)
                assert q == len([b for b in P.local_blocks(nblk, r, world) if b <= j])
    assert P.mg_padded_dim(65536, 512, 8) == 65536 and P.mg_padded_dim(1000, 256, 3) == 1536


def test_prefix_cols_matches_brute_force():
    nb = 256
    for world in (1, 2, 4):
        for rank in range(world):
            for grow_end in range(0, 10 * nb + 1, 128):
                brute = sum(nb for j in range(rank, 64, world) if j * nb < grow_end)
                assert P.prefix_cols(grow_end, nb, rank, world) == brute


def test_sharding_partitions():
    for m in (0, 1, 7, 100, 2048):
        for world in (1, 2, 3, 8):
            cuts = [P.shard_range(m, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == m
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1
    assert P.shard_classes(10, 1, 8) == [1, 9] and P.shard_classes(10, 7, 8) == [7]
    assert sorted(sum((P.shard_classes(10, r, 8) for r in range(8)), [])) == list(range(10))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ident = bytes(range(128)) if rank == 0 else b""
        got = P.broadcast_bytes(ident, 128, src=0)
        assert got == bytes(range(128))
        m = 11
        lo, hi = P.shard_range(m, rank, world)
        local = np.arange(lo, hi, dtype=np.float64)[:, None] * np.ones((1, 3))
        full = P.gather_slices(local, m)
        assert full.shape == (m, 3) and np.array_equal(full[:, 0], np.arange(m))
        t = torch.full((4,), float(rank + 1), dtype=torch.float64)
        P.allreduce_sum_(t)
        assert torch.all(t == sum(range(1, world + 1)))
        # class-sharded sum == full sum (the structure of the multiclass E_c reduction)
        C = 5
        rs = np.random.RandomState(0)
        E = rs.randn(C, 6, 6)
        part = torch.zeros(6, 6, dtype=torch.float64)
        for c in P.shard_classes(C, rank, world):
            part += torch.from_numpy(E[c])
        P.allreduce_sum_(part)
        assert np.allclose(part.numpy(), E.sum(0))
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


def test_gloo_world2_exchanges():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    assert dict(ret) == {0: 1, 1: 1}
