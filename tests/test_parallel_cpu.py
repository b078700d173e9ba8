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


@pytest.mark.parametrize("snake", [True, False])
def test_block_cyclic_maps_are_consistent(snake):
    for world in (1, 2, 3, 8):
        nblk = 6 * world
        seen = []
        for r in range(world):
            blocks = P.local_blocks(nblk, r, world, snake)
            assert len(blocks) == nblk // world                      # every rank owns the same number of blocks
            for q, j in enumerate(blocks):
                assert P.owner_of_block(j, world, snake) == r and P.local_index_of_block(j, world, snake) == q
                assert P.global_block(q, r, world, snake) == j
            seen += blocks
        assert sorted(seen) == list(range(nblk))
        for r in range(world):
            for j in range(nblk):
                q = P.first_local_block_after(j, r, world, snake)
                assert q == len([b for b in P.local_blocks(nblk, r, world, snake) if b <= j])
    assert P.mg_padded_dim(65536, 512, 8, snake) == 65536
    assert P.mg_padded_dim(1000, 256, 3, False) == 1536 and P.mg_padded_dim(1000, 256, 3, True) == 1536
    assert P.mg_padded_dim(1600, 256, 3, True) == 3072


def test_snake_map_balances_triangular_work():
    """Work of a column of a lower-triangular matrix ~ (rows below it)^2: plain cyclic gives rank 0 ~9 % more than rank 7 at
    N = 65536, nb = 256, P = 8; the boustrophedon map brings the spread to 0.25 %."""
    nblk, world = 256, 8
    for snake, bound in ((False, 0.05), (True, 0.005)):
        w = [sum((nblk - j) ** 2 for j in P.local_blocks(nblk, r, world, snake)) for r in range(world)]
        spread = (max(w) - min(w)) / (sum(w) / world)
        assert (spread > bound) if not snake else (spread < bound), (snake, spread)


@pytest.mark.parametrize("snake", [True, False])
def test_library_block_maps_match_the_python_mirror(snake):
    """The C maps the kernels use (gpx_mg_block_*; no GPU needed) against gaussian_process_b200.parallel."""
    from gaussian_process_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        pytest.skip("libgpx.so not built")
    lib = _lib.load()
    lib.gpx_mg_set_layout(int(snake))
    try:
        for world in (1, 2, 3, 8):
            for j in range(0, 7 * world):
                assert lib.gpx_mg_block_owner(j, world) == P.owner_of_block(j, world, snake)
                for r in range(world):
                    assert lib.gpx_mg_blocks_below(j, world, r) == P.blocks_below(j, r, world, snake)
            for r in range(world):
                for q in range(9):
                    assert lib.gpx_mg_block_global(q, world, r) == P.global_block(q, r, world, snake)
            assert lib.gpx_mg_padded_dim(1000, 256, world) == P.mg_padded_dim(1000, 256, world, snake)
    finally:
        lib.gpx_mg_set_layout(1)


@pytest.mark.parametrize("snake", [True, False])
def test_prefix_cols_matches_brute_force(snake):
    nb = 256
    for world in (1, 2, 4):
        for rank in range(world):
            owned = P.local_blocks(64, rank, world, snake)
            for grow_end in range(0, 10 * nb + 1, 128):
                brute = sum(nb for j in owned if j * nb < grow_end)
                assert P.prefix_cols(grow_end, nb, rank, world, snake) == brute


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
