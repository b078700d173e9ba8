"""Host-side logic of the multi-GPU paths (pure Python, no CUDA): block-cyclic ownership maps, sharding of test
points and classes, and the small torch.distributed exchanges around libgpx's own NCCL communicator.

Everything here runs on the CPU with the ``gloo`` backend (tests/test_parallel_cpu.py, world size 2); on the GPU box
the same functions run over ``nccl``.  SURVEY.md section 8(e)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


# ---- 1-D block-cyclic layout of block columns (the P x 1 case of the 2-D block-cyclic scheme) ----------------
# Two block -> rank maps (csrc/common.cuh gpx_cyc_*): plain cyclic (j mod P) and the default boustrophedon ("snake")
# order 0..P-1, P-1..0, 0.. which pairs a long column of the triangular matrix with a short one on every rank.
def mg_padded_dim(n: int, nb: int, world: int, snake: bool = True) -> int:
    unit = nb * world * (2 if snake else 1)
    return ((n + unit - 1) // unit) * unit


def owner_of_block(j: int, world: int, snake: bool = True) -> int:
    if not snake:
        return j % world
    pos = j % (2 * world)
    return pos if pos < world else 2 * world - 1 - pos


def local_index_of_block(j: int, world: int, snake: bool = True) -> int:
    if not snake:
        return j // world
    return 2 * (j // (2 * world)) + (1 if j % (2 * world) >= world else 0)


def global_block(q: int, rank: int, world: int, snake: bool = True) -> int:
    if not snake:
        return q * world + rank
    return (q // 2) * 2 * world + (2 * world - 1 - rank if q % 2 else rank)


def local_blocks(nblk: int, rank: int, world: int, snake: bool = True) -> List[int]:
    """Global block indices owned by ``rank``."""
    return [j for j in range(nblk) if owner_of_block(j, world, snake) == rank]


def blocks_below(j: int, rank: int, world: int, snake: bool = True) -> int:
    """Number of the rank's blocks with global index < j."""
    if j <= 0:
        return 0
    if not snake:
        return (j - rank + world - 1) // world if j > rank else 0
    cyc, rem = divmod(j, 2 * world)
    return 2 * cyc + (1 if rem > rank else 0) + (1 if rem > 2 * world - 1 - rank else 0)


def first_local_block_after(j: int, rank: int, world: int, snake: bool = True) -> int:
    """Smallest local index q whose global block is > j (trailing-update range)."""
    return blocks_below(j + 1, rank, world, snake)


def prefix_cols(grow_end: int, nb: int, rank: int, world: int, snake: bool = True) -> int:
    """Number of local columns whose global block starts below row ``grow_end`` (structure of L^-1's columns)."""
    return blocks_below((grow_end + nb - 1) // nb, rank, world, snake) * nb


# ---- sharding of independent work ---------------------------------------------------------------------------
def shard_range(m: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of m independent items (test points) for this rank; sizes differ by at most 1."""
    base, rem = divmod(m, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_classes(C: int, rank: int, world: int) -> List[int]:
    """Classes c = rank, rank+world, ... (10 classes on 8 GPUs -> at most 2 per GPU; SURVEY 8e multiclass row)."""
    return list(range(rank, C, world))


def gather_slices(local: np.ndarray, m: int, group=None) -> np.ndarray:
    """All-gather the per-rank slices produced with ``shard_range`` back into one array of length m (axis 0)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = [shard_range(m, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    buf = torch.zeros((width,) + tuple(local.shape[1:]), dtype=torch.float64, device=dev)
    buf[:local.shape[0]] = torch.as_tensor(np.ascontiguousarray(local), dtype=torch.float64, device=dev)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    parts = [o[:hi - lo].cpu().numpy() for o, (lo, hi) in zip(outs, sizes)]
    return np.concatenate(parts, axis=0)


def broadcast_bytes(payload: bytes, nbytes: int, src: int = 0, group=None) -> bytes:
    """Broadcast a small byte string (libgpx's 128-byte NCCL unique id) through torch.distributed."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    raw = list(payload) if dist.get_rank(group) == src else [0] * nbytes
    t = torch.tensor(raw, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().tolist())


def allreduce_sum_(tensor, group=None):
    """In-place sum over ranks of a device tensor (the E_c / R^T c / f reductions of the multiclass Newton step)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, group=group)
    return tensor
