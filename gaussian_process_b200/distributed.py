"""Multi-GPU drivers for the parts of the path that shard without a data-path exchange:
test-point prediction (independent columns of K_s) and the per-class factorisations of the multiclass Laplace
step.  One process per GPU; torch.distributed is the plumbing."""
from __future__ import annotations

import numpy as np

from . import parallel as P
from .engine import Engine, GPFit
from .laplace import MultiLaplaceNewton


def _world_rank():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def predict_sharded(eng: Engine, fit: GPFit, Xs: np.ndarray):
    """Every rank holds the factor (fit replicated or re-done per rank: it is O(N^3/3) once, prediction is
    O(N^2 m)); the m test points are split contiguously over ranks and the (mu, var) slices all-gathered."""
    world, rank = _world_rank()
    Xs = np.asarray(Xs, dtype=np.float64)
    m = Xs.shape[0]
    lo, hi = P.shard_range(m, rank, world)
    if hi > lo:
        mu, var, _ = eng.predict(fit, Xs[lo:hi])
        loc = np.stack([eng.to_host(mu), eng.to_host(var)], axis=1)
    else:
        loc = np.zeros((0, 2))
    full = P.gather_slices(loc, m)
    return full[:, 0].copy(), full[:, 1].copy()


def multiclass_newton_sharded(eng: Engine, Ksub_dev, y, C: int, n: int, tolerance: float = 1e-8, max_iter: int = 100):
    """Textbook softmax Laplace with class c on rank c mod P: per-class B_c factorisations and E_c stay local,
    sum_c E_c (n x n), R^T c (n) and f (C x n) are all-reduced.  Returns the fitted MultiLaplaceNewton."""
    world, rank = _world_rank()
    if world > 1:
        eng.mg_init()          # libgpx's own NCCL communicator: the step all-reduces sum_c E_c, R^T c and f inside the call
    model = MultiLaplaceNewton(eng, Ksub_dev, C, n, classes=P.shard_classes(C, rank, world))
    model.fit(y, tolerance, max_iter)
    return model
