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
    """Every rank holds the factor -- the replicated factor of the distributed fit (``Engine.mg_fit``; nothing is
    refitted per rank) or a single-GPU fit done on every rank; the m test points are split contiguously over ranks and
    the (mu, var) slices all-gathered."""
    world, rank = _world_rank()
    Xs = np.asarray(Xs, dtype=np.float64)
    m = Xs.shape[0]
    lo, hi = P.shard_range(m, rank, world)
    if hi > lo:
        mu, var, _ = eng.predict(fit, Xs[lo:hi], m_total=m, row0=lo)
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


class BinaryLaplaceDistributed:
    """Binary Laplace approximation (textbook Newton, R&W Alg. 3.1) with every B = I + W^1/2 K W^1/2 factored by the
    block-cyclic multi-GPU Cholesky (SURVEY 8e row "binary Laplace": GP_binary_classification.py:107 on P GPUs).  One
    Newton iteration is ONE C-ABI call (gpx_mg_laplace_binary_step); B is built block-cyclically from X, never whole."""

    def __init__(self, eng: Engine, X, sigma: float = 1.0, l: float = 1.0, nb: int = 256):
        from ._lib import COV_SE
        world, _ = _world_rank()
        if world > 1:
            eng.mg_init()
        self.eng, self.nb, self.kind = eng, int(nb), COV_SE
        self.theta = np.array([float(sigma), float(l)])
        self.Xd = eng.to_device(X)
        self.n = self.Xd.shape[0]
        lay = eng.mg_layout(self.n, self.nb)
        self.npad = lay["npad"]
        self.ws = eng.mg_workspace(self.n, self.nb)
        self.K = eng.cov(COV_SE, self.Xd, self.Xd, self.theta, same_x=True, n1p=None)   # replicated, for the two mat-vecs
        self.vws = eng.empty(8 * self.npad)
        self.err = eng.empty(2)
        self.errors = []
        self.f = self.g = self.w = self.sw = None

    def step(self, yd, f, f_new):
        import ctypes
        from ._lib import check
        eng = self.eng
        thp = self.theta.ctypes.data_as(ctypes.c_void_p)
        check(eng.lib.gpx_mg_laplace_binary_step(eng.h, self.kind, eng._p(self.Xd), self.n, self.Xd.shape[1], thp, 2, eng._p(self.K),
                                                 self.K.stride(0), eng._p(yd), eng._p(f), self.nb, eng._p(self.ws), eng._p(self.vws),
                                                 eng._p(f_new), eng._p(self.err)), "gpx_mg_laplace_binary_step")

    def fit_newton(self, y, tolerance: float = 1e-10, max_iter: int = 100):
        eng = self.eng
        yd = eng.zeros(self.npad)
        yd[:self.n] = eng.to_device(np.asarray(y, dtype=np.float64).reshape(-1))
        f, f_new = eng.zeros(self.npad), eng.zeros(self.npad)
        self.errors = []
        for _ in range(max_iter):
            eng._sync_stream()
            self.step(yd, f, f_new)
            err = float(self.err[0].item())
            self.errors.append(err)
            f, f_new = f_new, f
            if err <= tolerance:
                break
        self.f = f
        npd = self.npad
        self.g, self.w, self.sw = eng.zeros(npd), eng.zeros(npd), eng.zeros(npd)
        from ._lib import check
        eng._sync_stream()
        check(eng.lib.gpx_logistic_terms(eng.h, 1, self.n, eng._p(yd), eng._p(f), eng._p(self.g), eng._p(self.w), eng._p(self.sw)),
              "gpx_logistic_terms")                     # terms at the converged mode (what prediction uses)
        return len(self.errors)
