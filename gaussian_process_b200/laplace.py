"""Device-resident Laplace approximations (binary logistic, multiclass softmax) on the gpx engine.

Two modes each:
  * reference-faithful -- reproduces the shipped arithmetic of GP_binary_classification.py:86-133 and
    GP_multi_classification.py:129-176 including their quirks (W and the gradient frozen at f_prior;
    y=-1 gradient sign; point-major Pi vs class-major D; literal stride 60; one triangular factor in
    the multiclass update),
  * textbook -- Rasmussen & Williams Alg. 3.1 / 3.3 (Newton on B = I + W^1/2 K W^1/2, per-class
    factorisations), the modes the large configurations use.
All vectors live in HBM, padded to multiples of 128 with zeros; the only per-iteration host traffic
is the scalar convergence error.
"""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional

import numpy as np

from ._lib import COV_SE, GPX_TILE, check
from .engine import Engine, padded


def _chk(eng: Engine, status: int, what: str):
    check(status, what)


class BinaryLaplace:
    """Binary GP classifier, logistic likelihood, Laplace approximation."""

    def __init__(self, eng: Engine, K, n: int):
        """``K``: padded (npad, npad) device tensor holding the full symmetric covariance in [:n,:n]."""
        self.eng, self.K, self.n, self.npad = eng, K, int(n), K.shape[0]
        self.L = self.dinv = self.g = self.w = self.sw = self.f = None
        self.Linv = None   # optional explicit inv(L) (padded, lower) used instead of L for prediction
        self.errors: List[float] = []

    # -- shared pieces ---------------------------------------------------------------------------
    def _terms(self, mode: int, y, f):
        eng = self.eng
        g, w, sw = eng.zeros(self.npad), eng.zeros(self.npad), eng.zeros(self.npad)
        eng._sync_stream()
        _chk(eng, eng.lib.gpx_logistic_terms(eng.h, mode, self.n, eng._p(y), eng._p(f), eng._p(g), eng._p(w), eng._p(sw)),
             "gpx_logistic_terms")
        return g, w, sw

    def _factor_B(self, sw):
        eng = self.eng
        B = eng.empty(self.npad, self.npad)
        eng._sync_stream()
        _chk(eng, eng.lib.gpx_build_B(eng.h, eng._p(self.K), eng._p(sw), self.n, self.npad, B.stride(0), eng._p(B)), "gpx_build_B")
        dinv = eng.potrf(B)
        return B, dinv

    def _newton_step(self, f, g, w, sw, L, dinv, Linv=None):
        """b = W f + g; a = b - W^1/2 B^-1 W^1/2 K b; f_new = K a  (GP_binary...:109-111).
        With ``Linv`` (explicit inv(L), as the reference forms at :108) B^-1 is applied as L_inv^T (L_inv .): two
        HBM-bound GEMVs without the sequential dependency of a triangular solve."""
        eng, n, npad = self.eng, self.n, self.npad
        b = eng.vec_op(5, n, eng.zeros(npad), x=w, y=f, z=g)
        t = eng.gemv(self.K, b, eng.zeros(npad), m=n, n=n)
        eng.vec_op(2, n, t, x=sw, y=t)
        if Linv is not None:
            u = eng.gemv(Linv, t, eng.zeros(npad), m=n, n=n)
            t = eng.gemv(Linv, u, eng.zeros(npad), trans=True, m=n, n=n)
        else:
            eng.potrs_vec(L, dinv, t, big=eng.block_inverses(L, dinv))
        a = eng.vec_op(3, n, eng.zeros(npad), x=b, y=sw, z=t)
        f_new = eng.gemv(self.K, a, eng.zeros(npad), m=n, n=n)
        d = eng.vec_op(6, n, eng.zeros(npad), x=f_new, y=f)
        err = math.sqrt(eng.dot(d, d, n))
        return f_new, a, err

    # -- reference-faithful ------------------------------------------------------------------------
    def _workspace(self):
        if getattr(self, "_ws", None) is None:
            eng = self.eng
            self._ws = eng.empty(int(eng.lib.gpx_laplace_binary_ws_elems(self.npad)))
            self._B = eng.empty(self.npad, self.npad)
            self._dinv = eng.empty(self.npad // GPX_TILE, GPX_TILE, GPX_TILE)
            self._err = eng.empty(2)
        return self._ws

    def fit_reference(self, y, f_prior, tolerance: float = 1e-4, max_iter: int = 10000, on_iter=None):
        """W, gradient evaluated at ``f_prior`` every iteration (never at f): B is factored once and the
        iterate converges linearly.  The whole loop is ONE libgpx call (gpx_laplace_binary_ref_fit).
        Returns the number of iterations."""
        eng = self.eng
        yd = self._pad_vec(y)
        fp = self._pad_vec(f_prior)
        ws = self._workspace()
        Linv = eng.empty(self.npad, self.npad)
        self.g, self.w, self.sw, f = eng.empty(self.npad), eng.empty(self.npad), eng.empty(self.npad), eng.empty(self.npad)
        errs = np.zeros(max_iter)
        iters = ctypes.c_int(0)
        eng._sync_stream()
        _chk(eng, eng.lib.gpx_laplace_binary_ref_fit(
            eng.h, eng._p(self.K), self.n, self.npad, self.K.stride(0), eng._p(yd), eng._p(fp), float(tolerance), int(max_iter),
            eng._p(self._B), eng._p(self._dinv), eng._p(Linv), eng._p(ws), eng._p(f), eng._p(self.g), eng._p(self.w),
            eng._p(self.sw), ctypes.c_void_p(errs.ctypes.data), ctypes.byref(iters)), "gpx_laplace_binary_ref_fit")
        self.L, self.dinv, self._Linv_full, self.f = self._B, self._dinv, Linv, f
        self.errors = [float(e) for e in errs[:iters.value]]
        if on_iter is not None:
            for i, e in enumerate(self.errors):
                on_iter(i, e)
        return len(self.errors)

    # -- textbook Newton (R&W Alg 3.1) -----------------------------------------------------------
    def newton_step(self, yd, f, f_new):
        """One device-resident Newton iteration (gpx_laplace_binary_step): enqueues only, no allocation, no sync."""
        eng = self.eng
        ws = self._workspace()
        _chk(eng, eng.lib.gpx_laplace_binary_step(eng.h, eng._p(self.K), self.n, self.npad, self.K.stride(0), eng._p(yd),
                                                  eng._p(f), eng._p(self._B), eng._p(self._dinv), eng._p(ws), eng._p(f_new),
                                                  eng._p(self._err)), "gpx_laplace_binary_step")

    def fit_newton(self, y, f0=None, tolerance: float = 1e-10, max_iter: int = 100, on_iter=None):
        eng = self.eng
        yd = self._pad_vec(y)
        f = eng.zeros(self.npad) if f0 is None else self._pad_vec(f0)
        f_new = eng.zeros(self.npad)
        self._workspace()
        self.errors = []
        for i in range(max_iter):
            eng._sync_stream()
            self.newton_step(yd, f, f_new)
            err = float(self._err[0].item())             # the one read-back of the iteration (convergence decision)
            eng.potrf_check()
            self.errors.append(err)
            f, f_new = f_new, f
            if on_iter is not None:
                on_iter(i, err)
            if err <= tolerance:
                break
        self.g, self.w, self.sw = self._terms(1, yd, f)
        self.L, self.dinv = self._factor_B(self.sw)
        self.f = f
        return len(self.errors)

    def _pad_vec(self, v):
        eng = self.eng
        out = eng.zeros(self.npad)
        out[:self.n] = eng.to_device(np.asarray(v, dtype=np.float64).reshape(-1) if not hasattr(v, "data_ptr") else v.reshape(-1)[:self.n])
        return out

    # -- prediction ----------------------------------------------------------------------------------
    def predict(self, X_train_dev, Xs_dev, sigma: float, l: float = 1.0):
        """f*_mean = k*^T g, var = k** - |L^-1 W^1/2 k*|^2 for many test points (GP_binary...:148-153).
        Returns host arrays (f_mean, var)."""
        eng = self.eng
        m = Xs_dev.shape[0]
        theta = [float(sigma), float(l)]
        Ks = eng.cov(COV_SE, X_train_dev, Xs_dev, theta)
        mu, var = eng.empty(m), eng.empty(m)
        eng._sync_stream()
        _chk(eng, eng.lib.gpx_predict_moments(eng.h, eng._p(Ks), None, self.n, m, Ks.stride(0), eng._p(self.g), None,
                                              eng._p(mu), None), "gpx_predict_moments")
        _chk(eng, eng.lib.gpx_scale_rows(eng.h, self.n, Ks.shape[1], Ks.stride(0), eng._p(self.sw), eng._p(Ks)), "gpx_scale_rows")
        if self.Linv is not None:                                   # v = L_inv (W^1/2 k*) as a GEMM
            V = eng.empty(self.npad, Ks.shape[1])
            eng.gemm(self.Linv, Ks, V, a_kmajor=True, b_kmajor=False, M=self.npad, N=Ks.shape[1], K=self.npad)
            Ks = V
        else:
            eng.trsm(self.L, self.dinv, Ks, trans=False)
        kss = eng.vec_op(4, m, eng.empty(m), a=float(sigma) ** 2)
        _chk(eng, eng.lib.gpx_predict_moments(eng.h, None, eng._p(Ks), self.n, m, Ks.stride(0), None, eng._p(kss), None,
                                              eng._p(var)), "gpx_predict_moments")
        return eng.to_host(mu), eng.to_host(var)

    def L_inverse_host(self) -> np.ndarray:
        """Dense inv(L) on the host (the reference returns it, GP_binary...:108,133)."""
        eng = self.eng
        Li = getattr(self, "_Linv_full", None)
        if Li is None:
            Li = self.L.clone()
            eng.trtri(Li, self.dinv)
        return eng.to_host(Li[:self.n, :self.n])


class MultiLaplaceReference:
    """Reference-faithful GP_multi_classification.model_training2 (:129-176) on the device."""

    def __init__(self, eng: Engine, K_full, C: int, n: int, stride: int = 60, s: float = 3.0, step_size: float = 1e-4):
        self.eng, self.C, self.n, self.stride, self.s, self.step = eng, int(C), int(n), int(stride), float(s), float(step_size)
        self.N = self.C * self.n
        self.npad = K_full.shape[0]
        if (self.C - 1) * self.stride + self.n > self.N:
            # the reference would raise IndexError at GP_multi...:55 (f[j*60+i] out of range)
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % ((self.C - 1) * self.stride + self.n - 1, self.N))
        # L = chol(s I + K); K^-1 := L^-T L^-1 (constant across iterations; the reference recomputes it)
        A = K_full.clone()
        if self.npad > self.N:  # identity padding so the padded factorisation equals the unpadded one
            ones = eng.vec_op(4, self.npad - self.N, eng.empty(self.npad - self.N), a=1.0)
            eng._sync_stream()
            p0 = ctypes.c_void_p(A.data_ptr() + 8 * (self.N * A.stride(0) + self.N))
            _chk(eng, eng.lib.gpx_copy_strided(eng.h, self.npad - self.N, eng._p(ones), 1, p0, A.stride(0) + 1), "gpx_copy_strided")
        d = eng.diag(A, self.N)
        eng.vec_op(1, self.N, d, a=self.s, x=d, y=eng.vec_op(4, self.N, eng.empty(self.N), a=1.0))
        eng._sync_stream()
        _chk(eng, eng.lib.gpx_copy_strided(eng.h, self.N, eng._p(d), 1, eng._p(A), A.stride(0) + 1), "gpx_copy_strided")
        dinv = eng.potrf(A)
        eng.trtri(A, dinv)
        self.Kinv = eng.lauum(A)
        _chk(eng, eng.lib.gpx_symmetrize(eng.h, self.npad, eng._p(self.Kinv), self.Kinv.stride(0)), "gpx_symmetrize")
        self.errors: List[float] = []
        self.pi = self.f = None

    def fit(self, y, tolerance: float = 0.01, max_iter: int = 10000, on_iter=None):
        eng, C, n, N, npad = self.eng, self.C, self.n, self.N, self.npad
        yd = eng.zeros(npad)
        yd[:N] = eng.to_device(np.asarray(y, dtype=np.float64).reshape(-1))
        f = eng.zeros(npad)
        lib = eng.lib
        self.errors = []
        pi = eng.zeros(npad)
        for it in range(max_iter):
            eng._sync_stream()
            _chk(eng, lib.gpx_softmax_classes(eng.h, C, n, self.stride, eng._p(f), eng._p(pi)), "gpx_softmax_classes")
            H = eng.empty(npad, npad)
            _chk(eng, lib.gpx_multi_ref_hessian(eng.h, C, n, self.stride, eng._p(self.Kinv), self.Kinv.stride(0), eng._p(pi),
                                                self.s, npad, eng._p(H)), "gpx_multi_ref_hessian")
            dinv2 = eng.potrf(H)                                           # GP_multi...:155
            rhs = eng.gemv(self.Kinv, f, eng.zeros(npad), alpha=1.0 - self.step, m=N, n=N)
            wf = eng.zeros(npad)
            _chk(eng, lib.gpx_multi_ref_wf(eng.h, C, n, self.stride, eng._p(pi), eng._p(f), eng._p(wf)), "gpx_multi_ref_wf")
            eng.vec_op(1, N, rhs, a=1.0, x=rhs, y=wf)
            eng.vec_op(1, N, rhs, a=1.0, x=rhs, y=yd)
            eng.vec_op(1, N, rhs, a=1.0, x=rhs, y=pi)                      # :157 (+ pi, as shipped)
            eng.trsv(H, dinv2, rhs, trans=False)                           # :158 inv(L2) . sum
            d = eng.vec_op(6, N, eng.zeros(npad), x=rhs, y=f)
            err = math.sqrt(eng.dot(d, d, N))
            self.errors.append(err)
            f = rhs
            if on_iter is not None:
                on_iter(it, err)
            if err <= tolerance:
                break
        self.f, self.pi = f, pi
        return len(self.errors)


class MultiLaplaceNewton:
    """Textbook multiclass Laplace (R&W Alg. 3.3) with one shared n x n covariance block ``Ksub`` --
    the structure of the reference's block_diag(K_sub x C) (GP_multi...:233-238).  Classes may be sharded
    across ranks: ``classes`` lists the classes this process owns; sums over classes (sum_c E_c, R^T c, f) are
    all-reduced inside libgpx over the engine's NCCL communicator (``Engine.mg_init()``) when the world has > 1 rank.
    One iteration is ONE C-ABI call (``gpx_laplace_multi_step``) on a workspace allocated once."""

    def __init__(self, eng: Engine, Ksub, C: int, n: int, classes: Optional[List[int]] = None, allreduce=None,
                 concurrency: int = 4):
        self.eng, self.K, self.C, self.n, self.npad = eng, Ksub, int(C), int(n), Ksub.shape[0]
        self.classes = list(range(C)) if classes is None else list(classes)
        self.errors: List[float] = []
        self.f = self.pi = None
        # the per-class factorisations are independent: issue them through `concurrency` handles / CUDA streams so the
        # latency-bound parts of one class overlap with the DMMA-bound parts of another
        self.nstreams = max(1, min(int(concurrency), len(self.classes)))
        self._lanes = None
        self.allreduce = allreduce   # kept for API compatibility: libgpx all-reduces over its own communicator

    def _make_lanes(self):
        from .engine import new_engine
        eng = self.eng
        T = eng.torch
        lanes = []
        for k in range(self.nstreams):
            st = T.cuda.Stream(device=eng.device)
            with T.cuda.stream(st):
                e = new_engine(eng.device.index)      # own handle: own helper streams, scratch and pivot flag
                e._sync_stream()
                lanes.append(dict(eng=e, stream=st))
        return lanes

    def _prepare(self):
        if self._lanes is None:
            eng, npad = self.eng, self.npad
            self._lanes = self._make_lanes()
            nl = len(self._lanes)
            self._lane_arr = (ctypes.c_void_p * nl)(*[ln["eng"].h for ln in self._lanes])
            self._ws = eng.empty(int(eng.lib.gpx_laplace_multi_ws_elems(npad, self.C, nl)))
            self._Linv = eng.empty(max(1, len(self.classes)), npad, npad)
            self._cls = (ctypes.c_int * max(1, len(self.classes)))(*self.classes)
            self._err = eng.empty(2)

    def step(self, yd, f, f_new, pi):
        """One device-resident Alg-3.3 iteration (gpx_laplace_multi_step): enqueues only."""
        eng = self.eng
        _chk(eng, eng.lib.gpx_laplace_multi_step(
            eng.h, self._lane_arr, len(self._lanes), eng._p(self.K), self.n, self.npad, self.K.stride(0), self.C, self._cls,
            len(self.classes), eng._p(yd), eng._p(f), eng._p(self._ws), eng._p(self._Linv), eng._p(f_new), eng._p(pi),
            eng._p(self._err)), "gpx_laplace_multi_step")

    def fit(self, y, tolerance: float = 1e-8, max_iter: int = 100, on_iter=None):
        eng, C, n = self.eng, self.C, self.n
        yd = eng.to_device(np.asarray(y, dtype=np.float64).reshape(C, n)) if not hasattr(y, "data_ptr") else y.reshape(C, n)
        yd = yd.contiguous()
        f, f_new, pi = eng.zeros(C, n), eng.zeros(C, n), eng.zeros(C, n)
        self._prepare()
        self.errors = []
        for it in range(max_iter):
            eng._sync_stream()
            self.step(yd, f, f_new, pi)
            err = float(self._err[0].item())             # the one read-back of the iteration
            for ln in self._lanes:
                ln["eng"].potrf_check()
            eng.potrf_check()
            self.errors.append(err)
            f, f_new = f_new, f
            if on_iter is not None:
                on_iter(it, err)
            if err <= tolerance:
                break
        eng._sync_stream()
        _chk(eng, eng.lib.gpx_softmax_classes(eng.h, C, n, n, eng._p(f), eng._p(pi)), "gpx_softmax_classes")
        self.f, self.pi = f, pi
        return len(self.errors)

    def predict(self, X_train_dev, Xs_dev, y, sigma: float = 1.0, l: float = 1.0):
        """f*_c = k*^T (y_c - pi_c) for every class and test point (GP_multi...:193-197) -> (m, C) host."""
        eng, C, n = self.eng, self.C, self.n
        m = Xs_dev.shape[0]
        Ks = eng.cov(COV_SE, X_train_dev, Xs_dev, [float(sigma), float(l)])
        yd = eng.to_device(np.asarray(y, dtype=np.float64).reshape(C, n))
        out = np.empty((m, C))
        for c in range(C):
            r = eng.vec_op(6, n, eng.zeros(self.npad), x=yd[c], y=self.pi[c])
            mu = eng.empty(m)
            eng._sync_stream()
            _chk(eng, eng.lib.gpx_predict_moments(eng.h, eng._p(Ks), None, n, m, Ks.stride(0), eng._p(r), None, eng._p(mu), None),
                 "gpx_predict_moments")
            out[:, c] = eng.to_host(mu)
        return out
