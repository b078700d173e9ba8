"""Host-side engine: thin, typed wrappers over the libgpx C ABI.

torch is used only as the device-memory allocator / stream provider (``tensor.data_ptr()`` crosses
the ABI); every floating-point operation of the GP path runs in libgpx's sm_100a kernels.  There is
no CPU fallback -- constructing an :class:`Engine` without the library or without a B200 raises.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import COV_CO2, COV_DELTA, COV_LIN, COV_LOWER, COV_PER, COV_SAME_X, COV_SE, GPX_TILE, GpxError, check

_KIND_NTHETA = {COV_SE: 2, COV_LIN: 1, COV_PER: 2, COV_CO2: 11}


def padded(n: int) -> int:
    return ((int(n) + GPX_TILE - 1) // GPX_TILE) * GPX_TILE


def _theta_array(theta: Sequence[float]):
    t = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).reshape(-1))
    return t, t.ctypes.data_as(ctypes.c_void_p)


@dataclass
class GPFit:
    """Device-resident state of one exact-GP fit (K + sI = L L^T)."""
    kind: int
    theta: np.ndarray
    s: float
    n: int
    npad: int
    X: "object"          # torch (n, D) device
    y: "object"          # torch (n,) device
    L: "object"          # torch (npad, npad) device: lower factor (upper zero); L^-1 after fit_grad
    dinv: "object"       # torch (npad/128, 128, 128): inverses of L's diagonal blocks
    alpha: "object"      # torch (npad,)
    lml: float
    y_alpha: float
    sum_log_diag: float
    grad: Optional[np.ndarray] = None
    Kinv: "object" = None
    L_is_inverse: bool = False
    big: "object" = None  # (D, bs): explicit inverses of the bs x bs diagonal blocks of L (lazy, for prediction)


class Engine:
    """One libgpx handle bound to one CUDA device."""

    def __init__(self, device: int = 0):
        import torch

        self.torch = torch
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise GpxError("no CUDA device visible: the gpx engine has no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        h = ctypes.c_void_p()
        check(self.lib.gpx_create(device, ctypes.byref(h)), "gpx_create")
        self.h = h
        self._sync_stream()

    # ------------------------------------------------------------------ plumbing
    def _sync_stream(self):
        s = self.torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.gpx_set_stream(self.h, ctypes.c_void_p(s)), "gpx_set_stream")

    def close(self):
        if getattr(self, "h", None):
            self.lib.gpx_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def launches(self) -> int:
        return int(self.lib.gpx_launch_count(self.h))

    def synchronize(self):
        check(self.lib.gpx_synchronize(self.h), "gpx_synchronize")

    def empty(self, *shape):
        return self.torch.empty(*shape, dtype=self.torch.float64, device=self.device)

    def zeros(self, *shape):
        return self.torch.zeros(*shape, dtype=self.torch.float64, device=self.device)

    def to_device(self, a, pinned: bool = False):
        """NumPy -> device tensor (float64, C-contiguous)."""
        torch = self.torch
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=torch.float64).contiguous()
        arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        t = torch.from_numpy(arr)
        if pinned:
            t = t.pin_memory()
        return t.to(self.device, non_blocking=pinned)

    def to_host(self, t) -> np.ndarray:
        return t.detach().cpu().numpy()

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(0 if t is None else t.data_ptr())

    # ------------------------------------------------------------------ A1-A3 covariance
    def cov(self, kind: int, X1, X2, theta, diag_add: float = 0.0, same_x: bool = False, lower: bool = False,
            out=None, with_grad: bool = False, delta: bool = False, n1p: Optional[int] = None):
        """K (padded) = k(X1, X2; theta) [+ diag_add I].  Returns K or (K, dK[ntheta]) device tensors.
        ``n1p`` overrides the padded row count (a factor padded to the multi-GPU block unit has more rows)."""
        self._sync_stream()
        n1, D = X1.shape
        n2 = X2.shape[0]
        n1p, n2p = (padded(n1) if n1p is None else int(n1p)), padded(n2)
        K = out if out is not None else self.empty(n1p, n2p)
        th, thp = _theta_array(theta)
        flags = (COV_SAME_X if same_x else 0) | (COV_LOWER if lower else 0) | (COV_DELTA if delta else 0)
        dK = self.empty(len(th), n1p, n2p) if with_grad else None
        check(self.lib.gpx_cov_build(self.h, kind, self._p(X1), n1, self._p(X2), n2, D, thp, len(th), float(diag_add), flags,
                                     self._p(K), n1p, n2p, K.stride(0), self._p(dK), n1p * n2p), "gpx_cov_build")
        return (K, dK) if with_grad else K

    # ------------------------------------------------------------------ A4/A5 dense linear algebra
    def potrf(self, A):
        """In-place lower Cholesky of a padded square tensor; returns the leaf-inverse buffer."""
        self._sync_stream()
        n = A.shape[0]
        dinv = self.empty(n // GPX_TILE, GPX_TILE, GPX_TILE)
        check(self.lib.gpx_potrf(self.h, self._p(A), n, A.stride(0), self._p(dinv)), "gpx_potrf")
        return dinv

    def potrf_async(self, A):
        """Like potrf() but returns after enqueueing (no host synchronisation); call potrf_check() later."""
        self._sync_stream()
        n = A.shape[0]
        dinv = self.empty(n // GPX_TILE, GPX_TILE, GPX_TILE)
        check(self.lib.gpx_potrf_async(self.h, self._p(A), n, A.stride(0), self._p(dinv)), "gpx_potrf_async")
        return dinv

    def potrf_check(self):
        """Synchronise this handle's stream and raise LinAlgError if any async factorisation since the last check failed."""
        self._sync_stream()
        info = ctypes.c_int(0)
        check(self.lib.gpx_potrf_info(self.h, ctypes.byref(info)), "gpx_potrf_info")
        check(info.value, "gpx_potrf_async")

    def trsv(self, L, dinv, x, trans: bool = False):
        self._sync_stream()
        check(self.lib.gpx_trsv(self.h, self._p(L), L.shape[0], L.stride(0), self._p(dinv), int(trans), self._p(x)), "gpx_trsv")
        return x

    def trsm(self, L, dinv, B, trans: bool = False):
        self._sync_stream()
        check(self.lib.gpx_trsm(self.h, self._p(L), L.shape[0], L.stride(0), self._p(dinv), int(trans), self._p(B),
                                B.shape[1], B.stride(0)), "gpx_trsm")
        return B

    def block_inverses(self, L, dinv):
        """Explicit inverses of the bs x bs diagonal blocks of L (bs up to 1024) -> (D, bs) or (None, 128) when the
        matrix is too small to benefit.  They shorten the serial chain of trsv / trsm by bs/128."""
        self._sync_stream()
        n = L.shape[0]
        bs = int(self.lib.gpx_block_size_for(n))
        if bs <= GPX_TILE or n < 2 * bs:
            return None, GPX_TILE
        D = self.empty(n // bs, bs, bs)
        work = self.empty(n * bs // 4)
        check(self.lib.gpx_block_inverses(self.h, self._p(L), n, L.stride(0), self._p(dinv), bs, self._p(D), self._p(work)),
              "gpx_block_inverses")
        return D, bs

    def trsv_big(self, L, D, bs, x, trans: bool = False):
        self._sync_stream()
        tmp = self.empty(bs)
        check(self.lib.gpx_trsv_big(self.h, self._p(L), L.shape[0], L.stride(0), self._p(D), bs, int(trans), self._p(x), self._p(tmp)),
              "gpx_trsv_big")
        return x

    def trsm_big(self, L, D, bs, B, trans: bool = False):
        self._sync_stream()
        tmp = self.empty(bs * B.shape[1])
        check(self.lib.gpx_trsm_big(self.h, self._p(L), L.shape[0], L.stride(0), self._p(D), bs, int(trans), self._p(B), B.shape[1],
                                    B.stride(0), self._p(tmp)), "gpx_trsm_big")
        return B

    def potrs_vec(self, L, dinv, x, big=None):
        """x <- (L L^T)^-1 x.  ``big`` = (D, bs) from block_inverses() uses the short-chain solves."""
        if big is not None and big[0] is not None:
            self.trsv_big(L, big[0], big[1], x, False)
            return self.trsv_big(L, big[0], big[1], x, True)
        self.trsv(L, dinv, x, False)
        return self.trsv(L, dinv, x, True)

    def trtri(self, L, dinv, work=None):
        self._sync_stream()
        n = L.shape[0]
        if work is None and n > GPX_TILE:
            work = self.empty((n // 2 + GPX_TILE) * (n // 2 + GPX_TILE))
        check(self.lib.gpx_trtri(self.h, self._p(L), n, L.stride(0), self._p(dinv), self._p(work)), "gpx_trtri")
        return L

    def lauum(self, Linv, out=None):
        self._sync_stream()
        n = Linv.shape[0]
        if out is None:
            out = self.zeros(n, n)
        check(self.lib.gpx_lauum(self.h, self._p(Linv), n, Linv.stride(0), self._p(out), out.stride(0)), "gpx_lauum")
        return out

    def gemm(self, A, B, C, a_kmajor: bool, b_kmajor: bool, M: int, N: int, K: int, alpha=1.0, beta=0.0):
        self._sync_stream()
        check(self.lib.gpx_gemm(self.h, int(a_kmajor), int(b_kmajor), M, N, K, float(alpha), self._p(A), A.stride(0),
                                self._p(B), B.stride(0), float(beta), self._p(C), C.stride(0)), "gpx_gemm")
        return C

    def gemv(self, A, x, y, trans: bool = False, alpha=1.0, beta=0.0, m=None, n=None):
        self._sync_stream()
        m = A.shape[0] if m is None else m
        n = A.shape[1] if n is None else n
        check(self.lib.gpx_gemv(self.h, int(trans), m, n, float(alpha), self._p(A), A.stride(0), self._p(x), float(beta),
                                self._p(y)), "gpx_gemv")
        return y

    def symv_lower(self, S, x, y, n=None):
        self._sync_stream()
        n = S.shape[0] if n is None else n
        check(self.lib.gpx_symv_lower(self.h, n, self._p(S), S.stride(0), self._p(x), self._p(y)), "gpx_symv_lower")
        return y

    def vec_op(self, op: int, n: int, out, a=0.0, x=None, y=None, z=None):
        self._sync_stream()
        check(self.lib.gpx_vec_op(self.h, op, n, float(a), self._p(x), self._p(y), self._p(z), self._p(out)), "gpx_vec_op")
        return out

    def dot(self, x, y, n=None) -> float:
        self._sync_stream()
        n = x.numel() if n is None else n
        out = self.empty(1)
        check(self.lib.gpx_dot(self.h, n, self._p(x), self._p(y), self._p(out)), "gpx_dot")
        return float(out.item())

    def diag(self, A, n=None):
        self._sync_stream()
        n = A.shape[0] if n is None else n
        out = self.empty(n)
        check(self.lib.gpx_copy_strided(self.h, n, self._p(A), A.stride(0) + 1, self._p(out), 1), "gpx_copy_strided")
        return out

    # ------------------------------------------------------------------ fit / LML / gradient
    def fit(self, kind: int, X, y, theta, s: float, with_grad: bool = False) -> GPFit:
        """K = cov(X,X;theta) + s I = L L^T, alpha = K^-1 y, LML [and dLML/dtheta]."""
        self._sync_stream()
        Xd = self.to_device(X)
        yd = self.to_device(np.asarray(y).reshape(-1) if not hasattr(y, "data_ptr") else y.reshape(-1))
        n, D = Xd.shape
        npad = padded(n)
        th, thp = _theta_array(theta)
        A = self.empty(npad, npad)
        dinv = self.empty(npad // GPX_TILE, GPX_TILE, GPX_TILE)
        alpha = self.empty(npad)
        out = self.empty(3 + 11)
        Kinv = None
        if with_grad:
            Kinv = self.empty(npad, npad)
            st = self.lib.gpx_gp_fit_grad(self.h, kind, self._p(Xd), n, D, thp, len(th), float(s), self._p(yd), self._p(A),
                                          npad, A.stride(0), self._p(dinv), self._p(Kinv), self._p(alpha), self._p(out),
                                          ctypes.c_void_p(out.data_ptr() + 24))
        else:
            st = self.lib.gpx_gp_fit(self.h, kind, self._p(Xd), n, D, thp, len(th), float(s), self._p(yd), self._p(A), npad,
                                     A.stride(0), self._p(dinv), self._p(alpha), self._p(out))
        check(st, "gpx_gp_fit")
        o = self.to_host(out)
        return GPFit(kind, th, float(s), n, npad, Xd, yd, A, dinv, alpha, float(o[0]), float(o[1]), float(o[2]),
                     grad=o[3:3 + len(th)].copy() if with_grad else None, Kinv=Kinv, L_is_inverse=with_grad)

    def ascend(self, kind: int, X, y, theta0, mask, s: float, step: float, tol: float, max_iter: int, use_graph: bool = True,
               ws=None):
        """Device-resident gradient ascent on the LML over the masked hyper-parameters (gpx_gp_ascent; tune...:121-153
        generalised).  Returns dict(theta, theta_used, iterations, lml, error, converged, history)."""
        self._sync_stream()
        Xd = self.to_device(X)
        yd = self.to_device(np.asarray(y).reshape(-1) if not hasattr(y, "data_ptr") else y.reshape(-1))
        n, D = Xd.shape
        th = np.ascontiguousarray(np.asarray(theta0, dtype=np.float64).reshape(-1)).copy()
        mk = np.ascontiguousarray(np.asarray(mask, dtype=np.int32).reshape(-1))
        if mk.size != th.size:
            raise ValueError("mask must have one entry per hyper-parameter")
        if ws is None:
            ws = self.empty(int(self.lib.gpx_gp_ascent_ws_elems(n)))
        used = np.empty_like(th)
        out4 = np.zeros(4)
        hist = np.zeros(int(max_iter))
        st = self.lib.gpx_gp_ascent(self.h, kind, self._p(Xd), n, D, ctypes.c_void_p(th.ctypes.data), th.size,
                                    ctypes.c_void_p(mk.ctypes.data), float(s), self._p(yd), float(step), float(tol), int(max_iter),
                                    int(bool(use_graph)), self._p(ws), ctypes.c_void_p(used.ctypes.data),
                                    ctypes.c_void_p(out4.ctypes.data), ctypes.c_void_p(hist.ctypes.data))
        check(st, "gpx_gp_ascent")
        it = int(out4[0])
        return dict(theta=th, theta_used=used, iterations=it, lml=float(out4[1]), error=float(out4[2]), converged=bool(out4[3]),
                    history=hist[:it].copy())

    def lml_grad(self, kind: int, Xd, theta, Kinv, alpha, n=None) -> np.ndarray:
        """.5 sum_ik (alpha_i alpha_k - Kinv_ik) dK_ik/dtheta_j for every theta_j (fused kernel) -> host array."""
        self._sync_stream()
        n = Xd.shape[0] if n is None else n
        th, thp = _theta_array(theta)
        grad = self.empty(len(th))
        check(self.lib.gpx_lml_grad(self.h, kind, self._p(Xd), n, Xd.shape[1], thp, len(th), self._p(Kinv), Kinv.stride(0),
                                    self._p(alpha), self._p(grad)), "gpx_lml_grad")
        return self.to_host(grad)

    def inverse_from_factor(self, fit: "GPFit"):
        """K_y^-1 (lower triangle valid) from a fit: L <- L^-1 in place, then Linv^T Linv (tune...:144)."""
        work = self.empty(fit.npad, fit.npad)
        self.trtri(fit.L, fit.dinv, work)
        fit.L_is_inverse = True
        fit.Kinv = self.lauum(fit.L, work)
        return fit.Kinv

    # ------------------------------------------------------------------ fused small-problem posterior
    def small_max(self) -> int:
        return int(self.lib.gpx_small_max())

    def small_posterior(self, kind: int, X, y, Xs, theta, s: float, jitter: float, Z=None):
        """mu, var, f_post (or None), lml of GP_regression.py:109-156 in ONE kernel launch; host arrays in and out.

        Requires N <= small_max(); with Z (the caller's standard normals, (n, num_fun)) also n <= small_max(),
        without Z any number of test points (one thread block per small_max() of them)."""
        self._sync_stream()
        X = np.ascontiguousarray(X, dtype=np.float64)
        Xs = np.ascontiguousarray(Xs, dtype=np.float64)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        N, D = X.shape
        n = Xs.shape[0]
        th, thp = _theta_array(theta)
        mu = np.empty(n)
        var = np.empty(n)
        lml = ctypes.c_double(0.0)
        nf = 0
        fpost = None
        zp = fp = ctypes.c_void_p(0)
        if Z is not None:
            Z = np.ascontiguousarray(Z, dtype=np.float64)
            nf = Z.shape[1]
            fpost = np.empty((n, nf))
            zp, fp = ctypes.c_void_p(Z.ctypes.data), ctypes.c_void_p(fpost.ctypes.data)
        st = self.lib.gpx_gp_small_posterior_host(
            self.h, kind, ctypes.c_void_p(X.ctypes.data), N, D, ctypes.c_void_p(y.ctypes.data),
            ctypes.c_void_p(Xs.ctypes.data), n, thp, len(th), float(s), float(jitter), zp, nf,
            ctypes.c_void_p(mu.ctypes.data), ctypes.c_void_p(var.ctypes.data), fp, ctypes.byref(lml))
        check(st, "gpx_gp_small_posterior_host")
        return mu, var, fpost, float(lml.value)

    def small_fit(self, kind: int, X, y, Xs, theta, s: float, jitter: float):
        """First half of the two-step small posterior: (mu, var, lml); the sampling factor stays on the device."""
        self._sync_stream()
        X = np.ascontiguousarray(X, dtype=np.float64)
        Xs = np.ascontiguousarray(Xs, dtype=np.float64)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        N, D = X.shape
        n = Xs.shape[0]
        th, thp = _theta_array(theta)
        mu = np.empty(n)
        var = np.empty(n)
        lml = ctypes.c_double(0.0)
        st = self.lib.gpx_gp_small_fit_host(
            self.h, kind, ctypes.c_void_p(X.ctypes.data), N, D, ctypes.c_void_p(y.ctypes.data),
            ctypes.c_void_p(Xs.ctypes.data), n, thp, len(th), float(s), float(jitter),
            ctypes.c_void_p(mu.ctypes.data), ctypes.c_void_p(var.ctypes.data), ctypes.byref(lml))
        check(st, "gpx_gp_small_fit_host")
        return mu, var, float(lml.value)

    def small_prior_factor(self, kind: int, Xs, theta, s: float) -> int:
        """chol(k(Xs,Xs) + s I) in one launch, kept on the device for small_sample (GP_regression.py:71-92)."""
        self._sync_stream()
        Xs = np.ascontiguousarray(Xs, dtype=np.float64)
        th, thp = _theta_array(theta)
        check(self.lib.gpx_gp_small_prior_factor_host(self.h, kind, ctypes.c_void_p(Xs.ctypes.data), Xs.shape[0], Xs.shape[1],
                                                      thp, len(th), float(s)), "gpx_gp_small_prior_factor_host")
        return Xs.shape[0]

    def small_sample(self, n: int, Z) -> np.ndarray:
        """Second half: mu + L_ Z for the factor kept by the last small_fit (Z: (n, num_fun) standard normals)."""
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        fpost = np.empty(Z.shape)
        check(self.lib.gpx_gp_small_sample_host(self.h, n, Z.shape[1], ctypes.c_void_p(Z.ctypes.data),
                                                ctypes.c_void_p(fpost.ctypes.data)), "gpx_gp_small_sample_host")
        return fpost

    def small_lml_grad(self, kind: int, X, y, theta, s: float, with_grad: bool = True):
        """(lml, grad or None) for N <= small_max() in one kernel launch (host arrays in, host scalars out)."""
        self._sync_stream()
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        th, thp = _theta_array(theta)
        lml = ctypes.c_double(0.0)
        grad = np.empty(len(th)) if with_grad else None
        st = self.lib.gpx_gp_small_lml_grad_host(
            self.h, kind, ctypes.c_void_p(X.ctypes.data), X.shape[0], X.shape[1], ctypes.c_void_p(y.ctypes.data), thp,
            len(th), float(s), ctypes.byref(lml), ctypes.c_void_p(grad.ctypes.data) if with_grad else None)
        check(st, "gpx_gp_small_lml_grad_host")
        return float(lml.value), grad

    def small_ascent(self, X, y, sigma: float, l0: float, s: float, step: float, tol: float, max_iter: int):
        """tune_hyperparms_regression.py:121-153 in one launch -> dict(iterations, l, l_used, lml, error, converged)."""
        self._sync_stream()
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        out = np.empty(6)
        st = self.lib.gpx_gp_small_ascent_host(
            self.h, ctypes.c_void_p(X.ctypes.data), X.shape[0], X.shape[1], ctypes.c_void_p(y.ctypes.data), float(sigma),
            float(l0), float(s), float(step), float(tol), int(max_iter), ctypes.c_void_p(out.ctypes.data))
        check(st, "gpx_gp_small_ascent_host")
        return dict(iterations=int(out[0]), l=float(out[1]), l_used=float(out[2]), lml=float(out[3]), error=float(out[4]),
                    converged=bool(out[5]))

    # ------------------------------------------------------------------ A6 prediction
    def predict(self, fit: GPFit, Xs, want_v: bool = False, kss_diag=None, m_total: Optional[int] = None, row0: int = 0):
        """mu = K_s^T alpha, var = diag(K_ss) - colsum((L^-1 K_s)^2)  (GP_regression.py:143-147).

        Returns (mu, var, V or None) as device tensors of the true test size (V is padded).  When ``Xs`` is one shard of
        a larger test set, ``m_total`` is the size of the whole set and ``row0`` the shard's first index in it: the CO2
        kernel's "square block" delta (CO2_example.py:58-66) is decided from the WHOLE block's shape and lands on the
        global diagonal."""
        self._sync_stream()
        if fit.L_is_inverse:
            raise GpxError("predict() needs the factor L; this fit holds L^-1 (fit_grad overwrote it)")
        Xsd = self.to_device(Xs)
        m = Xsd.shape[0]
        # CO2_example.py:58-66 adds theta_11^2 * I to ANY square block, so also to K_s when N == n
        square_co2 = fit.kind == COV_CO2 and fit.n == (m if m_total is None else int(m_total))
        if square_co2 and m_total is not None and (m != m_total or row0 != 0):
            # a shard of a square cross block: build without the delta, then add theta_11^2 on the global diagonal
            Ks = self.cov(fit.kind, fit.X, Xsd, fit.theta, n1p=fit.npad)
            cnt = max(0, min(m, fit.n - row0))
            if cnt > 0:
                d = self.vec_op(4, cnt, self.empty(cnt), a=float(fit.theta[10]) ** 2)
                view = Ks[row0:row0 + cnt, :cnt]
                view.diagonal().add_(d)      # torch as buffer arithmetic on cnt entries (host-side logic of a rare corner case)
        else:
            Ks = self.cov(fit.kind, fit.X, Xsd, fit.theta, delta=square_co2, n1p=fit.npad)    # (npad, mpad)
        mu = self.empty(m)
        var = self.empty(m)
        lib = self.lib
        check(lib.gpx_predict_moments(self.h, self._p(Ks), None, fit.n, m, Ks.stride(0), self._p(fit.alpha), None,
                                      self._p(mu), None), "gpx_predict_moments(mu)")
        if getattr(fit, "big", None) is None:
            fit.big = self.block_inverses(fit.L, fit.dinv)
        if fit.big[0] is not None:
            self.trsm_big(fit.L, fit.big[0], fit.big[1], Ks, trans=False)    # Ks <- V = L^-1 K_s (short chain)
        else:
            self.trsm(fit.L, fit.dinv, Ks, trans=False)                      # Ks <- V = L^-1 K_s
        if kss_diag is None:
            Kss = self.cov(fit.kind, Xsd, Xsd, fit.theta, same_x=True)
            kss_diag = self.diag(Kss, m)
        check(lib.gpx_predict_moments(self.h, None, self._p(Ks), fit.n, m, Ks.stride(0), None, self._p(kss_diag), None,
                                      self._p(var)), "gpx_predict_moments(var)")
        return mu, var, (Ks if want_v else None)

    def posterior_sample_factor(self, kind: int, theta, Xs_dev, V, jitter: float = 1e-6):
        """L_ = chol(K_ss + jitter I - V^T V)  (GP_regression.py:154).  Returns (L_ padded, K_ss diag)."""
        m = Xs_dev.shape[0]
        Kss = self.cov(kind, Xs_dev, Xs_dev, theta, same_x=True)
        kss_diag = self.diag(Kss, m)
        # add jitter on the true diagonal only: rebuild with diag_add (padding stays identity)
        C = self.cov(kind, Xs_dev, Xs_dev, theta, diag_add=jitter, same_x=True)
        mp = C.shape[0]
        self.gemm(V, V, C, a_kmajor=False, b_kmajor=False, M=mp, N=mp, K=V.shape[0], alpha=-1.0, beta=1.0)
        self.potrf(C)
        return C, kss_diag

    def tri_times(self, Lfac, Z_host: np.ndarray, n: int) -> np.ndarray:
        """(L_ z) for host normals z (n, num_fun): GEMM on device, result on host."""
        mp = Lfac.shape[0]
        nf = Z_host.shape[1]
        Zp = self.zeros(mp, padded(nf))
        Zp[:n, :nf] = self.to_device(Z_host)
        out = self.empty(mp, padded(nf))
        self.gemm(Lfac, Zp, out, a_kmajor=True, b_kmajor=False, M=mp, N=padded(nf), K=mp)
        return self.to_host(out[:n, :nf])

    # ------------------------------------------------------------------ multi-GPU (one process per GPU)
    def mg_init(self):
        """Create libgpx's own NCCL communicator over the ranks of the initialised torch.distributed group
        (the 128-byte NCCL unique id travels through torch.distributed; NCCL itself is dlopen'ed by libgpx)."""
        import torch.distributed as dist
        if getattr(self, "_mg_ready", False):
            return
        world, rank = dist.get_world_size(), dist.get_rank()
        if world > 1:
            import nvidia.nccl as _n  # torch-bundled libnccl.so.2
            path = os.path.join(os.path.dirname(_n.__file__) if getattr(_n, "__file__", None) else list(_n.__path__)[0],
                                "lib", "libnccl.so.2")
            check(self.lib.gpx_nccl_load(path.encode()), "gpx_nccl_load")
            ident = (ctypes.c_char * 128)()
            if rank == 0:
                check(self.lib.gpx_nccl_unique_id(ident), "gpx_nccl_unique_id")
            t = self.torch.tensor(list(ident.raw), dtype=self.torch.uint8, device=self.device)
            dist.broadcast(t, src=0)
            raw = bytes(t.cpu().tolist())
            buf = (ctypes.c_char * 128).from_buffer_copy(raw)
            check(self.lib.gpx_nccl_init(self.h, buf, rank, world), "gpx_nccl_init")
        self._mg_ready = True
        self._mg_world, self._mg_rank = world, rank

    def mg_fit_grad(self, kind: int, X, y, theta, s: float, nb: int = 512, with_grad: bool = True, ws=None):
        """Distributed fit + LML (+ gradient) over the ranks of ``mg_init`` -> (lml, grad ndarray or None, alpha dev)."""
        self._sync_stream()
        world = getattr(self, "_mg_world", 1)
        Xd, yd = self.to_device(X), self.to_device(np.asarray(y).reshape(-1) if not hasattr(y, "data_ptr") else y.reshape(-1))
        n, D = Xd.shape
        th, thp = _theta_array(theta)
        npad = int(self.lib.gpx_mg_padded_dim(n, nb, world))
        if ws is None:
            ws = self.empty(int(self.lib.gpx_mg_workspace_elems(n, nb, world)))
        alpha = self.empty(npad)
        out = self.empty(3 + 11)
        st = self.lib.gpx_mg_fit_grad(self.h, kind, self._p(Xd), n, D, thp, len(th), float(s), self._p(yd), nb, self._p(ws),
                                      self._p(alpha), self._p(out), ctypes.c_void_p(out.data_ptr() + 24), int(with_grad))
        check(st, "gpx_mg_fit_grad")
        o = self.to_host(out)
        return float(o[0]), (o[3:3 + len(th)].copy() if with_grad else None), alpha

    def mg_layout(self, n: int, nb: int):
        """dict(Aloc, Kloc, Lfull, dinv (element offsets into the workspace), npad, wloc)."""
        world = getattr(self, "_mg_world", 1)
        out = (ctypes.c_int64 * 6)()
        check(self.lib.gpx_mg_workspace_layout(n, nb, world, out), "gpx_mg_workspace_layout")
        return dict(Aloc=out[0], Kloc=out[1], Lfull=out[2], dinv=out[3], npad=out[4], wloc=out[5])

    def mg_workspace(self, n: int, nb: int):
        world = getattr(self, "_mg_world", 1)
        return self.empty(int(self.lib.gpx_mg_workspace_elems(n, nb, world)))

    def mg_fit(self, kind: int, X, y, theta, s: float, nb: int = 256, ws=None) -> GPFit:
        """Distributed fit (no gradient) -> a GPFit whose factor is the REPLICATED factor inside the workspace: every
        rank can predict its own shard of test points from it without refitting (``mg_predict``)."""
        self._sync_stream()
        Xd = self.to_device(X)
        yd = self.to_device(np.asarray(y).reshape(-1) if not hasattr(y, "data_ptr") else y.reshape(-1))
        n, D = Xd.shape
        th, thp = _theta_array(theta)
        ws = self.mg_workspace(n, nb) if ws is None else ws
        lay = self.mg_layout(n, nb)
        npad = lay["npad"]
        alpha = self.empty(npad)
        out = self.empty(3 + 11)
        check(self.lib.gpx_mg_fit_grad(self.h, kind, self._p(Xd), n, D, thp, len(th), float(s), self._p(yd), nb, self._p(ws),
                                       self._p(alpha), self._p(out), ctypes.c_void_p(out.data_ptr() + 24), 0), "gpx_mg_fit_grad")
        o = self.to_host(out)
        L = ws[lay["Lfull"]:lay["Lfull"] + npad * npad].view(npad, npad)
        dinv = ws[lay["dinv"]:lay["dinv"] + npad * GPX_TILE].view(npad // GPX_TILE, GPX_TILE, GPX_TILE)
        fit = GPFit(kind, th, float(s), n, npad, Xd, yd, L, dinv, alpha, float(o[0]), float(o[1]), float(o[2]))
        fit.ws = ws          # keeps the workspace alive
        return fit

    def mg_predict(self, fit: GPFit, Xs):
        """Test-point prediction split over the ranks: every rank predicts a contiguous shard of ``Xs`` from
        the replicated factor of ``mg_fit`` and the (mu, var) slices are all-gathered.  Host arrays out."""
        from . import parallel as P
        import torch.distributed as dist
        Xs = np.asarray(Xs, dtype=np.float64)
        m = Xs.shape[0]
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        lo, hi = P.shard_range(m, rank, world)
        if hi > lo:
            mu, var, _ = self.predict(fit, Xs[lo:hi], m_total=m, row0=lo)
            loc = np.stack([self.to_host(mu), self.to_host(var)], axis=1)
        else:
            loc = np.zeros((0, 2))
        full = P.gather_slices(loc, m) if world > 1 else loc
        return full[:, 0].copy(), full[:, 1].copy()

    def mg_factor(self, kind: int, Xd, theta, diag_add: float, scale, nb: int, ws):
        """Distributed Cholesky of diag(scale) k(X,X) diag(scale) + diag_add I into the workspace's replicated factor."""
        self._sync_stream()
        th, thp = _theta_array(theta)
        check(self.lib.gpx_mg_factor(self.h, kind, self._p(Xd), Xd.shape[0], Xd.shape[1], thp, len(th), float(diag_add),
                                     self._p(scale), nb, self._p(ws)), "gpx_mg_factor")

    def mg_emulate_fit_grad(self, P: int, kind: int, X, y, theta, s: float, nb: int = 256):
        """P virtual ranks on this one GPU (test helper for the block-cyclic index maps)."""
        self._sync_stream()
        Xd, yd = self.to_device(X), self.to_device(np.asarray(y).reshape(-1))
        n, D = Xd.shape
        th, thp = _theta_array(theta)
        npad = int(self.lib.gpx_mg_padded_dim(n, nb, P))
        ws = self.empty(P * int(self.lib.gpx_mg_workspace_elems(n, nb, P)))
        alpha = self.empty(npad)
        out = self.empty(3 + 11)
        st = self.lib.gpx_mg_emulate_fit_grad(self.h, P, kind, self._p(Xd), n, D, thp, len(th), float(s), self._p(yd), nb,
                                              self._p(ws), self._p(alpha), self._p(out), ctypes.c_void_p(out.data_ptr() + 24))
        check(st, "gpx_mg_emulate_fit_grad")
        o = self.to_host(out)
        return float(o[0]), o[3:3 + len(th)].copy(), self.to_host(alpha[:n])

    # ------------------------------------------------------------------ measurement
    def fp64_peak(self, dmma: bool = True, iters: int = 4096):
        tf = ctypes.c_double()
        ms = ctypes.c_double()
        check(self.lib.gpx_bench_fp64_peak(self.h, int(dmma), iters, ctypes.byref(tf), ctypes.byref(ms)), "gpx_bench_fp64_peak")
        return tf.value, ms.value


_engines = {}


def new_engine(device: int = 0) -> Engine:
    """A fresh handle on the same device (own scratch and helper streams): independent problems issued through different
    engines, each under its own ``torch.cuda.stream``, run concurrently."""
    return Engine(device)


def get_engine(device: int = 0) -> Engine:
    """Process-wide engine per device (handles are cheap but workspaces are not)."""
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]
