"""Size-independent properties of the path, checked at large sizes where the CPU oracle cannot follow
(BASELINE.json configs C3/C5 sizes):
  * (K + sI) alpha = y            -- residual through an independent row-panel K build x alpha product
  * L L^T = K + sI                -- on random probe vectors
  * K^-1 (K + sI) v = v           -- the explicit inverse used by the gradient
  * dLML/dtheta                   -- against central finite differences of the engine's own LML
  * symmetry / positive diagonal of K, invariance of the LML to a translation of the inputs."""
import numpy as np
import pytest

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from gaussian_process_b200 import get_engine
    return get_engine(0)


def _residual(eng, kind, Xd, theta, s, alpha, y, n, panel=4096):
    """max |(K + sI) alpha - y| / max |y| with K rebuilt in row panels (never the factorised copy)."""
    worst = 0.0
    for r0 in range(0, n, panel):
        r1 = min(n, r0 + panel)
        Kp = eng.cov(kind, Xd[r0:r1].contiguous(), Xd, theta)          # (rows_pad, npad), no noise term
        out = eng.gemv(Kp, alpha, eng.zeros(Kp.shape[0]), m=r1 - r0, n=n)
        res = eng.to_host(out[:r1 - r0]) + s * eng.to_host(alpha[r0:r1]) - y[r0:r1]
        worst = max(worst, float(np.max(np.abs(res))))
    return worst / float(np.max(np.abs(y)))


@pytest.mark.parametrize("N", [16384, 65536])
def test_fit_residual_at_full_size(eng, N):
    from gaussian_process_b200._lib import COV_SE
    import torch
    X, y = O.synth_c5(N, 16)
    theta, s = [1.0, 4.0], 5e-4
    fit = eng.fit(COV_SE, X, y, theta, s)
    assert np.isfinite(fit.lml)
    r = _residual(eng, COV_SE, fit.X, theta, s, fit.alpha, y, N)
    # cond(K + sI) ~ N / s ~ 1e8: a backward-stable solve leaves a residual ~ eps * cond-independent factor
    assert r < 1e-8, r
    # log-determinant sanity: N log(s) < log det < N log(sigma^2 N + s)
    assert N * np.log(s) / 2 < fit.sum_log_diag < N * np.log(N + s) / 2
    del fit
    torch.cuda.empty_cache()


def test_factor_and_inverse_consistency_16k(eng):
    from gaussian_process_b200._lib import COV_SE
    import torch
    N = 16384
    X, y = O.synth_c5(N, 16)
    theta, s = [1.0, 4.0], 5e-4
    Xd = eng.to_device(X)
    K = eng.cov(COV_SE, Xd, Xd, theta, diag_add=s, same_x=True)      # full symmetric K + sI
    A = K.clone()
    dinv = eng.potrf(A)
    rs = np.random.RandomState(0)
    v = eng.to_device(rs.randn(N))
    # L (L^T v) == (K + sI) v
    t = eng.gemv(A, v, eng.zeros(N), trans=True)
    llt = eng.gemv(A, t, eng.zeros(N))
    kv = eng.gemv(K, v, eng.zeros(N))
    num = eng.to_host(llt) - eng.to_host(kv)
    assert np.max(np.abs(num)) / np.max(np.abs(eng.to_host(kv))) < 1e-12
    # K^-1 (lower, from trtri + lauum) times (K + sI) v == v
    eng.trtri(A, dinv)
    Kinv = eng.lauum(A)
    out = eng.symv_lower(Kinv, kv, eng.zeros(N))
    assert np.max(np.abs(eng.to_host(out) - eng.to_host(v))) / np.max(np.abs(eng.to_host(v))) < 1e-6   # cond ~ 3e7
    del K, A, Kinv
    torch.cuda.empty_cache()


def test_gradient_matches_finite_differences_of_engine_lml(eng):
    from gaussian_process_b200._lib import COV_CO2, COV_SE
    X, y = O.synth_c5(4096, 16)
    th = np.array([1.0, 4.0])
    fit = eng.fit(COV_SE, X, y, th, 5e-4, with_grad=True)
    for j in range(2):
        h = 1e-4 * th[j]
        e = np.zeros(2)
        e[j] = h
        f = lambda t: eng.fit(COV_SE, X, y, t, 5e-4).lml   # noqa: E731
        fd = (8 * (f(th + e) - f(th - e)) - (f(th + 2 * e) - f(th - 2 * e))) / (12 * h)
        assert abs(fd - fit.grad[j]) <= 1e-5 * abs(fit.grad[j]), (j, fd, fit.grad[j])
    Xc, yc, _ = O.synth_c2(2048)
    thc = O.CO2_THETA_BOOK.copy()
    fc = eng.fit(COV_CO2, Xc, yc, thc, 1.0, with_grad=True)          # s = 1 keeps the FD well conditioned
    for j in (1, 4, 7, 10):
        h = 1e-3 * thc[j]
        e = np.zeros(11)
        e[j] = h
        f = lambda t: eng.fit(COV_CO2, Xc, yc, t, 1.0).lml   # noqa: E731
        fd = (8 * (f(thc + e) - f(thc - e)) - (f(thc + 2 * e) - f(thc - 2 * e))) / (12 * h)
        assert abs(fd - fc.grad[j]) <= 1e-5 * max(1.0, abs(fc.grad[j])), (j, fd, fc.grad[j])


def test_covariance_symmetry_and_translation_invariance(eng):
    from gaussian_process_b200._lib import COV_CO2, COV_SE
    rs = np.random.RandomState(3)
    X = rs.randn(1000, 8)
    Xd = eng.to_device(X)
    K = eng.to_host(eng.cov(COV_SE, Xd, Xd, [1.2, 0.8], same_x=True))[:1000, :1000]
    assert np.array_equal(K, K.T) and np.all(np.diag(K) == 1.2 ** 2)
    y = np.sin(X.sum(1))
    a = eng.fit(COV_SE, X, y, [1.2, 0.8], 5e-4).lml
    b = eng.fit(COV_SE, X + 1000.0, y, [1.2, 0.8], 5e-4).lml      # direct differences: no cancellation in the Gram term
    assert abs(a - b) <= 1e-9 * abs(a)
    t = 1958 + np.arange(600) / 12.0
    yc = np.sin(t)
    c = eng.fit(COV_CO2, t[:, None], yc, O.CO2_THETA_BOOK, 5e-4).lml
    d = eng.fit(COV_CO2, (t - 1958.0)[:, None], yc, O.CO2_THETA_BOOK, 5e-4).lml
    assert abs(c - d) <= 1e-8 * abs(c)
