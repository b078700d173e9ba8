"""GPU parity tests of the libgpx primitives (through the C ABI) against NumPy on seeded inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def eng():
    from gaussian_process_b200 import get_engine
    return get_engine(0)


def spd(n, seed=0, cond_shift=1.0):
    rs = np.random.RandomState(seed)
    A = rs.randn(n, n)
    return A @ A.T / n + cond_shift * np.eye(n)


@pytest.mark.parametrize("akm,bkm", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_gemm_all_layouts(eng, akm, bkm):
    rs = np.random.RandomState(1)
    M, N, K = 256, 384, 272
    A = rs.randn(M, K)
    B = rs.randn(K, N)
    C0 = rs.randn(M, N)
    Ad = eng.to_device(A if akm else A.T.copy())
    Bd = eng.to_device(B.T.copy() if bkm else B)
    Cd = eng.to_device(C0)
    eng.gemm(Ad, Bd, Cd, bool(akm), bool(bkm), M, N, K, alpha=-1.5, beta=0.5)
    ref = -1.5 * A @ B + 0.5 * C0
    assert rel(eng.to_host(Cd), ref) < 1e-13


@pytest.mark.parametrize("n", [128, 256, 384, 1024, 1152])
def test_potrf_matches_numpy(eng, n):
    A = spd(n, seed=n)
    Ad = eng.to_device(A)
    dinv = eng.potrf(Ad)
    L = np.linalg.cholesky(A)
    Lg = eng.to_host(Ad)
    assert np.all(np.triu(Lg, 1) == 0.0)
    assert rel(Lg, L) < 1e-12
    # leaf inverses
    D = eng.to_host(dinv)
    for b in range(n // 128):
        blk = L[b * 128:(b + 1) * 128, b * 128:(b + 1) * 128]
        assert rel(D[b] @ blk, np.eye(128)) < 1e-11


@pytest.mark.parametrize("n,group", [(2432, 0), (4224, 0), (5248, 0), (2048, 4), (3200, 3)])
def test_potrf_lookahead_groups_match_numpy(eng, n, group):
    """The three-stream look-ahead factorisation (substitution chain, grouped K = G*128 bulk updates): panel counts that
    leave a short last group (19, 33, 41 panels) with the default group sizes, and forced group sizes that do not divide
    the panel count.  Factor, strict-upper zeros and the batched leaf inverses against NumPy."""
    A = spd(n, seed=n)
    Ad = eng.to_device(A)
    assert eng.lib.gpx_potrf_set_group(group) == 0
    try:
        dinv = eng.potrf(Ad)
    finally:
        eng.lib.gpx_potrf_set_group(0)
    L = np.linalg.cholesky(A)
    Lg = eng.to_host(Ad)
    assert np.all(np.triu(Lg, 1) == 0.0)
    assert rel(Lg, L) < 1e-12
    D = eng.to_host(dinv)
    for b in (0, 1, n // 256, n // 128 - 1):
        blk = L[b * 128:(b + 1) * 128, b * 128:(b + 1) * 128]
        assert rel(D[b] @ blk, np.eye(128)) < 1e-11


@pytest.mark.parametrize("n,bad", [(2048, 1500), (4224, 4200), (1152, 5)])
def test_potrf_lookahead_not_positive_definite_reports_the_pivot(eng, n, bad):
    """A failing pivot inside the look-ahead range (factor-only leaves on the chain stream): LinAlgError, as
    np.linalg.cholesky raises, and the handle stays usable afterwards."""
    A = spd(n, seed=7)
    A[bad, bad] = -5.0
    with pytest.raises(np.linalg.LinAlgError):
        eng.potrf(eng.to_device(A))
    B = spd(256, seed=8)
    Bd = eng.to_device(B)
    eng.potrf(Bd)
    assert rel(eng.to_host(Bd), np.linalg.cholesky(B)) < 1e-12


def test_potrf_not_positive_definite_raises(eng):
    A = spd(256, seed=3)
    A[200, 200] = -5.0
    with pytest.raises(np.linalg.LinAlgError):
        eng.potrf(eng.to_device(A))


@pytest.mark.parametrize("n", [128, 640])
def test_trsv_trsm_trtri_lauum(eng, n):
    A = spd(n, seed=7 + n)
    L = np.linalg.cholesky(A)
    Ad = eng.to_device(A)
    dinv = eng.potrf(Ad)
    rs = np.random.RandomState(2)
    b = rs.randn(n)
    x = eng.to_device(b.copy())
    eng.trsv(Ad, dinv, x, trans=False)
    assert rel(eng.to_host(x), np.linalg.solve(L, b)) < 1e-11
    x = eng.to_device(b.copy())
    eng.trsv(Ad, dinv, x, trans=True)
    assert rel(eng.to_host(x), np.linalg.solve(L.T, b)) < 1e-11
    Bm = rs.randn(n, 256)
    Xd = eng.to_device(Bm.copy())
    eng.trsm(Ad, dinv, Xd, trans=False)
    assert rel(eng.to_host(Xd), np.linalg.solve(L, Bm)) < 1e-11
    Xd = eng.to_device(Bm.copy())
    eng.trsm(Ad, dinv, Xd, trans=True)
    assert rel(eng.to_host(Xd), np.linalg.solve(L.T, Bm)) < 1e-11
    Li = Ad.clone()
    eng.trtri(Li, dinv)
    Linv = np.linalg.inv(L)
    assert rel(eng.to_host(Li), Linv) < 1e-11
    Kinv = eng.to_host(eng.lauum(Li))
    ref = Linv.T @ Linv
    assert rel(np.tril(Kinv), np.tril(ref)) < 1e-11


def test_gemv_symv_dot(eng):
    rs = np.random.RandomState(4)
    m, n = 777, 513
    A = rs.randn(m, n)
    x, xt = rs.randn(n), rs.randn(m)
    y0, yt0 = rs.randn(m), rs.randn(n)
    Ad = eng.to_device(A)
    y = eng.gemv(Ad, eng.to_device(x), eng.to_device(y0.copy()), alpha=0.7, beta=-0.3)
    assert rel(eng.to_host(y), 0.7 * A @ x - 0.3 * y0) < 1e-13
    yt = eng.gemv(Ad, eng.to_device(xt), eng.to_device(yt0.copy()), trans=True, alpha=1.1, beta=2.0)
    assert rel(eng.to_host(yt), 1.1 * A.T @ xt + 2.0 * yt0) < 1e-13
    S = spd(300, seed=9)
    xs = rs.randn(300)
    ys = eng.symv_lower(eng.to_device(np.tril(S)), eng.to_device(xs), eng.zeros(300))
    assert rel(eng.to_host(ys), S @ xs) < 1e-13
    assert abs(eng.dot(eng.to_device(x), eng.to_device(x)) - x @ x) < 1e-10


def test_cov_build_all_kinds_and_padding(eng):
    from gaussian_process_b200._lib import COV_CO2, COV_LIN, COV_PER, COV_SE
    from oracle import gp_oracle as O
    rs = np.random.RandomState(5)
    for D in (1, 3, 16, 40):
        a, b = rs.randn(150, D), rs.randn(70, D)
        ad, bd = eng.to_device(a), eng.to_device(b)
        K = eng.to_host(eng.cov(COV_SE, ad, bd, [1.3, 0.9]))
        assert K.shape == (256, 128)
        assert rel(K[:150, :70], O.rbf_kernel(a, b, 1.3, 0.9)) < 1e-14
        assert np.all(K[150:, :] == 0) and np.all(K[:, 70:] == 0)
        assert rel(eng.to_host(eng.cov(COV_LIN, ad, bd, [0.4]))[:150, :70], O.lin_kernel(a, b, 0.4)) < 1e-13
        th = O.CO2_THETA_BOOK * (0.7 + 0.6 * rs.rand(11))
        assert rel(eng.to_host(eng.cov(COV_CO2, ad, bd, th))[:150, :70], O.co2_covariance(a, b, th)) < 1e-13
        Kaa = eng.to_host(eng.cov(COV_CO2, ad, ad, th, diag_add=0.25, same_x=True))
        assert rel(Kaa[:150, :150], O.co2_covariance(a, a, th) + 0.25 * np.eye(150)) < 1e-13
        assert np.array_equal(Kaa[150:, 150:], np.eye(106))
        Klow = eng.to_host(eng.cov(COV_SE, ad, ad, [1.0, 2.0], same_x=True, lower=True))
        assert np.all(Klow[:128, 128:] == 0)
        assert rel(Klow[128:150, :150], O.rbf_kernel(a, a, 1.0, 2.0)[128:150]) < 1e-14
    a1, b1 = rs.randn(90, 1), rs.randn(33, 1)
    Kp = eng.to_host(eng.cov(COV_PER, eng.to_device(a1), eng.to_device(b1), [2.0, 1.5]))[:90, :33]
    assert rel(Kp, O.per_kernel(a1, b1, [2.0, 1.5])) < 1e-13


def test_cov_derivatives_same_pass(eng):
    """dK/dtheta written in the same pass equal the analytic matrices of the oracle."""
    from gaussian_process_b200._lib import COV_CO2, COV_SE
    from oracle import gp_oracle as O
    X, y, _ = O.synth_c2(140)
    th = O.CO2_THETA_BOOK
    Xd = eng.to_device(X)
    K, dK = eng.cov(COV_CO2, Xd, Xd, th, same_x=True, with_grad=True)
    ref = O.co2_dcov(X, th)
    dK = eng.to_host(dK)
    for j in range(11):
        assert rel(dK[j, :140, :140], ref[j]) < 1e-12, j
    Xr = np.random.RandomState(0).randn(100, 5)
    _, dS = eng.cov(COV_SE, eng.to_device(Xr), eng.to_device(Xr), [1.4, 2.2], same_x=True, with_grad=True)
    rs = O.rbf_dcov(Xr, 1.4, 2.2)
    dS = eng.to_host(dS)
    assert rel(dS[0, :100, :100], rs[0]) < 1e-13 and rel(dS[1, :100, :100], rs[1]) < 1e-13


def test_fp64_peak_microbench_runs(eng):
    tf_dmma, _ = eng.fp64_peak(True, 512)
    tf_dfma, _ = eng.fp64_peak(False, 512)
    assert tf_dmma > 1.0 and tf_dfma > 1.0


@pytest.mark.parametrize("n", [2048, 2560, 3072])
def test_block_inverses_and_short_chain_solves(eng, n):
    A = spd(n, seed=n + 1)
    L = np.linalg.cholesky(A)
    Ad = eng.to_device(A)
    dinv = eng.potrf(Ad)
    D, bs = eng.block_inverses(Ad, dinv)
    assert D is not None and n % bs == 0
    Dh = eng.to_host(D)
    for b in range(n // bs):
        blk = L[b * bs:(b + 1) * bs, b * bs:(b + 1) * bs]
        assert rel(Dh[b] @ blk, np.eye(bs)) < 1e-10
        assert np.all(np.triu(Dh[b], 1) == 0.0)
    rs = np.random.RandomState(1)
    b0 = rs.randn(n)
    for trans in (False, True):
        x = eng.to_device(b0.copy())
        eng.trsv_big(Ad, D, bs, x, trans=trans)
        assert rel(eng.to_host(x), np.linalg.solve(L.T if trans else L, b0)) < 1e-10
        Bm = rs.randn(n, 256)
        Xd = eng.to_device(Bm.copy())
        eng.trsm_big(Ad, D, bs, Xd, trans=trans)
        assert rel(eng.to_host(Xd), np.linalg.solve(L.T if trans else L, Bm)) < 1e-10
