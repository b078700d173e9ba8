"""GPU parity tests of the one-launch small-problem posterior (csrc/small.cu, gpx_gp_small_posterior_host): the
as-shipped GP_regression sizes (N=5, n=100; BASELINE.json configs[0]) against the oracle, against the golden
vectors, and against the tiled path of the same library."""
import numpy as np
import pytest

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-8


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture
def G():
    from gaussian_process_b200 import GP_regression as mod
    old = mod.FUSED_SMALL_PATH
    yield mod
    mod.FUSED_SMALL_PATH = old


@pytest.mark.parametrize("fused", [True, False])
def test_c1_as_shipped_both_paths(G, golden, fused):
    """N=5 / N=64, n=100: golden vectors of the unmodified reference through either path."""
    G.FUSED_SMALL_PATH = fused
    g = golden("ka1_regression.npz")
    for tag in ("n64", "n5"):
        X, y, Xs = g[tag + "_X"], g[tag + "_y"], g[tag + "_Xs"]
        for kc, par in (("rbf", 1), ("per", [2.0, 1.5])):
            np.random.seed(7)
            mu, sd, fp = G.prediction(X, Xs, y, kc, par, 10)
            assert mu.shape == (100,) and sd.shape == (100,) and fp.shape == (100, 10)
            assert rel(mu, g["%s_%s_mu" % (tag, kc)]) < TOL
            assert rel(sd ** 2, g["%s_%s_sd" % (tag, kc)] ** 2) < TOL
            assert rel(fp, g["%s_%s_fpost" % (tag, kc)]) < 1e-5


def test_fused_launch_counts_and_matches_tiled_path(G):
    from gaussian_process_b200 import get_engine
    eng = get_engine()
    X, y, Xs = O.synth_c1(5, 100)
    G.FUSED_SMALL_PATH = True
    np.random.seed(3)
    before = eng.launches()
    mu, sd, fp = G.prediction(X, Xs, y, 'rbf', 1, 10)
    assert eng.launches() - before == 2          # the posterior kernel + mu + L_ z
    from gaussian_process_b200._lib import COV_SE
    z = np.random.RandomState(3).normal(size=(100, 10))
    before = eng.launches()
    mu1, var1, fp1, _ = eng.small_posterior(COV_SE, X, y, Xs, [1.0, 1.0], 5e-4, 1e-6, z)
    assert eng.launches() - before == 1          # normals supplied up front: ONE launch
    assert np.array_equal(mu1, mu) and rel(var1, sd ** 2) < 1e-12 and rel(fp1, fp) < 1e-12
    G.FUSED_SMALL_PATH = False
    np.random.seed(3)
    mu_t, sd_t, fp_t = G.prediction(X, Xs, y, 'rbf', 1, 10)
    assert rel(mu, mu_t) < 1e-11 and rel(sd ** 2, sd_t ** 2) < 1e-9 and rel(fp, fp_t) < 1e-5


@pytest.mark.parametrize("N,n,D,kc,par", [(1, 1, 1, 'rbf', 1.0), (2, 128, 3, 'rbf', 0.8), (128, 128, 4, 'rbf', 1.5),
                                            (128, 7, 16, 'rbf', 4.0), (37, 101, 2, 'lin', 0.5), (90, 128, 1, 'per', [2.0, 1.5])])
def test_fused_vs_oracle_shapes_and_kernels(G, N, n, D, kc, par):
    rs = np.random.RandomState(100 + N + n)
    X = rs.uniform(-3, 3, (N, D))
    y = np.sin(X.sum(1)) + 0.05 * rs.randn(N)
    Xs = rs.uniform(-3, 3, (n, D))
    G.FUSED_SMALL_PATH = True
    np.random.seed(5)
    mu, sd, fp = G.prediction(X, Xs, y, kc, par, 4)
    np.random.seed(5)
    mu_o, sd_o, fp_o = O.regression_prediction(X, Xs, y, kc, par, 4)
    assert mu.shape == (n,) and fp.shape == (n, 4)
    assert rel(mu, mu_o) < TOL
    # variance is a difference of O(1) terms: absolute agreement at the 1e-8 level of k**
    assert np.max(np.abs(sd ** 2 - sd_o ** 2)) < 1e-8 * max(1.0, np.max(np.abs(sd_o ** 2)))
    assert rel(fp, fp_o) < 1e-4


def test_fused_lml_and_many_test_points(G):
    """Without sampling the kernel runs one block per 128 test points: BO-style call (N <= 15, n = 500)."""
    from gaussian_process_b200 import get_engine
    from gaussian_process_b200._lib import COV_SE
    eng = get_engine()
    X, y, Xs = O.synth_c1(12, 500)
    mu, var, fp, lml = eng.small_posterior(COV_SE, X, y, Xs, [1.0, 1.3], 5e-4, 1e-6, None)
    assert fp is None and mu.shape == (500,)
    K = O.rbf_kernel(X, X, 1.0, 1.3)
    L, alpha = O._fit_solve(K, y, 5e-4)
    Ks = O.rbf_kernel(X, Xs, 1.0, 1.3)
    v = np.linalg.solve(L, Ks)
    assert rel(mu, Ks.T @ alpha) < TOL
    assert np.max(np.abs(var - (1.0 - np.sum(v ** 2, axis=0)))) < 1e-8
    assert rel(lml, O.lml_from(L, y, alpha)) < TOL
    with pytest.raises(Exception):
        eng.small_posterior(COV_SE, X, y, Xs, [1.0, 1.3], 5e-4, 1e-6, np.zeros((500, 1)))   # sampling needs n <= 128


def test_co2_small_prediction_and_square_block_delta(G):
    """CO2_example.make_prediction at a small size through both paths, including N == n where the reference adds
    theta_11^2 to the diagonal of the (square) cross block K_s as well (CO2_example.py:58-66)."""
    from gaussian_process_b200 import CO2_example as C2
    th = O.CO2_THETA_BOOK
    for N, n in ((100, 60), (80, 80)):
        X, y, _ = O.synth_c2(N)
        Xs = X[-1:] + (1 + np.arange(n))[:, None] / 12.0
        np.random.seed(2)
        mu_o, sd_o, fp_o = O.co2_make_prediction(X, Xs, y, th)
        for fused in (True, False):
            G.FUSED_SMALL_PATH = fused
            np.random.seed(2)
            mu, sd, fp = C2.make_prediction(X, Xs, y, th)
            assert rel(mu, mu_o) < TOL, (N, n, fused)
            assert rel(sd ** 2, sd_o ** 2) < 1e-7, (N, n, fused)
            assert rel(fp, fp_o) < 1e-4, (N, n, fused)


def test_co2_bo_posterior_both_paths(G, golden):
    """GP over the 11-D theta space (CO2_example.py:152-179): golden vectors through the fused and the tiled path."""
    from gaussian_process_b200 import CO2_example as C2
    g = golden("ka3_co2.npz")
    mu_o, sd_o = O.co2_bo_posterior(g["bo_theta_train"], g["bo_theta_test"], g["bo_y"])
    assert rel(mu_o, g["bo_mu"]) < 1e-10
    ok = np.isfinite(sd_o)
    for fused in (True, False):
        G.FUSED_SMALL_PATH = fused
        mu, sd = C2.bayesian_opt(g["bo_theta_train"], g["bo_theta_test"], g["bo_y"])
        assert rel(mu, g["bo_mu"]) < TOL, fused
        assert np.max(np.abs(sd[ok] ** 2 - sd_o[ok] ** 2)) < 1e-7 * np.max(sd_o[ok] ** 2), fused


def test_fused_not_positive_definite_raises_and_restores_rng(G):
    """Duplicated test points with zero jitter headroom: LinAlgError like GP_regression.py:154, and the global RNG
    is left where the reference leaves it (no draw happened)."""
    from gaussian_process_b200 import get_engine
    from gaussian_process_b200._lib import COV_SE
    eng = get_engine()
    X, y, _ = O.synth_c1(6, 10)
    Xs = np.zeros((20, 1))
    with pytest.raises(np.linalg.LinAlgError):
        eng.small_posterior(COV_SE, X, y, Xs, [1.0, 1.0], 5e-4, -1.0, np.zeros((20, 1)))
    with pytest.raises(np.linalg.LinAlgError):
        eng.small_posterior(COV_SE, X, y, Xs, [1.0, 1.0], -5.0, 1e-6, np.zeros((20, 1)))
    G.FUSED_SMALL_PATH = True
    old_s = G.NOISE_VARIANCE
    try:
        G.NOISE_VARIANCE = -5.0
        np.random.seed(42)
        st = np.random.get_state()
        with pytest.raises(np.linalg.LinAlgError):
            G.prediction(X, Xs, y, 'rbf', 1, 2)
        now = np.random.get_state()
        assert np.array_equal(now[1], st[1]) and now[2:] == st[2:]
    finally:
        G.NOISE_VARIANCE = old_s


# ------------------------------------------------------------------ fused LML / gradient / ascent loop (N <= 128)
@pytest.mark.parametrize("fused", [True, False])
def test_tune_first_loop_both_paths(G, golden, fused):
    """tune_hyperparms_first at N=8: same iteration count, moments and LML as the unmodified reference, whether the
    ascent loop runs inside one kernel (fused) or as one fused fit+grad call per iteration (tiled)."""
    import contextlib
    import io
    from gaussian_process_b200 import tune_hyperparms_regression as T
    G.FUSED_SMALL_PATH = fused
    g = golden("ka2_tune_first.npz")
    X, y, Xs = O.synth_c1(8, 100)
    np.random.seed(3)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        mu, sd, fp, lml = T.tune_hyperparms_first(X, Xs, y, 2, 1, np.array([1.7]))
    ref_txt = str(g["stdout"])
    assert buf.getvalue().splitlines()[0] == ref_txt.splitlines()[0]          # same iteration count
    assert rel(mu, g["mu"]) < 1e-7 and rel(sd ** 2, g["sd"] ** 2) < 1e-6 and rel(lml, g["lml"]) < TOL
    assert fp.shape == (100, 2)


def test_small_ascent_matches_oracle_loop():
    from gaussian_process_b200 import get_engine
    eng = get_engine()
    X, y, Xs = O.synth_c1(12, 20)
    np.random.seed(0)
    mu_o, sd_o, fp_o, lml_o, l_o, it_o = O.tune_first(X, Xs, y, 1, 1.0, np.array([0.6]))
    res = eng.small_ascent(X, y, 1.0, 0.6, 5e-4, 0.01, 1e-3, 10000)
    assert res["iterations"] == it_o and res["converged"]
    assert rel(res["l"], l_o) < 1e-8 and rel(res["lml"], lml_o) < TOL
    capped = eng.small_ascent(X, y, 1.0, 0.6, 5e-4, 0.01, 0.0, 7)
    assert capped["iterations"] == 7 and not capped["converged"]


@pytest.mark.parametrize("N,D", [(1, 1), (100, 3), (128, 16)])
def test_small_lml_grad_se_vs_oracle(N, D):
    from gaussian_process_b200 import get_engine
    from gaussian_process_b200._lib import COV_SE
    eng = get_engine()
    rs = np.random.RandomState(N)
    X = rs.randn(N, D)
    y = np.sin(X.sum(1)) + 0.05 * rs.randn(N)
    lml, grad = eng.small_lml_grad(COV_SE, X, y, [1.3, 2.1], 5e-4)
    K = O.rbf_kernel(X, X, 1.3, 2.1) + 5e-4 * np.eye(N)
    L = np.linalg.cholesky(K)
    alpha = np.linalg.solve(L.T, np.linalg.solve(L, y))
    Kinv = np.dot(np.linalg.inv(L.T), np.linalg.inv(L))
    assert rel(lml, O.lml_from(L, y, alpha)) < TOL
    ref = O.lml_grad_from(alpha, Kinv, O.rbf_dcov(X, 1.3, 2.1))
    assert np.all(np.abs(grad - ref) <= 1e-7 * np.maximum(1.0, np.abs(ref))), (grad, ref)
    # the reference's own GEMM + trace form for dLML/dl (tune...:54-57)
    assert abs(grad[1] - O.rbf_grad_l(X, 1.3, 2.1, alpha, Kinv)) <= 1e-7 * max(1.0, abs(grad[1]))
    lml_only, none = eng.small_lml_grad(COV_SE, X, y, [1.3, 2.1], 5e-4, with_grad=False)
    assert none is None and lml_only == lml


def test_small_co2_lml_and_gradient_both_paths(G):
    from gaussian_process_b200 import CO2_example as C2
    X, y, _ = O.synth_c2(120)
    th = O.CO2_THETA_BOOK
    K = O.co2_covariance(X, X, th) + O.S_NOISE * np.eye(120)
    Kinv = np.linalg.inv(K)
    ref = O.lml_grad_from(Kinv @ y, Kinv, O.co2_dcov(X, th))
    lml_o = O.co2_lml(X, y, th)
    for fused in (True, False):
        G.FUSED_SMALL_PATH = fused
        assert rel(C2.compute_mar_likelihood(X, y, th), lml_o) < TOL, fused
        lml, grad = C2.compute_mar_likelihood_gradient(X, y, th)
        assert rel(lml, lml_o) < TOL, fused
        assert np.all(np.abs(grad - ref) <= 1e-6 * np.maximum(1.0, np.abs(ref))), (fused, grad, ref)


def test_small_lml_not_positive_definite_raises():
    from gaussian_process_b200 import get_engine
    from gaussian_process_b200._lib import COV_SE
    X, y, _ = O.synth_c1(10, 5)
    with pytest.raises(np.linalg.LinAlgError):
        get_engine().small_lml_grad(COV_SE, X, y, [1.0, 1.0], -3.0)


@pytest.mark.parametrize("fused", [True, False])
def test_prior_draws_both_paths(G, golden, fused):
    """GP_regression.f_prior (:71-92): golden draws of the unmodified reference; the fused path factors
    k(X*,X*) + s I in one launch and forms L z in a second."""
    from gaussian_process_b200 import get_engine
    G.FUSED_SMALL_PATH = fused
    g = golden("ka1_regression.npz")
    Xs = g["n5_Xs"]
    np.random.seed(11)
    before = get_engine().launches()
    fp = G.f_prior(Xs, np.zeros((100, 1)), "rbf", 1, 3)
    if fused:
        assert get_engine().launches() - before == 2
    assert fp.shape == (100, 3) and rel(fp, g["n5_fprior"]) < 1e-8
    np.random.seed(4)
    per = G.f_prior(Xs, np.ones((100, 1)), "per", [2.0, 1.5], 2)
    np.random.seed(4)
    assert rel(per, O.f_prior(Xs, np.ones((100, 1)), "per", [2.0, 1.5], 2)) < 1e-7
    # a negative "noise" makes K + s I indefinite: LinAlgError, as np.linalg.cholesky at :90
    old = G.NOISE_VARIANCE
    try:
        G.NOISE_VARIANCE = -2.0
        with pytest.raises(np.linalg.LinAlgError):
            G.f_prior(Xs, np.zeros((100, 1)), "rbf", 1, 1)
    finally:
        G.NOISE_VARIANCE = old
