"""CPU tests: the NumPy oracle port (oracle/gp_oracle.py) against (a) the committed golden vectors
generated from the unmodified reference and (b) the live reference when /root/reference exists."""
import contextlib
import io
import os
import tempfile

import numpy as np
import pytest

from oracle import gp_oracle as O
from oracle.ref_loader import load_reference, reference_available

RTOL = 1e-12


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# ----------------------------------------------------------------------------- goldens
def test_golden_ka1_regression(golden):
    g = golden("ka1_regression.npz")
    for tag, N in (("n64", 64), ("n5", 5)):
        X, y, Xs = O.synth_c1(N, 100)
        assert np.array_equal(X, g[tag + "_X"]) and np.array_equal(y, g[tag + "_y"])
        for kc, par in (("rbf", 1), ("per", [2.0, 1.5])):
            np.random.seed(7)
            mu, sd, fp = O.regression_prediction(X, Xs, y, kc, par, 10)
            assert rel(mu, g["%s_%s_mu" % (tag, kc)]) < RTOL
            assert rel(sd, g["%s_%s_sd" % (tag, kc)]) < 1e-10
            assert rel(fp, g["%s_%s_fpost" % (tag, kc)]) < 1e-9
        assert rel(O.rbf_lml(X, y, 1, 1), g[tag + "_lml"]) < RTOL
        assert rel(O.rbf_kernel(X, Xs, 1.3, 0.7), g[tag + "_K_rbf"]) < 1e-15
        assert rel(O.lin_kernel(X, Xs, 0.5), g[tag + "_K_lin"]) < 1e-15
        assert rel(O.per_kernel(X, Xs, [2.0, 1.5]), g[tag + "_K_per"]) < 1e-15
        np.random.seed(11)
        assert rel(O.f_prior(Xs, np.zeros((100, 1)), "rbf", 1, 3), g[tag + "_fprior"]) < 1e-12


def test_known_answers_appendix_b(golden):
    """SURVEY.md Appendix B literal values (KA1, KA2, KA3)."""
    g = golden("ka1_regression.npz")
    assert abs(float(g["n64_lml"]) - 101.72916291948931) < 1e-9
    assert abs(g["n64_rbf_mu"][0] - 0.9716562751701314) < 1e-12
    g2 = golden("ka2_lml_grad.npz")
    assert abs(float(g2["lml"]) - (-2697.120092406408)) < 1e-8
    assert abs(float(g2["dlml_dl"]) - (-3205.55525254022)) < 1e-6
    g3 = golden("ka3_co2.npz")
    assert abs(float(g3["lml_468"]) - (-268.28556910319685)) < 1e-8
    assert abs(float(g3["lml_2048"]) - (-1157.7137968966385)) < 1e-7


def test_golden_ka2_lml_grad(golden):
    g = golden("ka2_lml_grad.npz")
    X, y = O.synth_c5(512, 16)
    lml, grad, alpha = O.rbf_fit_lml_grad(X, y, 1.0, 4.0)
    assert rel(lml, g["lml"]) < RTOL
    assert rel(grad, g["dlml_dl"]) < 1e-9
    assert rel(alpha, g["alpha"]) < 1e-9
    # Hadamard-sum restatement == the reference's GEMM+trace form
    L = np.linalg.cholesky(O.rbf_kernel(X, X, 1.0, 4.0) + 5e-4 * np.eye(512))
    Kinv = np.dot(np.linalg.inv(L.T), np.linalg.inv(L))
    gh = O.lml_grad_from(alpha, Kinv, O.rbf_dcov(X, 1.0, 4.0))
    assert rel(gh[1], g["dlml_dl"]) < 1e-9
    # the "best-effort CPU" timing baseline (dpotrf/dpotri, O(N^2) trace) computes the same numbers
    tm = {}
    lml_b, grad_b, alpha_b = O.rbf_fit_lml_grad_best_effort(X, y, 1.0, 4.0, chunk=100, timings=tm)
    assert rel(lml_b, g["lml"]) < RTOL and rel(grad_b, g["dlml_dl"]) < 1e-9 and rel(alpha_b, g["alpha"]) < 1e-9
    assert tm["n2"] > 0 and tm["n3"] > 0


def test_golden_ka2_tune_and_bo(golden):
    g = golden("ka2_tune_first.npz")
    X, y, Xs = O.synth_c1(8, 100)
    np.random.seed(3)
    mu, sd, fp, lml, l, it = O.tune_first(X, Xs, y, 2, 1, np.array([1.7]))
    assert rel(mu, g["mu"]) < 1e-10 and rel(sd, g["sd"]) < 1e-9
    assert rel(lml, g["lml"]) < 1e-11
    assert ("after %d iterations" % it) in str(g["stdout"])
    assert rel(fp, g["fpost"]) < 1e-7
    b = golden("ka2_bo.npz")
    np.random.seed(5)
    mu, sd, fp = O.bo_posterior_1d(b["lt"], b["ltest"], b["yl"])
    assert rel(mu, b["mu"]) < RTOL and rel(sd, b["sd"]) < 1e-10 and rel(fp, b["fpost"]) < 1e-9


def test_golden_ka3_co2(golden):
    g = golden("ka3_co2.npz")
    X, y, Xs = O.synth_c2(468)
    th = O.CO2_THETA_BOOK
    assert rel(O.co2_covariance(X, X, th), g["K_468"]) < 1e-15
    assert rel(O.co2_covariance(X, Xs, th), g["Ks_468"]) < 1e-15
    assert rel(O.co2_lml(X, y, th), g["lml_468"]) < 1e-12
    np.random.seed(9)
    mu, sd, fp = O.co2_make_prediction(X, Xs, y, th)
    assert rel(mu, g["mu"]) < 1e-12 and rel(sd, g["sd"]) < 1e-10 and rel(fp, g["fpost"]) < 1e-8
    mu_bo, sd_bo = O.co2_bo_posterior(g["bo_theta_train"], g["bo_theta_test"], g["bo_y"])
    assert rel(mu_bo, g["bo_mu"]) < 1e-12
    assert np.allclose(sd_bo, g["bo_sd"], rtol=1e-9, equal_nan=True)


def test_co2_gradient_finite_differences():
    """Appendix C derivatives are not in the reference: pin them by 4th-order central differences of
    the (reference-pinned) LML.  Noise s=1 keeps the FD itself well conditioned (rel. err ~1e-8)."""
    X, y, _ = O.synth_c2(96)
    th = O.CO2_THETA_BOOK.copy()
    s = 1.0
    K = O.co2_covariance(X, X, th) + s * np.eye(96)
    Kinv = np.linalg.inv(K)
    alpha = Kinv @ y
    g = O.lml_grad_from(alpha, Kinv, O.co2_dcov(X, th))

    def f(t):
        return O.co2_lml(X, y, t, s=s)

    for j in range(11):
        h = 1e-3 * th[j]
        e = np.zeros(11)
        e[j] = h
        fd = (8 * (f(th + e) - f(th - e)) - (f(th + 2 * e) - f(th - 2 * e))) / (12 * h)
        assert abs(fd - g[j]) <= 1e-6 * max(1.0, abs(g[j])), (j, fd, g[j])


def test_golden_ka4_binary(golden):
    g = golden("ka4_binary.npz")
    X, y, fpr = g["X"], g["y"], g["f_prior"]
    K = O.rbf_kernel(X, X, 1, 1)
    W, L_inv, fd, f, errs = O.binary_training_reference(K, y, fpr, 1)
    assert len(errs) == len(g["errors"]) == 138
    assert rel(errs, g["errors"]) < 1e-9
    assert rel(np.diag(W), g["Wdiag"]) < 1e-15 and rel(L_inv, g["L_inv"]) < 1e-12
    assert rel(fd, g["first_deri"]) < 1e-15
    assert abs(np.trace(W) - 30.169878124609628) < 1e-10
    fs, var, lab = O.binary_predict_reference(g["Xq"], X, L_inv, W, fd, 1)
    assert rel(fs, g["fbar"]) < 1e-12 and rel(var, g["var"]) < 1e-10
    assert np.array_equal(lab == 1, g["is_plus"])


def test_binary_newton_is_a_mode():
    """Textbook mode: at the returned f, f = K grad log p(y|f) (stationarity of the Laplace objective)."""
    rs = np.random.RandomState(5)
    X = rs.randn(96, 2)
    y = np.where(X[:, 0] * X[:, 1] > 0, 1.0, -1.0)
    K = O.rbf_kernel(X, X, 1, 1)
    f, w, g, L, it = O.binary_training_newton(K, y)
    assert it < 30
    assert np.max(np.abs(f - K @ g)) < 1e-9


def test_golden_ka5_multi(golden):
    g = golden("ka5_multi.npz")
    Ks = O.rbf_kernel(g["Xtr"], g["Xtr"], 1, 1)
    from scipy.linalg import block_diag
    pi, f, errs = O.multi_training_reference(block_diag(Ks, Ks, Ks), g["y_targets"], 3, 60)
    assert len(errs) == len(g["errors"]) == 18
    assert rel(errs, g["errors"]) < 1e-9 and rel(pi, g["pi_vector"]) < 1e-10
    _, am = O.multi_predict_reference(g["Xte"], g["Xtr"], 3, g["y_targets"], pi, 1)
    assert np.array_equal(am == g["yte"], g["hits"])
    assert abs(np.mean(g["hits"]) - 0.875) < 1e-12
    pv, pm = O.compute_pi(g["fprobe"], 3, 60)
    assert rel(pv, g["pi_probe"]) < 1e-15 and rel(pm, g["pim_probe"]) < 1e-15


def test_multi_newton_is_a_mode():
    """Textbook Alg 3.3: at convergence f_c = K (y_c - pi_c) for every class."""
    X, labels, y, _, _ = O.synth_c4(n=60, C=3, D=2, n_test=4)
    Ks = O.rbf_kernel(X, X, 1, 1)
    p, f, it = O.multi_training_newton(Ks, y, 3, 60)
    assert it < 50
    resid = y.reshape(3, 60) - p
    assert np.max(np.abs(f - (Ks @ resid.T).T)) < 1e-7


# ----------------------------------------------------------------------------- live reference
needs_ref = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@needs_ref
def test_oracle_matches_live_reference_regression():
    R = load_reference()
    G, T, C2 = R["GP_regression"], R["tune_hyperparms_regression"], R["CO2_example"]
    rs = np.random.RandomState(42)
    for D in (1, 3, 16):
        a, b = rs.randn(37, D), rs.randn(23, D)
        assert rel(O.rbf_kernel(a, b, 0.9, 1.7), G.RBF_kernel(a, b, 0.9, 1.7)) < 1e-15
        assert rel(O.rbf_kernel(a, b, 0.9, 1.7, chunk=5), G.RBF_kernel(a, b, 0.9, 1.7)) < 1e-15
        th = O.CO2_THETA_BOOK * (0.8 + 0.4 * rs.rand(11))
        assert rel(O.co2_covariance(a, b, th), C2.covariance_function(a, b, th)) < 1e-14
        assert rel(O.co2_covariance(a, a, th), C2.covariance_function(a, a, th)) < 1e-14
    X, y = O.synth_c5(200, 16)
    assert rel(O.rbf_lml(X, y, 1.2, 3.0), T.compute_mar_likelihood(X, None, y, 1.2, 3.0)) < 1e-13
    K = G.RBF_kernel(X, X, 1.2, 3.0)
    L = np.linalg.cholesky(K + 5e-4 * np.eye(200))
    alpha = np.linalg.solve(L.T, np.linalg.solve(L, y))
    Kinv = np.dot(np.linalg.inv(L.T), np.linalg.inv(L))
    _, l_new = T.gradient_ascent(X, X, 1.2, 3.0, alpha.reshape(-1, 1), Kinv)
    assert rel(O.rbf_grad_l(X, 1.2, 3.0, alpha, Kinv), (l_new - 3.0) / 0.01) < 1e-9


@needs_ref
def test_oracle_matches_live_reference_classifiers():
    R = load_reference()
    G, B, M = R["GP_regression"], R["GP_binary_classification"], R["GP_multi_classification"]
    rs = np.random.RandomState(8)
    X = rs.randn(40, 2)
    y = np.where(X[:, 0] + X[:, 1] > 0, 1, -1).reshape(-1, 1)
    K = G.RBF_kernel(X, X, 1, 1)
    fpr = 0.3 * rs.randn(40, 1)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            W, L_inv, g = B.model_training(K, y, fpr, 1)
    finally:
        os.chdir(cwd)
    W2, L_inv2, g2, f2, errs = O.binary_training_reference(K, y, fpr, 1)
    assert rel(W2, W) < 1e-15 and rel(L_inv2, L_inv) < 1e-13 and rel(g2, g) < 1e-15
    f = rs.randn(180)
    pv, pm = M.compute_pi(f, 3, 60)
    pv2, pm2 = O.compute_pi(f, 3, 60)
    assert rel(pv2, pv) < 1e-15 and rel(pm2, pm) < 1e-15


@needs_ref
def test_oracle_matches_live_reference_on_random_shapes():
    """Property-style pinning: 20 seeded random sizes / dimensions / hyper-parameters, every regression-path function
    of the oracle against the unmodified reference (prediction, lin / per kernels, CO2 LML and prediction, SE LML)."""
    R = load_reference()
    G, T, C2 = R["GP_regression"], R["tune_hyperparms_regression"], R["CO2_example"]
    cfg = np.random.RandomState(2024)
    for case in range(20):
        N, n, D = int(cfg.randint(2, 41)), int(cfg.randint(1, 31)), int(cfg.randint(1, 6))
        l, nf, seed = float(cfg.uniform(0.4, 3.0)), int(cfg.randint(1, 5)), int(cfg.randint(0, 10 ** 6))
        rs = np.random.RandomState(seed)
        X, Xs = rs.uniform(-3, 3, (N, D)), rs.uniform(-3, 3, (n, D))
        y = np.sin(X.sum(1)) + 0.05 * rs.randn(N)
        np.random.seed(seed)
        mu, sd, fp = O.regression_prediction(X, Xs, y, 'rbf', l, nf)
        np.random.seed(seed)
        mu_r, sd_r, fp_r = G.prediction(X, Xs, y, 'rbf', l, nf)
        assert rel(mu, mu_r) < 1e-12 and rel(fp, fp_r) < 1e-9, case
        assert np.allclose(sd ** 2, sd_r ** 2, rtol=0, atol=1e-12, equal_nan=True), case
        assert rel(O.lin_kernel(X, Xs, l), G.lin_kernel(X, Xs, l)) < 1e-14, case
        x1, xs1 = X[:, :1], Xs[:, :1]
        assert rel(O.per_kernel(x1, xs1, [2.0, l]), G.per_kernel(x1, xs1, [2.0, l])) < 1e-14, case
        # CO2 composite on time-like 1-D inputs (its real use).  With theta_1 = 66 some draws are numerically
        # indefinite: then BOTH sides raise LinAlgError (CO2_example.py:143 / :212); otherwise the numbers agree.
        t = 1958 + np.sort(rs.uniform(0, 40, N))[:, None]
        ts = 1998 + np.sort(rs.uniform(0, 5, n))[:, None]
        yt = 3 * np.sin(2 * np.pi * t.ravel()) + 0.1 * (t.ravel() - 1958) ** 2 + 0.3 * rs.randn(N)
        yt -= yt.mean()
        th = O.CO2_THETA_BOOK * (0.7 + 0.6 * rs.rand(11))

        def both(fo, fr, *a):
            outs = []
            for fn in (fo, fr):
                np.random.seed(seed)
                try:
                    with np.errstate(invalid="ignore"):
                        outs.append(fn(*a))
                except np.linalg.LinAlgError:
                    outs.append(None)
            assert (outs[0] is None) == (outs[1] is None), case
            return outs

        lo, lr = both(O.co2_lml, C2.compute_mar_likelihood, t, yt, th)
        if lo is not None:
            assert rel(lo, lr) < 1e-10, case
        po, pr = both(O.co2_make_prediction, C2.make_prediction, t, ts, yt, th)
        if po is not None:
            assert rel(po[0], pr[0]) < 1e-9 and np.allclose(po[1] ** 2, pr[1] ** 2, rtol=1e-8, atol=1e-8, equal_nan=True), case
        assert rel(O.rbf_lml(X, y, 1.0, l), T.compute_mar_likelihood(X, None, y, 1.0, l)) < 1e-12, case
