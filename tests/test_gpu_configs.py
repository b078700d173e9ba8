"""Parity at the BASELINE.json configuration sizes and coverage of the drop-in entry points that
had no test: C2 at N=8192 (cond ~ 1e8, the hard case of SURVEY 7.3), C3 textbook Newton at N=4096 D=8, C4 Alg-3.3 at
C=10 n=1024, kernel_1..4, the CO2 Bayesian-optimisation driver against a golden run of the reference, the softmax
kernel with overlapping strides, synthetic_mauna_loa and the scripts' __main__ demos.  Everything goes through the C ABI.

Tolerances (north_star): mean / variance / LML 1e-8 relative (max-norm), Laplace modes 1e-6.
"""
import contextlib
import io
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def eng():
    from gaussian_process_b200 import get_engine
    return get_engine(0)


# ------------------------------------------------------------------------------------------------ C2 at full size
@pytest.fixture(scope="module")
def c2_full():
    return O.synth_c2(8192, 240)


def test_c2_lml_at_n8192_matches_oracle(c2_full):
    """CO2_example.py:131-149 at the BASELINE size; the oracle follows the reference's LU-solve / inv(L) path."""
    from gaussian_process_b200 import CO2_example as C2
    X, y, _ = c2_full
    lml = C2.compute_mar_likelihood(X, y, O.CO2_THETA_BOOK)
    lml_o = O.co2_lml(X, y, O.CO2_THETA_BOOK)
    assert abs(lml - lml_o) <= 1e-8 * abs(lml_o), (lml, lml_o)


def test_c2_prediction_at_n8192_matches_oracle(c2_full):
    """CO2_example.py:182-214 at N=8192, 240 test points: mean and variance within 1e-8 (max-norm relative)."""
    from gaussian_process_b200 import CO2_example as C2
    X, y, Xs = c2_full
    np.random.seed(4)
    mu, sd, fp = C2.make_prediction(X, Xs, y, O.CO2_THETA_BOOK)
    np.random.seed(4)
    mu_o, sd_o, fp_o = O.co2_make_prediction(X, Xs, y, O.CO2_THETA_BOOK)
    assert rel(mu, mu_o) < 1e-8
    assert rel(sd ** 2, sd_o ** 2) < 1e-8
    assert rel(fp, fp_o) < 1e-5     # one posterior draw: chol of a jittered 240 x 240 difference matrix


# ------------------------------------------------------------------------------------------------ C3 / C4 reduced
def test_c3_newton_n4096_d8_matches_oracle(eng):
    """GP_binary_classification.py:86-133 in textbook mode (R&W Alg. 3.1) at N=4096, D=8: mode within 1e-6."""
    from gaussian_process_b200._lib import COV_SE
    from gaussian_process_b200.laplace import BinaryLaplace
    X, y, _ = O.synth_c3(4096, 8)
    Xd = eng.to_device(X)
    Kd = eng.cov(COV_SE, Xd, Xd, [1.0, 1.0], same_x=True)
    m = BinaryLaplace(eng, Kd, 4096)
    it = m.fit_newton(y, tolerance=1e-8)
    f_o, w_o, g_o, L_o, it_o = O.binary_training_newton(O.rbf_kernel(X, X, 1, 1), y, tolerance=1e-8)
    assert it == it_o
    assert rel(eng.to_host(m.f[:4096]), f_o) < 1e-6
    assert rel(eng.to_host(m.g[:4096]), g_o) < 1e-6
    assert rel(eng.to_host(m.w[:4096]), w_o) < 1e-6


def test_c4_alg33_c10_n1024_matches_oracle(eng):
    """GP_multi_classification.py:66-126 (textbook Alg. 3.3) at C=10, n=1024, D=16: mode and pi within 1e-6."""
    from gaussian_process_b200._lib import COV_SE
    from gaussian_process_b200.laplace import MultiLaplaceNewton
    X, labels, y, Xt, tl = O.synth_c4(1024, 10, 16, 16)
    Xd = eng.to_device(X)
    Kd = eng.cov(COV_SE, Xd, Xd, [1.0, 1.0], same_x=True)
    model = MultiLaplaceNewton(eng, Kd, 10, 1024)
    it = model.fit(y, tolerance=1e-8)
    p_o, f_o, it_o = O.multi_training_newton(O.rbf_kernel(X, X, 1, 1), y, 10, 1024, tolerance=1e-8)
    assert it == it_o
    assert rel(eng.to_host(model.f), f_o) < 1e-6
    assert rel(eng.to_host(model.pi), p_o) < 1e-6


# ------------------------------------------------------------------------------------------------ A3 kernel_1..4
def test_co2_kernel_terms_match_reference_formulas():
    """kernel_1 .. kernel_4 (CO2_example.py:9-66) as element-wise device maps of a host sqdist matrix, and their sum
    against covariance_function's golden matrix."""
    from gaussian_process_b200 import CO2_example as C2
    g = np.load(os.path.join(ROOT, "tests", "golden", "ka3_co2.npz"))
    th = O.CO2_THETA_BOOK
    X, y, Xs = O.synth_c2(468)
    d2 = O.sqdist(X, X)
    r = np.sqrt(d2)
    k1 = C2.kernel_1(d2, th[0], th[1])
    k2 = C2.kernel_2(r, d2, th[2], th[3], th[4])
    k3 = C2.kernel_3(d2, th[5], th[6], th[7])
    k4 = C2.kernel_4(d2, th[8], th[9], th[10])
    assert rel(k1, th[0] ** 2 * np.exp(-.5 * d2 / th[1] ** 2)) < 1e-14
    assert rel(k2, th[2] ** 2 * np.exp(-.5 * d2 / th[3] ** 2 + -2 * (np.sin(np.pi * r) / th[4]) ** 2)) < 1e-13
    assert rel(k3, th[5] ** 2 * (1.0 / np.power(1 + .5 * d2 / (th[7] * th[6] ** 2), th[7]))) < 1e-13
    assert rel(k4, th[8] ** 2 * np.exp(-.5 * d2 / th[9] ** 2) + th[10] ** 2 * np.eye(468)) < 1e-14
    assert rel(k1 + k2 + k3 + k4, g["K_468"]) < 1e-13
    d2s = O.sqdist(X, Xs)                       # rectangular block: no delta term (:60-63)
    assert rel(C2.kernel_4(d2s, th[8], th[9], th[10]), th[8] ** 2 * np.exp(-.5 * d2s / th[9] ** 2)) < 1e-14


# ------------------------------------------------------------------------------------------------ N2 BO driver
@pytest.mark.parametrize("fused", [True, False])
def test_co2_bo_driver_matches_reference_golden(fused):
    """CO2_example.tune_hyperparameters_BO (:330-379) end to end against a run of the unmodified reference
    (oracle/gen_golden.py KA6: 4 labels x 10 iterations x 500 candidates, `random`/NumPy seeded)."""
    from gaussian_process_b200 import CO2_example as C2
    from gaussian_process_b200 import GP_regression as G
    g = np.load(os.path.join(ROOT, "tests", "golden", "ka6_co2_bo.npz"))
    X, y, Xs = O.synth_c2(int(g["N"]), 24)
    old = G.FUSED_SMALL_PATH
    G.FUSED_SMALL_PATH = fused
    try:
        random.seed(42)
        np.random.seed(42)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            th = C2.tune_hyperparameters_BO(X, Xs, y)
    finally:
        G.FUSED_SMALL_PATH = old
    lines = buf.getvalue().splitlines()
    best = np.array([float(lines[i + 1]) for i, ln in enumerate(lines) if ln.endswith("th iteration!")])
    assert len(lines) == int(g["nlines"])
    assert rel(best, g["best_lml"]) < 1e-8
    assert abs(float(lines[-1]) - float(g["book_lml"])) <= 1e-8 * abs(float(g["book_lml"]))
    assert rel(th, g["theta"]) < 1e-12


# ------------------------------------------------------------------------------------------------ A11 softmax strides
@pytest.mark.parametrize("C,n,stride", [(3, 60, 60), (3, 100, 60), (4, 37, 50), (10, 1000, 60)])
def test_softmax_classes_overlapping_stride_is_last_writer_wins(eng, C, n, stride):
    """GP_multi_classification.py:36-63 with the literal stride 60: for n > 60 several (class, point) pairs write the
    same pi_vector entry and the reference's sequential loop keeps the last one."""
    ln = (C - 1) * stride + n
    f = np.random.RandomState(3).randn(ln)
    fd, pid = eng.to_device(f), eng.zeros(ln)
    eng._sync_stream()
    from gaussian_process_b200._lib import check
    check(eng.lib.gpx_softmax_classes(eng.h, C, n, stride, eng._p(fd), eng._p(pid)), "gpx_softmax_classes")
    pv, _ = O.compute_pi(f, C, n, stride)
    assert rel(eng.to_host(pid), pv) < 1e-14


def test_laplace_kernels_accept_65536_rows(eng):
    """build_B / scale_rows put rows on grid.x: N = 65536 (the headline size) exceeds the 65535 grid.y limit."""
    from gaussian_process_b200._lib import check
    rows, cols = 65536 + 128, 128
    M = eng.torch.ones(rows, cols, dtype=eng.torch.float64, device=eng.device)
    s = eng.torch.arange(rows, dtype=eng.torch.float64, device=eng.device)
    eng._sync_stream()
    check(eng.lib.gpx_scale_rows(eng.h, rows, cols, cols, eng._p(s), eng._p(M)), "gpx_scale_rows")
    assert float(M[65540, 7]) == 65540.0 and float(M[3, 0]) == 3.0


# ------------------------------------------------------------------------------------------------ N4 data / scripts
def test_synthetic_mauna_loa_shape_and_fit():
    from gaussian_process_b200 import CO2_example as C2
    X, y = C2.synthetic_mauna_loa()
    assert X.shape == (468, 1) and y.shape == (468,) and abs(X[0, 0] - 1958) < 1e-12 and abs(X[12, 0] - 1959) < 1e-12
    X2, y2 = C2.synthetic_mauna_loa()
    assert np.array_equal(y, y2)                       # seeded
    yc = y - y.mean()
    lml = C2.compute_mar_likelihood(X, yc, O.CO2_THETA_BOOK)
    assert abs(lml - O.co2_lml(X, yc, O.CO2_THETA_BOOK)) <= 1e-8 * abs(lml)


@pytest.mark.parametrize("module,needle", [("GP_regression", "posterior mean range"),
                                           ("GP_binary_classification", None),
                                           ("GP_multi_classification", None),
                                           ("tune_hyperparms_regression", None)])
def test_script_main_blocks_run_end_to_end(module, needle):
    """The reference is five scripts; each drop-in module's __main__ demo must run to completion on the GPU."""
    env = dict(os.environ, MPLBACKEND="Agg", PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([sys.executable, "-m", "gaussian_process_b200." + module], capture_output=True, text=True, timeout=900,
                         env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    if needle:
        assert needle in out.stdout


# ------------------------------------------------------------------------------------------------ N1 optimiser drivers
def _numpy_ascent(lml_grad_fn, theta0, mask, step, tol, max_iter):
    """The loop of tune_hyperparms_regression.py:121-153 around any (LML, gradient) oracle."""
    th = np.array(theta0, dtype=np.float64)
    prev, hist = 0.0, []
    used = th.copy()
    for it in range(max_iter):
        lml, g = lml_grad_fn(th)
        used = th.copy()
        th = th + step * np.where(mask, g, 0.0)
        err = abs(lml - prev)
        prev = lml
        hist.append(lml)
        if err <= tol:
            break
    return th, used, np.array(hist)


@pytest.mark.parametrize("use_graph", [True, False])
def test_ascent_sigma_and_l_matches_numpy_loop(eng, use_graph):
    """gpx_gp_ascent over BOTH SE hyper-parameters (CUDA-graph replay and eager) vs the same loop on the oracle."""
    from gaussian_process_b200._lib import COV_SE
    X, y = O.synth_c5(300, 3)

    def f(th):
        K = O.rbf_kernel(X, X, th[0], th[1]) + O.S_NOISE * np.eye(300)
        Kinv = np.linalg.inv(K)
        a = Kinv @ y
        lml = -.5 * y @ a - .5 * np.linalg.slogdet(K)[1] - 150 * np.log(2 * np.pi)
        return lml, O.lml_grad_from(a, Kinv, O.rbf_dcov(X, th[0], th[1]))

    th_o, used_o, hist_o = _numpy_ascent(f, [1.0, 1.5], np.array([1, 1]), 1e-4, 1e-3, 12)
    res = eng.ascend(COV_SE, X, y, [1.0, 1.5], [1, 1], O.S_NOISE, 1e-4, 1e-3, 12, use_graph=use_graph)
    assert res["iterations"] == len(hist_o)
    assert rel(res["history"], hist_o) < 1e-8
    assert rel(res["theta"], th_o) < 1e-7 and rel(res["theta_used"], used_o) < 1e-7


def test_ascent_co2_eleven_theta_matches_oracle_gradient_loop(eng):
    """All 11 CO2 hyper-parameters ascended on the device (CO2_example.tune_hyperparameters_gradient) vs the oracle."""
    from gaussian_process_b200 import CO2_example as C2
    X, y, _ = O.synth_c2(200)
    th0 = O.CO2_THETA_BOOK * 1.05

    def f(th):
        K = O.co2_covariance(X, X, th) + O.S_NOISE * np.eye(200)
        Kinv = np.linalg.inv(K)
        a = Kinv @ y
        return O.co2_lml(X, y, th), O.lml_grad_from(a, Kinv, O.co2_dcov(X, th))

    mask = np.array([1, 0, 1, 0, 1, 1, 1, 1, 1, 1, 1])
    th_o, used_o, hist_o = _numpy_ascent(f, th0, mask, 1e-6, 1e-9, 4)
    th, lml, it = C2.tune_hyperparameters_gradient(X, y, th0, mask=mask, step_size=1e-6, tolerance=1e-9, max_iter=4)
    assert it == 4 and abs(lml - hist_o[-1]) <= 1e-8 * abs(hist_o[-1])
    assert np.all(np.abs(th - th_o) <= 1e-6 * np.maximum(1.0, np.abs(th_o)))
    assert th[1] == th0[1] and th[3] == th0[3]          # masked hyper-parameters do not move


def test_tune_first_tiled_path_is_one_device_loop(eng):
    """tune_hyperparms_first above the fused small-problem size: same iterations / LML / moments as the oracle loop."""
    from gaussian_process_b200 import tune_hyperparms_regression as T
    X, y, Xs = O.synth_c1(200, 150)
    np.random.seed(2)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        mu, sd, fp, lml = T.tune_hyperparms_first(X, Xs, y, 2, 1.0, np.array([0.8]))
    np.random.seed(2)
    mu_o, sd_o, fp_o, lml_o, l_o, it_o = O.tune_first(X, Xs, y, 2, 1.0, np.array([0.8]))
    assert ("after %d iterations" % it_o) in buf.getvalue()
    assert abs(lml - lml_o) <= 1e-8 * abs(lml_o)
    assert rel(mu, mu_o) < 1e-7 and rel(sd ** 2, sd_o ** 2) < 1e-6
