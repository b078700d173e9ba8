"""Block-cyclic multi-GPU path: (a) the per-rank routines driven for P virtual ranks on one GPU (index maps,
prefix-structured inverse, batched triangular products), (b) the real NCCL path at world size 1, and
(c) when >= 2 GPUs are visible, a 2-rank torchrun job checked against the oracle."""
import os
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


@pytest.mark.parametrize("P,nb,n", [(1, 128, 300), (2, 128, 700), (3, 256, 1000), (4, 128, 1500), (8, 128, 1100)])
def test_emulated_ranks_match_oracle(eng, P, nb, n):
    from gaussian_process_b200._lib import COV_SE
    X, y = O.synth_c5(n, 16)
    lml, grad, alpha = eng.mg_emulate_fit_grad(P, COV_SE, X, y, [1.0, 4.0], 5e-4, nb=nb)
    lml_o, grad_o, alpha_o = O.rbf_fit_lml_grad(X, y, 1.0, 4.0)
    assert rel(lml, lml_o) < 1e-8
    assert rel(alpha, alpha_o) < 1e-7
    assert rel(grad[1], grad_o) < 1e-7


@pytest.mark.parametrize("group_k", [128, 256, 512])
def test_emulated_ranks_group_sizes(eng, group_k):
    """Grouped / deferred trailing updates: every group size (incl. groups of one panel and a short last group)."""
    from gaussian_process_b200._lib import COV_SE, check
    X, y = O.synth_c5(1700, 16)
    check(eng.lib.gpx_mg_set_group_k(group_k), "gpx_mg_set_group_k")
    try:
        lml, grad, alpha = eng.mg_emulate_fit_grad(3, COV_SE, X, y, [1.0, 4.0], 5e-4, nb=128)
    finally:
        check(eng.lib.gpx_mg_set_group_k(1024), "gpx_mg_set_group_k")
    lml_o, grad_o, alpha_o = O.rbf_fit_lml_grad(X, y, 1.0, 4.0)
    assert rel(lml, lml_o) < 1e-8 and rel(alpha, alpha_o) < 1e-7 and rel(grad[1], grad_o) < 1e-7


@pytest.mark.parametrize("P,nb,n", [(2, 128, 900), (3, 128, 1000), (8, 128, 2100)])
def test_emulated_ranks_plain_cyclic_map(eng, P, nb, n):
    """The plain block-cyclic map (gpx_mg_set_layout(0)) stays supported next to the default boustrophedon map."""
    from gaussian_process_b200._lib import COV_SE, check
    X, y = O.synth_c5(n, 16)
    check(eng.lib.gpx_mg_set_layout(0), "gpx_mg_set_layout")
    try:
        lml, grad, alpha = eng.mg_emulate_fit_grad(P, COV_SE, X, y, [1.0, 4.0], 5e-4, nb=nb)
    finally:
        check(eng.lib.gpx_mg_set_layout(1), "gpx_mg_set_layout")
    lml_o, grad_o, alpha_o = O.rbf_fit_lml_grad(X, y, 1.0, 4.0)
    assert rel(lml, lml_o) < 1e-8 and rel(alpha, alpha_o) < 1e-7 and rel(grad[1], grad_o) < 1e-7


def test_emulated_ranks_co2_all_theta(eng):
    from gaussian_process_b200._lib import COV_CO2
    X, y, _ = O.synth_c2(500)
    th = O.CO2_THETA_BOOK
    lml, grad, _ = eng.mg_emulate_fit_grad(3, COV_CO2, X, y, th, 5e-4, nb=128)
    K = O.co2_covariance(X, X, th) + O.S_NOISE * np.eye(500)
    Kinv = np.linalg.inv(K)
    ref = O.lml_grad_from(Kinv @ y, Kinv, O.co2_dcov(X, th))
    assert rel(lml, O.co2_lml(X, y, th)) < 1e-8
    assert np.all(np.abs(grad - ref) <= 1e-6 * np.maximum(1.0, np.abs(ref)))


def test_world1_driver_matches_single_gpu_path(eng):
    from gaussian_process_b200._lib import COV_SE
    X, y = O.synth_c5(900, 16)
    lml, grad, alpha = eng.mg_fit_grad(COV_SE, X, y, [1.0, 4.0], 5e-4, nb=256)
    fit = eng.fit(COV_SE, X, y, [1.0, 4.0], 5e-4, with_grad=True)
    assert rel(lml, fit.lml) < 1e-10 and rel(grad, fit.grad) < 1e-8


def test_world1_factor_and_solve_of_scaled_matrix(eng):
    """gpx_mg_factor on B = I + diag(sw) K diag(sw) (the Laplace matrix) + gpx_mg_potrs_vec vs NumPy."""
    from gaussian_process_b200._lib import COV_SE, check
    n, nb = 700, 256
    X, y, _ = O.synth_c3(n, 8)
    sw = np.sqrt(np.random.RandomState(0).uniform(0.05, 0.25, n))
    Xd = eng.to_device(X)
    ws = eng.mg_workspace(n, nb)
    lay = eng.mg_layout(n, nb)
    swd = eng.zeros(lay["npad"])
    swd[:n] = eng.to_device(sw)
    eng.mg_factor(COV_SE, Xd, [1.0, 1.0], 1.0, swd, nb, ws)
    B = np.eye(n) + sw[:, None] * O.rbf_kernel(X, X, 1, 1) * sw[None, :]
    L = eng.to_host(ws[lay["Lfull"]:lay["Lfull"] + lay["npad"] ** 2].view(lay["npad"], lay["npad"]))
    assert rel(L[:n, :n], np.linalg.cholesky(B)) < 1e-12
    assert np.all(L[np.triu_indices(lay["npad"], 1)] == 0.0)
    rhs = np.random.RandomState(1).randn(n)
    x = eng.zeros(lay["npad"])
    x[:n] = eng.to_device(rhs)
    eng._sync_stream()
    check(eng.lib.gpx_mg_potrs_vec(eng.h, n, nb, eng._p(ws), eng._p(x)), "gpx_mg_potrs_vec")
    assert rel(eng.to_host(x[:n]), np.linalg.solve(B, rhs)) < 1e-11


def test_world1_mg_fit_then_predict_from_replicated_factor(eng):
    """mg_fit keeps the replicated factor; mg_predict predicts from it (no refit): GP_regression.py:143-148."""
    from gaussian_process_b200._lib import COV_SE
    Xc, yc, Xs = O.synth_c1(300, 333)
    fit = eng.mg_fit(COV_SE, Xc, yc, [1.0, 1.0], 5e-4, nb=256)
    mu, var = eng.mg_predict(fit, Xs)
    np.random.seed(0)
    mu_o, sd_o, _ = O.regression_prediction(Xc, Xs, yc, 'rbf', 1, 1)
    assert rel(mu, mu_o) < 1e-8 and rel(var, sd_o ** 2) < 1e-8
    assert abs(fit.lml - O.rbf_lml(Xc, yc, 1, 1)) <= 1e-8 * abs(fit.lml)


def test_world1_distributed_binary_laplace_matches_oracle(eng):
    """BinaryLaplaceDistributed (B factored by the block-cyclic driver, one C call per Newton step) vs the oracle."""
    from gaussian_process_b200.distributed import BinaryLaplaceDistributed
    X, y, _ = O.synth_c3(900, 8)
    m = BinaryLaplaceDistributed(eng, X, 1.0, 1.0, nb=256)
    it = m.fit_newton(y, tolerance=1e-9)
    f_o, w_o, g_o, L_o, it_o = O.binary_training_newton(O.rbf_kernel(X, X, 1, 1), y, tolerance=1e-9)
    assert it == it_o
    assert rel(eng.to_host(m.f[:900]), f_o) < 1e-6 and rel(eng.to_host(m.g[:900]), g_o) < 1e-6


def test_two_rank_nccl_job():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "mg_check.py"), "--npoints", "3000", "--block", "256"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MG_CHECK_OK" in out.stdout and "MG_SHARD_OK" in out.stdout and "MG_LAPLACE_OK" in out.stdout
