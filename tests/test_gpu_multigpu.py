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


def test_two_rank_nccl_job():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "mg_check.py"), "--npoints", "3000", "--block", "256"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MG_CHECK_OK" in out.stdout and "MG_SHARD_OK" in out.stdout
