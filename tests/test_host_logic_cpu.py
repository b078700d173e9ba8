"""CPU tests of the host-side logic of the drop-in modules (no GPU needed): the Bayesian-optimisation layer
(acquisition functions, candidate samplers, index helpers) against the live reference when /root/reference exists,
and invariants that hold everywhere.  These functions never touch the engine."""
import os
import contextlib
import io
import random

import numpy as np
import pytest

from oracle.ref_loader import load_reference, reference_available

needs_ref = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


def _mods():
    from gaussian_process_b200 import CO2_example as C2, tune_hyperparms_regression as T
    return T, C2


def test_overlap_and_candidate_grid_invariants():
    T, C2 = _mods()
    a = np.array([0.01, 2.505, 4.0])
    grid = np.linspace(0.01, 5, 3)
    ia, ib = T.overlap(a, grid)
    assert ia.tolist() == [0, 1] and ib.tolist() == [0, 1] and ib.dtype.kind == "i"
    random.seed(1)
    done = np.array([0.5, 3.5])
    cand = T.random_gen_test_parms(100, done)
    assert cand.shape == (100, 1) and np.all(np.diff(cand[:, 0]) > 0)
    assert not np.isin(cand[:, 0], done).any() and cand.min() >= 0.01 and cand.max() <= 5
    th = C2.init_hyperms(5, 11)
    assert th.shape == (5, 11) and np.allclose(th[2] - C2.HYPERMS_BOOK, 0.5 * 7)
    random.seed(2)
    tp = C2.random_sample_test_parms(50, th)
    assert tp.shape == (50, 11)
    assert np.all(tp >= C2.HYPERMS_BOOK * 0.3 - 1e-12) and np.all(tp <= C2.HYPERMS_BOOK * 1.5 + 1e-12)


def test_acquisition_functions_choose_the_expected_points():
    T, C2 = _mods()
    params = np.linspace(0, 1, 11).reshape(-1, 1)
    means = -(params[:, 0] - 0.6) ** 2
    sd = np.full(11, 0.1)
    done = np.array([0.0, 1.0])
    y = np.array([-0.36, -0.16])
    assert float(T.EI(params, means, sd, done, y, 3, 0)[0]) == pytest.approx(0.6)
    assert float(np.ravel(T.UCB(done, params, means, sd, 3, 0))[0]) == pytest.approx(0.6)
    random.seed(0)
    assert float(np.ravel(T.PI(params, means, sd, done, y, 3, 0))[0]) == pytest.approx(0.6)
    # early stop: improvement impossible -> PI returns True
    with contextlib.redirect_stdout(io.StringIO()):
        assert T.PI(params, means - 100, sd, done, y, 3, 0) is True
    th_test = np.arange(44, dtype=float).reshape(4, 11)
    mu = np.array([0.1, 0.9, 0.3, 0.2])
    s = np.array([0.1, 0.1, 0.1, 0.1])
    assert np.array_equal(C2.EI(th_test, mu, s, np.array([0.0])), th_test[1])
    assert np.array_equal(C2.UBC(th_test[:1], th_test, mu, s), th_test[1])
    assert C2.UBC(th_test[1:2], th_test, mu, s) is True
    random.seed(0)
    assert np.array_equal(C2.PI(th_test, mu, s, np.array([0.0])), th_test[1])
    # as shipped, the whole `choice` list falls through to PI (CO2_example.py:312-313,359)
    random.seed(0)
    assert np.array_equal(C2.acquisition_fun(['UCB', 'TS', 'EI', 'PI'], th_test[:1], th_test, mu, s, np.array([0.0])), th_test[1])


@needs_ref
def test_acquisition_functions_match_the_live_reference():
    T, C2 = _mods()
    R = load_reference()
    RT, RC = R["tune_hyperparms_regression"], R["CO2_example"]
    rs = np.random.RandomState(4)
    params = np.sort(rs.uniform(0.01, 5, 60)).reshape(-1, 1)
    means = rs.randn(60)
    sd = 0.1 + rs.rand(60)
    done = np.array([0.7, 2.2])
    y = np.array([-1.0, 0.2])
    assert np.array_equal(T.EI(params, means, sd, done, y, 3, 0), RT.EI(params, means, sd, done, y, 3, 0))
    assert np.array_equal(T.UCB(done, params, means, sd, 3, 0), RT.UCB(done, params, means, sd, 3, 0))
    random.seed(3)
    a = T.PI(params, means, sd, done, y, 3, 0)
    random.seed(3)
    b = RT.PI(params, means, sd, done, y, 3, 0)
    assert np.array_equal(a, b)
    ia, ib = T.overlap(np.array([0.01, 5.0]), np.linspace(0.01, 5, 7))
    ra, rb = RT.overlap(np.array([0.01, 5.0]), np.linspace(0.01, 5, 7))
    assert np.array_equal(ia, ra) and np.array_equal(ib, np.asarray(rb, dtype=np.int64))
    th_test = RC.init_hyperms(8, 11) * (0.5 + rs.rand(8, 11))
    mu, s, yv = rs.randn(8), 0.1 + rs.rand(8), rs.randn(5)
    assert np.array_equal(C2.EI(th_test, mu, s, yv), RC.EI(th_test, mu, s, yv))
    assert np.array_equal(C2.UBC(th_test[:2], th_test, mu, s), RC.UBC(th_test[:2], th_test, mu, s))
    assert np.array_equal(C2.init_hyperms(5, 11), RC.init_hyperms(5, 11))


def test_dataset_generators_are_seed_compatible_with_the_reference_recipe():
    """GP_regression.dataset_generator consumes the global NumPy RNG exactly like the reference (:58-68)."""
    from gaussian_process_b200 import GP_regression as G
    np.random.seed(11)
    f, X, y, Xs = G.dataset_generator(7, 50)
    np.random.seed(11)
    Xr = np.random.uniform(-5, 5, size=(7, 1))
    yr = np.sin(0.9 * Xr).flatten() + np.sqrt(0.0005) * np.random.randn(7)
    assert np.array_equal(X, Xr) and np.array_equal(y, yr) and Xs.shape == (50, 1)
    assert np.allclose(f(Xs), np.sin(0.9 * Xs).ravel())


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs next to ours) emits one JSON line with the contract keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-n", "256"], capture_output=True, text=True, timeout=300, check=True).stdout
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "s" and d["higher_is_better"] is False and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "N=65536" in d["config"]["workload"]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm; no GPU needed) prints ONE JSON line with the contract's keys, for the
    default config and for a per-config metric, also when torchrun's OMP_NUM_THREADS=1 is in the environment."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for extra, metric in ((["--cpu-sample-n", "256"], "gp_fit_lml_grad_seconds_n65536_fp64"),
                          (["--config", "c1"], "gp_regression_prediction_seconds_n5_fp64")):
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"] + extra,
                             capture_output=True, text=True, timeout=600, env=env, cwd=root)
        assert out.returncode == 0, out.stderr[-1500:]
        lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == "s" and d["higher_is_better"] is False
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0 and "workload" in d["config"]
