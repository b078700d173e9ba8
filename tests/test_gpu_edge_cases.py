"""Edge cases of the drop-in API on the GPU: tiny / ragged (non-multiple-of-128) sizes, one test point, duplicate
inputs, high dimension, reuse of the engine across differently sized problems, l passed as a shape-(1,) array."""
import numpy as np
import pytest

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("N,n,D", [(1, 1, 1), (2, 3, 1), (127, 5, 2), (128, 128, 3), (129, 257, 1), (300, 1, 7), (513, 100, 40)])
def test_prediction_ragged_sizes(N, n, D):
    from gaussian_process_b200 import GP_regression as G
    rs = np.random.RandomState(N * 7 + n)
    X = rs.uniform(-3, 3, (N, D))
    y = np.sin(X.sum(1)) + 0.05 * rs.randn(N)
    Xs = rs.uniform(-3, 3, (n, D))
    np.random.seed(4)
    mu, sd, fp = G.prediction(X, Xs, y, 'rbf', 1.3, 2)
    np.random.seed(4)
    mu_o, sd_o, fp_o = O.regression_prediction(X, Xs, y, 'rbf', 1.3, 2)
    assert mu.shape == (n,) and sd.shape == (n,) and fp.shape == (n, 2)
    assert rel(mu, mu_o) < 1e-8 and rel(sd ** 2, sd_o ** 2) < 1e-8


@pytest.mark.parametrize("N", [1, 3, 130, 1000])
def test_lml_ragged_sizes_and_array_lengthscale(N):
    from gaussian_process_b200 import tune_hyperparms_regression as T
    X, y = O.synth_c5(N, 16)
    l = np.array([2.5])                      # shape-(1,) array as in tune...:408
    got = T.compute_mar_likelihood(X, None, y, 1, l)
    assert rel(got, O.rbf_lml(X, y, 1, 2.5)) < 1e-8
    assert isinstance(got, np.float64)


def test_duplicate_training_points_are_handled_like_the_reference():
    """Exact duplicates make K singular; with s = 5e-4 the factorisation still succeeds in the reference and here."""
    from gaussian_process_b200 import GP_regression as G
    X = np.repeat(np.linspace(-2, 2, 20)[:, None], 3, axis=0)
    y = np.sin(X).ravel()
    Xs = np.linspace(-2, 2, 9)[:, None]
    np.random.seed(1)
    mu, sd, _ = G.prediction(X, Xs, y, 'rbf', 1, 1)
    np.random.seed(1)
    mu_o, sd_o, _ = O.regression_prediction(X, Xs, y, 'rbf', 1, 1)
    assert rel(mu, mu_o) < 1e-7 and rel(sd ** 2, sd_o ** 2) < 1e-6


def test_wrong_kernel_choice_and_bad_theta_fail_like_the_reference():
    from gaussian_process_b200 import CO2_example as C2, GP_regression as G
    X, y, Xs = O.synth_c1(8, 10)
    with pytest.raises(UnboundLocalError):
        G.prediction(X, Xs, y, 'matern', 1, 1)
    with pytest.raises(IndexError):
        C2.covariance_function(X, X, [1.0, 2.0, 3.0])


def test_engine_reuse_across_sizes_is_stateless():
    from gaussian_process_b200 import tune_hyperparms_regression as T
    vals = []
    for N in (700, 64, 700):
        X, y = O.synth_c5(N, 16)
        vals.append(float(T.compute_mar_likelihood(X, None, y, 1.0, 4.0)))
    assert vals[0] == vals[2]     # bit-identical: deterministic kernels, no atomics on the path


def test_co2_square_block_rule_and_nd_inputs():
    """delta is added iff the block is square (CO2_example.py:60-63) -- also for two *different* inputs of equal length."""
    from gaussian_process_b200 import CO2_example as C2
    rs = np.random.RandomState(2)
    a, b = rs.rand(30, 11) * 5, rs.rand(30, 11) * 5
    th = O.CO2_THETA_BOOK
    assert rel(C2.covariance_function(a, b, th), O.co2_covariance(a, b, th)) < 1e-13
    c = rs.rand(17, 11) * 5
    assert rel(C2.covariance_function(a, c, th), O.co2_covariance(a, c, th)) < 1e-13
