"""Drop-in for the reference's ``tune_hyperparms_regression`` module on the gpx B200 engine.

Hot path (tune_hyperparms_regression.py:104-162, 292-313, 31-64): K build -> Cholesky -> solves ->
LML -> K^-1 -> dLML/dl, repeated by a gradient-ascent loop.  Here one iteration is one fused libgpx
call (``gpx_gp_fit_grad``) with all state resident in HBM; the Bayesian-optimisation layer around it
(acquisition functions, candidate sampling) is host-side scalar work and stays Python.
"""
from __future__ import annotations

import random

import numpy as np
from scipy.stats import norm

from . import GP_regression as _gpr
from ._lib import COV_SE, GpxError
from .engine import get_engine, padded
from .GP_regression import RBF_kernel, dataset_generator, plot_posterior, plot_true_diff, prediction  # noqa: F401

NOISE_VARIANCE = 0.0005   # tune...:115,302
BO_NOISE = 0.0001         # tune...:75
STEP_SIZE = 0.01          # tune...:42
TOLERANCE = 0.001         # tune...:117

true_fun = None           # read by tune_hyperparms_gradient (tune...:414)


def _r(x):
    """repr() of a scalar the way the py2 reference printed it (no ``np.float64(...)`` wrapper)."""
    return repr(float(np.asarray(x).reshape(-1)[0]))


def plot_BO(X_train, y_train, X_test, f_post_fun, mu_post, stand_devi):
    plt = _gpr._plt()
    if plt is None:
        return
    plt.subplot(2, 1, 1)
    plt.plot(X_train, y_train, 'r+', ms=20)
    plt.gca().fill_between(X_test.flat, mu_post - 3 * stand_devi, mu_post + 3 * stand_devi, color="#dddddd")
    plt.plot(X_test, mu_post, linewidth=1)
    plt.title('Bayesian Optimization')
    plt.show()


def lml_gradient(a, sigma, l, alpha, K_y):
    """.5 tr((alpha alpha^T - K_y) dK/dtheta) for theta = (sigma, l) via the fused gradient kernel.

    ``K_y`` is the dense *inverse* of K + sI (the reference's argument name), ``alpha`` = K_y y."""
    eng = get_engine()
    a = np.asarray(a, dtype=np.float64)
    N = a.shape[0]
    npad = padded(N)
    Kinv = eng.zeros(npad, npad)
    Kinv[:N, :N] = eng.to_device(np.asarray(K_y, dtype=np.float64))
    al = eng.zeros(npad)
    al[:N] = eng.to_device(np.asarray(alpha, dtype=np.float64).reshape(-1))
    Xd = eng.to_device(a)
    theta = [float(np.asarray(sigma).reshape(-1)[0]), float(np.asarray(l).reshape(-1)[0])]
    return eng.lml_grad(COV_SE, Xd, theta, Kinv, al, n=N)


def gradient_ascent(a, b, sigma, l, alpha, K_y):
    """One ascent step on the length-scale: l <- l + 0.01 * dLML/dl (sigma is held fixed, as in the
    reference where its update is commented out).  tune_hyperparms_regression.py:31-64."""
    if b is not a and not (np.shape(a) == np.shape(b) and np.array_equal(a, b)):
        raise ValueError("gradient_ascent: the engine evaluates the trace form on a square block (a must equal b, "
                         "as at the reference's only call site tune_hyperparms_regression.py:145)")
    l_var = lml_gradient(a, sigma, l, alpha, K_y)[1]
    l = l + STEP_SIZE * l_var
    return sigma, l


def bayesian_opt(X_train, X_test, y_train):
    """GP posterior over the 1-D hyper-parameter axis (sigma = l = 1, s = 1e-4).  tune...:67-101."""
    mu_post, stand_devi, f_post_fun, _ = _gpr._fit_predict_sample(COV_SE, [1.0, 1.0], BO_NOISE, X_train, X_test, y_train, 1)
    return mu_post, stand_devi, f_post_fun


def tune_hyperparms_first(X_train, X_test, y_train, num_fun, sigma, l):
    """Gradient ascent on the log marginal likelihood w.r.t. the length-scale until |dLML| <= 1e-3.
    tune_hyperparms_regression.py:104-162 -> (mu_post, stand_devi, f_post_fun, optimal_likelihood).

    Per iteration the device does: fused K build, blocked DMMA Cholesky, two TRSVs, LML reduction,
    in-place triangular inverse, one triangular SYRK for K^-1 and the fused trace kernel."""
    eng = get_engine()
    X_train = np.asarray(X_train, dtype=np.float64)
    X_test = np.asarray(X_test, dtype=np.float64)
    if _gpr.FUSED_SMALL_PATH and num_fun >= 1 and X_train.shape[0] <= eng.small_max() and X_test.shape[0] <= eng.small_max():
        return _tune_first_small(eng, X_train, X_test, y_train, num_fun, sigma, l)
    Xd = eng.to_device(X_train)
    yd = eng.to_device(np.asarray(y_train, dtype=np.float64).reshape(-1))
    Xs = eng.to_device(X_test)
    N = Xs.shape[0]
    sig = float(np.asarray(sigma).reshape(-1)[0])
    l0 = float(np.asarray(l).reshape(-1)[0])
    # tune...:121-153 as ONE libgpx call: theta, LML history and the convergence test stay on the device, one iteration
    # (K build, Cholesky, solves, LML, K^-1, fused gradient, step on l) is one CUDA-graph launch.  Only l moves (the
    # reference's sigma update is commented out, :46-62).
    res = eng.ascend(COV_SE, Xd, yd, [sig, l0], [0, 1], NOISE_VARIANCE, STEP_SIZE, TOLERANCE, 10000)
    l = np.full(np.shape(l), res["theta"][1]) if isinstance(l, np.ndarray) else res["theta"][1]
    if res["converged"]:
        print("The hyperparameter tuning function has already converged after " + repr(res["iterations"]) + " iterations!")
        print("The error is " + _r(res["error"]))
        print("training end!")
    optimal_likelihood = np.float64(res["lml"])
    print('optimal lenghscalar is: ' + _r(l))
    print('maximum log marginal likelihood is: ' + _r(optimal_likelihood))
    theta = [sig, float(res["theta_used"][1])]
    fit = eng.fit(COV_SE, Xd, yd, theta, NOISE_VARIANCE)                    # tune...:132-137 at the theta of the last iteration
    mu, var, V = eng.predict(fit, Xs, want_v=True)
    mu_post, var = eng.to_host(mu), eng.to_host(var)
    with np.errstate(invalid="ignore"):
        stand_devi = np.sqrt(var)
    L_, _ = eng.posterior_sample_factor(COV_SE, theta, Xs, V, 1e-6)         # tune...:159
    f_post_fun = mu_post.reshape(-1, 1) + eng.tri_times(L_, np.random.normal(size=(N, num_fun)), N)
    return mu_post, stand_devi, f_post_fun, optimal_likelihood


def tune_hyperparms_all(X_train, y_train, sigma, l, step_size=STEP_SIZE, tolerance=TOLERANCE, max_iter=10000):
    """Gradient ascent on BOTH sigma and l (the update the reference leaves commented out, tune...:46-62, switched on):
    returns (sigma, l, log marginal likelihood, iterations).  Device-resident loop (gpx_gp_ascent)."""
    eng = get_engine()
    res = eng.ascend(COV_SE, np.asarray(X_train, dtype=np.float64), y_train,
                     [float(np.asarray(sigma).reshape(-1)[0]), float(np.asarray(l).reshape(-1)[0])], [1, 1], NOISE_VARIANCE,
                     step_size, tolerance, max_iter)
    return res["theta"][0], res["theta"][1], np.float64(res["lml"]), res["iterations"]


def _tune_first_small(eng, X_train, X_test, y_train, num_fun, sigma, l):
    """As-shipped sizes (N <= 128): the whole ascent loop is ONE kernel launch with all state in shared memory
    (csrc/small.cu, gp_small_grad_kernel), then the one-launch posterior at the length-scale of the last iteration."""
    sig = float(np.asarray(sigma).reshape(-1)[0])
    res = eng.small_ascent(X_train, y_train, sig, float(np.asarray(l).reshape(-1)[0]), NOISE_VARIANCE, STEP_SIZE,
                           TOLERANCE, 10000)
    l = np.full(np.shape(l), res["l"]) if isinstance(l, np.ndarray) else res["l"]
    if res["converged"]:
        print("The hyperparameter tuning function has already converged after " + repr(res["iterations"]) + " iterations!")
        print("The error is " + _r(res["error"]))
        print("training end!")
    optimal_likelihood = np.float64(res["lml"])
    print('optimal lenghscalar is: ' + _r(l))
    print('maximum log marginal likelihood is: ' + _r(optimal_likelihood))
    mu_post, var, _ = eng.small_fit(COV_SE, X_train, y_train, X_test, [sig, res["l_used"]], NOISE_VARIANCE, 1e-6)
    with np.errstate(invalid="ignore"):
        stand_devi = np.sqrt(var)
    z = np.random.normal(size=(X_test.shape[0], num_fun))                   # tune...:160
    return mu_post, stand_devi, eng.small_sample(X_test.shape[0], z), optimal_likelihood


def compute_mar_likelihood(X_train, X_test, y_train, sigma, l):
    """log p(y | X, sigma, l) with s = 5e-4.  tune_hyperparms_regression.py:292-313 (X_test unused)."""
    eng = get_engine()
    theta = [float(np.asarray(sigma).reshape(-1)[0]), float(np.asarray(l).reshape(-1)[0])]
    if _gpr.FUSED_SMALL_PATH and len(X_train) <= eng.small_max():
        return np.float64(eng.small_lml_grad(COV_SE, X_train, y_train, theta, NOISE_VARIANCE, with_grad=False)[0])
    fit = eng.fit(COV_SE, np.asarray(X_train, dtype=np.float64), y_train, theta, NOISE_VARIANCE)
    return np.float64(fit.lml)


# ----------------------------------------------------------------------------------------------
# Bayesian-optimisation layer (host-side).  py3 fixes: random.sample on a list,
# integer index arrays for np.delete.
# ----------------------------------------------------------------------------------------------
def PI(params, means, stand_devi, parms_done, y, n_iterations, k):
    """Probability of improvement (tune...:165-203); returns the next point or True to stop."""
    s = 0.0005
    stop_threshold = 0.001
    f_max = np.max(y) + s
    z = (means - f_max) / stand_devi
    cumu_gaussian = norm.cdf(z)
    if cumu_gaussian.sum() <= stop_threshold or np.max(cumu_gaussian) <= stop_threshold:
        print("all elements of cumulative are alost zeros!!!")
        return True
    indices = np.asarray(np.where(cumu_gaussian == np.max(cumu_gaussian))[0])
    done = np.asarray(parms_done).tolist()
    next_point = params[indices[random.randint(0, len(indices) - 1)]]
    condition = next_point in done
    while condition:
        next_point = params[indices[random.randint(0, len(indices) - 1)]]
        condition = next_point in done
        if len(next_point) == 1 and condition:
            return True
    return next_point


def UCB(parms_done, params, means, stand_devi, n_iterations, k):
    """Upper confidence bound, kappa = 1e-3 (tune...:206-229)."""
    kappa = 0.001
    objective = means + kappa * stand_devi
    indices = np.asarray(np.where(objective == np.max(objective))[0])
    next_point = params[indices[0]]
    if parms_done[len(parms_done) - 1] == next_point:
        return True
    return next_point


def TS(parms_done, params, y, n_iterations, k):
    """Thompson sampling: argmax of one posterior draw (tune...:232-248)."""
    mu_post, stand_devi, f_post_fun = prediction(np.asarray(parms_done).reshape(-1, 1), params, y, 'rbf', 1, 1)
    max_index = np.where(f_post_fun == np.max(f_post_fun))
    return params[max_index]


def EI(params, means, stand_devi, parms_done, y, n_iterations, k):
    """Expected improvement (tune...:251-273)."""
    f_max = np.max(y) + 0.0005
    z = (means - f_max) / stand_devi
    EI_vector = (means - f_max) * norm.cdf(z) + stand_devi * norm.pdf(z)
    return params[np.where(EI_vector == np.max(EI_vector))]


def acquisition_fun(params, means, stand_devi, parms_done, y, n_iterations, k):
    """Evaluates all four acquisition functions and returns PI's choice, like tune...:275-289."""
    next_point_PI = PI(params, means, stand_devi, parms_done, y, n_iterations, k)
    UCB(parms_done, params, means, stand_devi, n_iterations, k)
    TS(parms_done, params, y, n_iterations, k)
    EI(params, means, stand_devi, parms_done, y, n_iterations, k)
    return next_point_PI


def overlap(a, b):
    """Indices of a's entries that occur in b, and where (tune...:316-329); integer dtypes throughout."""
    a = np.asarray(a)
    b = np.asarray(b)
    ind_a = np.arange(len(a))[np.isin(a, b)]
    ind_b = np.array([np.argwhere(b == a[x]) for x in ind_a], dtype=np.int64).flatten()
    return ind_a, ind_b


def random_gen_test_parms(n, parms_done):
    """n sorted candidate length-scales on linspace(0.01, 5) minus the visited ones (tune...:331-346)."""
    num_gen = n + len(parms_done) + 10
    test_parms = np.linspace(0.01, 5, num_gen)
    _, ind_sample = overlap(parms_done, test_parms)
    test_parms = np.delete(test_parms, ind_sample)
    sampled = np.asarray(random.sample(list(test_parms), n))
    return np.sort(sampled).reshape(-1, 1)


def tune_hyperparms_second(X_train, X_test, y_train, num_fun, sigma, l):
    """Bayesian optimisation of the length-scale, 3 rounds (tune...:349-395)."""
    n = 100
    n_iterations = 3
    max_index = np.array([0])
    k = 0
    for k in range(n_iterations):
        l_test = random_gen_test_parms(n, l)
        log_marg_likelihood = np.array([compute_mar_likelihood(X_train, X_test, y_train, sigma, li) for li in l])
        mu_post, stand_devi, f_post_fun = bayesian_opt(l.reshape(-1, 1), l_test, log_marg_likelihood)
        next_point = acquisition_fun(l_test, mu_post, stand_devi, l, log_marg_likelihood, n_iterations, k)
        if next_point is True:
            max_index = np.where(log_marg_likelihood == np.max(log_marg_likelihood))[0]
            print("it takes " + repr(k + 1) + " iterations to get the optimal!")
            print("optimal lenghscalar is:" + _r(l[max_index][0]))
            break
        l = np.append(l, next_point)
        max_index = np.where(log_marg_likelihood == np.max(log_marg_likelihood))[0]
    log_marg_likelihood = np.array([compute_mar_likelihood(X_train, X_test, y_train, sigma, li) for li in l])
    print("it takes " + repr(k + 1) + " iterations to get the optimal!")
    print("optimal lenghscalar is:" + _r(l[max_index][0]))
    print("maximum likelihood is:" + _r(np.max(log_marg_likelihood)))
    l_test = random_gen_test_parms(n, l)
    mu_post, stand_devi, f_post_fun = bayesian_opt(l.reshape(-1, 1), l_test, log_marg_likelihood)
    plot_BO(l.reshape(-1, 1), log_marg_likelihood, l_test, f_post_fun, mu_post, stand_devi)
    return np.max(log_marg_likelihood)


def tune_hyperparms_gradient(X_train, X_test, y_train, num_fun):
    """Random initial l in (0, 5), then gradient ascent (tune...:398-415)."""
    sigma = 1
    l = np.random.uniform(0, 5, 1)
    mu_post, stand_devi, f_post_fun, optimal_likelihood = tune_hyperparms_first(X_train, X_test, y_train, num_fun, sigma, l)
    plot_posterior(X_test, f_post_fun, mu_post, stand_devi)
    if true_fun is not None:
        plot_true_diff(X_train, X_test, y_train, true_fun, mu_post, stand_devi)
    return optimal_likelihood


def tune_hyperparms_BO(X_train, X_test, y_train, num_fun):
    """Two random initial length-scales, then Bayesian optimisation (tune...:418-432)."""
    sigma = 1
    l = np.random.uniform(0.02, 5, 2)
    return tune_hyperparms_second(X_train, X_test, y_train, num_fun, sigma, l)


if __name__ == "__main__":
    N, n, num_fun = 3, 100, 10
    true_fun, X_train, y_train, X_test = dataset_generator(N, n)
    print("")
    print("------ Bayesian oprimization ------")
    optimal_likelihood_BO = tune_hyperparms_BO(X_train, X_test, y_train, num_fun)
    print("")
    print("------ gradient ascent------")
    optimal_likelihood_GA = tune_hyperparms_gradient(X_train, X_test, y_train, num_fun)
    error = np.abs(optimal_likelihood_BO - optimal_likelihood_GA) / max(np.abs(optimal_likelihood_BO), np.abs(optimal_likelihood_GA))
    print("")
    print("------ error rate ------")
    print("The error rate of optimal likelihood between two methods is: %.3f%%" % (error * 100))
