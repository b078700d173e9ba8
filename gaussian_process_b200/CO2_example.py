"""Drop-in for the reference's ``CO2_example`` module on the gpx B200 engine.

Composite covariance k1 + k2 + k3 + k4 (SE + SE x periodic + rational quadratic + SE-noise + delta),
CO2_example.py:9-94, evaluated by libgpx's fused builder in one pass; fit / LML / prediction follow
CO2_example.py:131-214.  The Bayesian-optimisation driver around it is host-side Python.
"""
from __future__ import annotations

import random

import numpy as np
from scipy.stats import norm

from . import GP_regression as _gpr
from ._lib import COV_CO2
from .engine import get_engine
from .tune_hyperparms_regression import overlap

NOISE_VARIANCE = 0.0005   # CO2...:139,191
BO_NOISE = 0.0001         # CO2...:160
HYPERMS_BOOK = np.array([66, 67, 2.4, 90, 1.3, .66, 1.2, .78, .18, 1.6, .19])  # CO2...:117,324,368,418

mu_post = None            # module global read by plot_prediction (CO2...:395)


def _theta(hyperparms):
    t = np.asarray(hyperparms, dtype=np.float64).reshape(-1)
    if t.size != 11:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (min(t.size, 10), t.size))
    return t


def _co2_term(term, sqdist, l2_norm, t0, t1, t2):
    import ctypes
    eng = get_engine()
    d2 = np.ascontiguousarray(sqdist, dtype=np.float64)
    if d2.ndim != 2:
        raise ValueError("sqdist must be a matrix")
    D2 = eng.to_device(d2)
    R = eng.to_device(np.ascontiguousarray(l2_norm, dtype=np.float64)) if l2_norm is not None else None
    out = eng.empty(*d2.shape)
    eng._sync_stream()
    from ._lib import check
    check(eng.lib.gpx_co2_term(eng.h, term, d2.shape[0], d2.shape[1], eng._p(D2), d2.shape[1], eng._p(R), d2.shape[1],
                               float(t0), float(t1), float(t2), eng._p(out), d2.shape[1]), "gpx_co2_term")
    return eng.to_host(out)


def kernel_1(sqdist, theta_1, theta_2):
    """theta_1^2 exp(-.5 sqdist / theta_2^2)  (CO2...:9-17) as an element-wise device map."""
    return _co2_term(1, sqdist, None, theta_1, theta_2, 0.0)


def kernel_2(l2_norm, sqdist, theta_3, theta_4, theta_5):
    """theta_3^2 exp(-.5 sqdist/theta_4^2 - 2 (sin(pi l2_norm)/theta_5)^2)  (CO2...:20-32)."""
    return _co2_term(2, sqdist, l2_norm, theta_3, theta_4, theta_5)


def kernel_3(sqdist, theta_6, theta_7, theta_8):
    """theta_6^2 (1 + .5 sqdist/(theta_8 theta_7^2))^-theta_8  (CO2...:35-46)."""
    return _co2_term(3, sqdist, None, theta_6, theta_7, theta_8)


def kernel_4(sqdist, theta_9, theta_10, theta_11):
    """theta_9^2 exp(-.5 sqdist/theta_10^2) + theta_11^2 I for a square block  (CO2...:49-66)."""
    return _co2_term(4, sqdist, None, theta_9, theta_10, theta_11)


def covariance_function(a, b, hyperparms):
    """K = k1 + k2 + k3 + k4 for inputs of any dimension (CO2...:69-94); the theta_11^2 delta term is
    added iff the block is square (:60-63), exactly as in the reference."""
    eng = get_engine()
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    square = a.shape[0] == b.shape[0]
    K = eng.cov(COV_CO2, eng.to_device(a), eng.to_device(b), _theta(hyperparms), same_x=square)
    return eng.to_host(K[:a.shape[0], :b.shape[0]])


def compute_mar_likelihood(X_train, y_train, hyperparms):
    """log p(y | X, theta), s = 5e-4 (CO2...:131-149)."""
    eng = get_engine()
    if _gpr.FUSED_SMALL_PATH and len(X_train) <= eng.small_max():
        return np.float64(eng.small_lml_grad(COV_CO2, X_train, y_train, _theta(hyperparms), NOISE_VARIANCE, with_grad=False)[0])
    fit = eng.fit(COV_CO2, np.asarray(X_train, dtype=np.float64), y_train, _theta(hyperparms), NOISE_VARIANCE)
    return np.float64(fit.lml)


def compute_mar_likelihood_gradient(X_train, y_train, hyperparms):
    """(LML, dLML/dtheta[11]) -- the multi-theta generalisation of tune...:31-64."""
    eng = get_engine()
    if _gpr.FUSED_SMALL_PATH and len(X_train) <= eng.small_max():
        lml, grad = eng.small_lml_grad(COV_CO2, X_train, y_train, _theta(hyperparms), NOISE_VARIANCE)
        return np.float64(lml), grad
    fit = eng.fit(COV_CO2, np.asarray(X_train, dtype=np.float64), y_train, _theta(hyperparms), NOISE_VARIANCE, with_grad=True)
    return np.float64(fit.lml), fit.grad


def bayesian_opt(hyperparms_train, hyperparms_test, y_train):
    """GP over the 11-D theta space using theta_train[0] as the kernel's own hyper-parameters
    (CO2...:152-179) -> (mu_post, stand_devi)."""
    eng = get_engine()
    hp = _theta(np.asarray(hyperparms_train)[0])
    if _gpr.FUSED_SMALL_PATH and len(hyperparms_train) <= eng.small_max():
        mu, var, _, _ = eng.small_posterior(COV_CO2, hyperparms_train, y_train, hyperparms_test, hp, BO_NOISE, 0.0, None)
        with np.errstate(invalid="ignore"):
            return mu, np.sqrt(var)
    fit = eng.fit(COV_CO2, np.asarray(hyperparms_train, dtype=np.float64), y_train, hp, BO_NOISE)
    mu, var, _ = eng.predict(fit, np.asarray(hyperparms_test, dtype=np.float64))
    with np.errstate(invalid="ignore"):
        return eng.to_host(mu), np.sqrt(eng.to_host(var))


def make_prediction(X_train, X_test, y_train, hyperparms):
    """Posterior mean / sd / one draw at X_test (CO2...:182-214)."""
    mu, sd, f_post_fun, _ = _gpr._fit_predict_sample(COV_CO2, _theta(hyperparms), NOISE_VARIANCE, X_train, X_test, y_train, 1)
    return mu, sd, f_post_fun


# ----------------------------------------------------------------------------------------------
# Bayesian-optimisation driver (host side)
# ----------------------------------------------------------------------------------------------
def random_sample_test_parms(n_test_hyperparms, train_parms):
    """Per-dimension candidates on [0.3, 1.5] x book value minus visited ones (CO2...:109-128)."""
    dim_parms = 11
    lower = HYPERMS_BOOK - HYPERMS_BOOK * .7
    upper = HYPERMS_BOOK + HYPERMS_BOOK * .5
    out = np.zeros(shape=(n_test_hyperparms, dim_parms))
    for i in range(dim_parms):
        num_gen = n_test_hyperparms + len(train_parms) + 10
        cand = np.linspace(lower[i], upper[i], num_gen)
        _, ind_sample = overlap(train_parms[:, i], cand)
        cand = np.delete(cand, ind_sample)
        out[:, i] = np.asarray(random.sample(list(cand), n_test_hyperparms))
    return out


def UBC(hyperparms_train, hyperparms_test, mu_post, stand_devi):
    """Upper confidence bound, kappa = 7 (CO2...:217-236)."""
    objective = mu_post + 7 * stand_devi
    idx = np.asarray(np.where(objective == np.max(objective))[0])
    next_point = hyperparms_test[idx[0]]
    if np.array_equal(hyperparms_train[len(hyperparms_train) - 1], next_point):
        return True
    return next_point


def TS(hyperparms_train, hyperparms_test, y_train):
    """Thompson sampling; like the reference this unpacks three values from bayesian_opt's pair and so
    raises ValueError when reached (CO2...:239-250)."""
    mu_post, stand_devi, f_post_fun = bayesian_opt(hyperparms_train, hyperparms_test, y_train)
    max_index = np.where(f_post_fun == np.max(f_post_fun))
    return hyperparms_test[max_index[0]].flatten()


def EI(hyperparms_test, mu_post, stand_devi, y):
    """Expected improvement (CO2...:253-270)."""
    f_max = np.max(y) + 0.0005
    z = (mu_post - f_max) / stand_devi
    ei = (mu_post - f_max) * norm.cdf(z) + stand_devi * norm.pdf(z)
    return hyperparms_test[np.where(ei == np.max(ei))].flatten()


def PI(hyperparms_test, mu_post, stand_devi, y):
    """Probability of improvement (CO2...:273-293)."""
    f_max = np.max(y) + 0.0005
    cdf = norm.cdf((mu_post - f_max) / stand_devi)
    idx = np.asarray(np.where(cdf == np.max(cdf))[0])
    return hyperparms_test[idx[random.randint(0, len(idx) - 1)]]


def acquisition_fun(choice, hyperparms_train, hyperparms_test, mu_post, stand_devi, y):
    """Dispatch on the acquisition name (CO2...:296-314): 'UBC' | 'TS' | 'EI' | anything else -> PI."""
    if isinstance(choice, str) and choice == 'UBC':
        return UBC(hyperparms_train, hyperparms_test, mu_post, stand_devi)
    if isinstance(choice, str) and choice == 'TS':
        return TS(hyperparms_train, hyperparms_test, y)
    if isinstance(choice, str) and choice == 'EI':
        return EI(hyperparms_test, mu_post, stand_devi, y)
    return PI(hyperparms_test, mu_post, stand_devi, y)


def init_hyperms(n_hyperms, dim_parms):
    """Initial theta rows: book + 0.5 (i + 5) (CO2...:317-327)."""
    out = np.zeros(shape=(n_hyperms, dim_parms))
    for i in range(n_hyperms):
        out[i] = HYPERMS_BOOK + 0.5 * (i + 5)
    return out


def tune_hyperparameters_BO(X_train, X_test, y_train, num_iterations=10, n_hyperparms_test=500):
    """BO over the 11 hyper-parameters for each acquisition label (CO2...:330-379).  As shipped, the whole
    ``choice`` list is handed to ``acquisition_fun`` (:359) so every label runs PI; preserved."""
    dim_parms, n_train = 11, 5
    choice = ['UCB', 'TS', 'EI', 'PI']
    hyperparms_train = max_index = None
    for j in choice:
        print(j)
        hyperparms_train = init_hyperms(n_train, dim_parms)
        y_axis = np.zeros(num_iterations)
        for k in range(num_iterations):
            hyperparms_test = random_sample_test_parms(n_hyperparms_test, hyperparms_train)
            lml = np.array([compute_mar_likelihood(X_train, y_train, th) for th in hyperparms_train])
            mu_bo, sd_bo = bayesian_opt(hyperparms_train, hyperparms_test, lml)
            next_point = acquisition_fun(choice, hyperparms_train, hyperparms_test, mu_bo, sd_bo, lml)
            hyperparms_train = np.append(hyperparms_train, [next_point], axis=0)
            max_index = np.where(lml == np.max(lml))[0]
            print("**")
            print("the " + repr(k + 1) + "th iteration!")
            y_axis[k] = np.max(lml)
            print(y_axis[k])
            print(hyperparms_train[max_index][0])
        print("hyperms in the book is:")
        print(HYPERMS_BOOK)
        print("marginal log likelihood in the book is:")
        print(compute_mar_likelihood(X_train, y_train, HYPERMS_BOOK))
    return hyperparms_train[max_index][0]


def tune_hyperparameters_gradient(X_train, y_train, hyperparms, mask=None, step_size=1e-4, tolerance=1e-3, max_iter=100):
    """Gradient ascent on the log marginal likelihood over the 11 hyper-parameters (or the subset with mask[j] != 0): the
    gradient-based counterpart of tune_hyperparameters_BO (CO2...:330-379), same loop as tune...:121-153.  All state stays on
    the device, one iteration is one CUDA-graph launch (gpx_gp_ascent).
    Returns (theta, log marginal likelihood of the last iteration, iterations)."""
    eng = get_engine()
    th = _theta(hyperparms)
    mk = np.ones(11, dtype=np.int32) if mask is None else np.asarray(mask, dtype=np.int32)
    res = eng.ascend(COV_CO2, np.asarray(X_train, dtype=np.float64), y_train, th, mk, NOISE_VARIANCE, step_size, tolerance,
                     max_iter)
    return res["theta"], np.float64(res["lml"]), res["iterations"]


def synthetic_mauna_loa(N=468, seed=0):
    """Mauna-Loa-shaped monthly series (the reference's fetch_mldata source is gone)."""
    t = 1958 + np.arange(N) / 12.0
    y = (315 + 1.3 * (t - 1958) + 0.012 * (t - 1958) ** 2 + 3 * np.sin(2 * np.pi * t) + 0.8 * np.sin(4 * np.pi * t)
         + 0.3 * np.random.RandomState(seed).randn(N))
    return t[:, None], y


def plot_prediction(X_train, X_test, y_train, y_test, stand_devi):
    plt = _gpr._plt()
    if plt is None:
        return
    plt.clf()
    plt.plot(X_train.reshape(-1, 1), y_train, label='training data')
    plt.plot(X_test.reshape(-1, 1), y_test, label='test data')
    plt.gca().fill_between(X_test.flat, mu_post - 3 * stand_devi, mu_post + 3 * stand_devi, color="#dddddd")
    plt.xlim(X_train.min(), X_test.max())
    plt.xlabel("Year")
    plt.legend(loc=4)
    plt.show()


if __name__ == "__main__":
    X_train, y_train = synthetic_mauna_loa()
    X_test = np.arange(X_train.max() // 1 + 1, X_train.max() // 1 + 21, 1. / 12)[:, np.newaxis]
    empirical_mean = np.mean(y_train)
    y_train = y_train - empirical_mean
    hyperms = tune_hyperparameters_BO(X_train.reshape(-1, 1), X_test, y_train, num_iterations=3, n_hyperparms_test=100)
    mu_post, stand_devi, f_post_fun = make_prediction(X_train, X_test, y_train, hyperms)
    print("forecast range: %.2f .. %.2f ppm" % ((mu_post + empirical_mean).min(), (mu_post + empirical_mean).max()))
