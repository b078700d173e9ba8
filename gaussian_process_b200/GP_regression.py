"""Drop-in for the reference's ``GP_regression`` module, running on the gpx B200 engine.

Same function names, argument order and return tuples as /root/reference/GP_regression.py; NumPy
float64 in / out; ``numpy.linalg.LinAlgError`` when a Cholesky is not positive definite.  All linear
algebra (covariance build, Cholesky, triangular solves, predictive moments, sampling factor) runs in
libgpx's sm_100a kernels; normal draws come from the global NumPy RNG on the host *after* the linear
algebra, in the reference's call order, so seeded runs reproduce the reference's stream.
"""
from __future__ import annotations

import numpy as np

from ._lib import COV_LIN, COV_PER, COV_SE
from .engine import get_engine

NOISE_VARIANCE = 0.0005      # `s` of GP_regression.py:58,81,120
SAMPLING_JITTER = 1e-6       # GP_regression.py:154
FUSED_SMALL_PATH = True      # N <= 128 and n <= 128: one-launch posterior (csrc/small.cu); False forces the tiled path

# module globals the reference drivers read (GP_regression.py:105,286,295)
n = 100
kernel_stand_deiv = 1
true_fun = None


def _scalar(v) -> float:
    return float(np.asarray(v, dtype=np.float64).reshape(-1)[0])


def _kind_theta(kernel_choice, parameter, sigma=1.0):
    """(kind, theta) of the reference's kernel_choice dispatch (GP_regression.py:84-89,125-136)."""
    if kernel_choice == 'rbf':
        return COV_SE, [_scalar(sigma), _scalar(parameter)]
    if kernel_choice == 'lin':
        return COV_LIN, [_scalar(parameter)]
    if kernel_choice == 'per':
        p, l = parameter
        return COV_PER, [_scalar(p), _scalar(l)]
    # the reference leaves `kernel` unassigned for any other choice
    raise UnboundLocalError("local variable 'kernel' referenced before assignment")


def _cov_host(kind, a, b, theta):
    eng = get_engine()
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    K = eng.cov(kind, eng.to_device(a), eng.to_device(b), theta)
    return eng.to_host(K[:a.shape[0], :b.shape[0]])


def RBF_kernel(a, b, sigma, l):
    """sigma^2 exp(-|a_i - b_j|^2 / (2 l^2)) -> ndarray (len(a), len(b)).  GP_regression.py:8-19."""
    return _cov_host(COV_SE, a, b, [_scalar(sigma), _scalar(l)])


def lin_kernel(a, b, c):
    """(a - c).(b - c) -> ndarray.  GP_regression.py:22-33."""
    return _cov_host(COV_LIN, a, b, [_scalar(c)])


def per_kernel(a, b, parameters):
    """exp(-2 sin^2(pi |a-b| / p) / l^2) -> ndarray.  GP_regression.py:36-50."""
    p, l = parameters
    return _cov_host(COV_PER, a, b, [_scalar(p), _scalar(l)])


def dataset_generator(N, n):
    """sin(0.9 x) + noise on U(-5, 5); host-side data synthesis (GP_regression.py:53-68)."""
    f = lambda x: np.sin(0.9 * x).flatten()  # noqa: E731
    X_train = np.random.uniform(-5, 5, size=(N, 1))
    y_train = f(X_train) + np.sqrt(NOISE_VARIANCE) * np.random.randn(N)
    X_test = np.linspace(-5, 5, n).reshape(-1, 1)
    return f, X_train, y_train, X_test


def f_prior(X_test, mu_prior, kernel_choice, kernel_parameter, num_fun):
    """Prior draws mu + chol(K + s I) z.  GP_regression.py:71-92."""
    eng = get_engine()
    kind, theta = _kind_theta(kernel_choice, kernel_parameter, 1.0)
    X_test = np.asarray(X_test, dtype=np.float64)
    if FUSED_SMALL_PATH and num_fun >= 1 and X_test.shape[0] <= eng.small_max():
        m = eng.small_prior_factor(kind, X_test, theta, NOISE_VARIANCE)      # one launch; raises like :90
        return mu_prior + eng.small_sample(m, np.random.normal(size=(m, num_fun)))
    Xd = eng.to_device(X_test)
    m = Xd.shape[0]
    B = eng.cov(kind, Xd, Xd, theta, diag_add=NOISE_VARIANCE, same_x=True)
    eng.potrf(B)
    z = np.random.normal(size=(m, num_fun))
    return mu_prior + eng.tri_times(B, z, m)


def prior_process(X_test, kernel_choice, kernel_parameter, num_fun):
    """Zero-mean prior draws; like the reference it sizes the mean with the module global ``n``."""
    mu_prior = np.zeros((n, 1))
    return f_prior(X_test, mu_prior, kernel_choice, kernel_parameter, num_fun)


def _fit_predict_sample(kind, theta, s, X_train, X_test, y_train, num_fun):
    """fit -> (mu, sd, f_post) shared by every regression-style entry point
    (GP_regression.py:138-156; tune...:85-101; CO2...:198-214)."""
    eng = get_engine()
    X_train = np.asarray(X_train, dtype=np.float64)
    X_test = np.asarray(X_test, dtype=np.float64)
    if FUSED_SMALL_PATH and num_fun >= 1 and X_train.shape[0] <= eng.small_max() and X_test.shape[0] <= eng.small_max():
        # as-shipped sizes (N=5, n=100): the whole linear algebra is ONE kernel launch (csrc/small.cu) that leaves
        # the sampling factor on the device; the normals are drawn afterwards, as in the reference (:155 comes after
        # the Cholesky calls that may raise), and a second tiny launch forms mu + L_ z.
        mu_post, var, _ = eng.small_fit(kind, X_train, y_train, X_test, theta, s, SAMPLING_JITTER)
        with np.errstate(invalid="ignore"):
            stand_devi = np.sqrt(var)                       # NaN where var < 0, as in the reference
        z = np.random.normal(size=(X_test.shape[0], num_fun))
        return mu_post, stand_devi, eng.small_sample(X_test.shape[0], z), None
    fit = eng.fit(kind, X_train, y_train, theta, s)
    Xs = eng.to_device(np.asarray(X_test, dtype=np.float64))
    m = Xs.shape[0]
    mu, var, V = eng.predict(fit, Xs, want_v=True)
    mu_post = eng.to_host(mu)
    with np.errstate(invalid="ignore"):
        stand_devi = np.sqrt(eng.to_host(var))          # NaN where var < 0, as in the reference
    L_, _ = eng.posterior_sample_factor(kind, theta, Xs, V, SAMPLING_JITTER)
    z = np.random.normal(size=(m, num_fun))
    f_post_fun = mu_post.reshape(-1, 1) + eng.tri_times(L_, z, m)
    return mu_post, stand_devi, f_post_fun, fit


def prediction(X_train, X_test, y_train, kernel_choice, l, num_fun):
    """GP posterior mean / standard deviation / draws.  GP_regression.py:109-156."""
    kind, theta = _kind_theta(kernel_choice, l, 1.0)
    mu_post, stand_devi, f_post_fun, _ = _fit_predict_sample(kind, theta, NOISE_VARIANCE, X_train, X_test, y_train, num_fun)
    return mu_post, stand_devi, f_post_fun


# ----------------------------------------------------------------------------------------------
# plotting (host-side visualisation, out of the hot path): active only when matplotlib exists
# ----------------------------------------------------------------------------------------------
def _plt():
    try:
        import matplotlib.pyplot as plt
        return plt
    except Exception:  # matplotlib is not part of the engine's requirements
        return None


def plot_kernel(kernel_choice):
    plt = _plt()
    if plt is None:
        return
    xs = np.linspace(-3, 3, 200).reshape(-1, 1)
    zero = np.zeros((1, 1))
    if kernel_choice == 'rbf':
        k = RBF_kernel(xs, zero, kernel_stand_deiv, 1)
    elif kernel_choice == 'per':
        k = per_kernel(xs, zero, [1, 1])
    else:
        k = lin_kernel(xs, np.ones((1, 1)), 0)
    plt.figure()
    plt.plot(xs, k)
    plt.title(kernel_choice + ' kernel')


def plot_prior(X_test, f_prior_fun, stand_deiv):
    plt = _plt()
    if plt is None:
        return
    plt.figure()
    plt.plot(X_test, f_prior_fun)
    plt.gca().fill_between(X_test.flat, -3 * stand_deiv, 3 * stand_deiv, color="#dddddd")
    plt.title('samples from the GP prior')


def plot_posterior(X_test, f_post_fun, mu_post, stand_devi):
    plt = _plt()
    if plt is None:
        return
    plt.figure()
    plt.plot(X_test, f_post_fun)
    plt.gca().fill_between(X_test.flat, mu_post - 3 * stand_devi, mu_post + 3 * stand_devi, color="#dddddd")
    plt.plot(X_test, mu_post, 'r--', lw=2)
    plt.title('samples from the GP posterior')


def plot_true_diff(X_train, X_test, y_train, true_fun, mu_post, stand_devi):
    plt = _plt()
    if plt is None:
        return
    plt.figure()
    plt.plot(X_train, y_train, 'r+', ms=20)
    plt.plot(X_test, true_fun(X_test), 'b-')
    plt.gca().fill_between(X_test.flat, mu_post - 3 * stand_devi, mu_post + 3 * stand_devi, color="#dddddd")
    plt.plot(X_test, mu_post, 'r--', lw=2)
    plt.title('Mean predictions plus 3 st.deviations')


def GP_regression(X_train, y_train, X_test, num_fun, kernel_choice, kernel_parameter):
    """Driver: prior draws, posterior, plots (GP_regression.py:268-297).  Returns the posterior tuple."""
    plot_kernel(kernel_choice)
    f_prior_fun = prior_process(X_test, kernel_choice, kernel_parameter, num_fun)
    plot_prior(X_test, f_prior_fun, kernel_stand_deiv)
    mu_post, stand_devi, f_post_fun = prediction(X_train, X_test, y_train, kernel_choice, kernel_parameter, num_fun)
    plot_posterior(X_test, f_post_fun, mu_post, stand_devi)
    if true_fun is not None:
        plot_true_diff(X_train, X_test, y_train, true_fun, mu_post, stand_devi)
    plt = _plt()
    if plt is not None:
        plt.show()
    return mu_post, stand_devi, f_post_fun


if __name__ == "__main__":
    N = 5
    n = 100
    num_fun = 10
    kernel_parameter = 1
    kernel_stand_deiv = 1
    kernel_choice = 'rbf'
    true_fun, X_train, y_train, X_test = dataset_generator(N, n)
    mu, sd, _ = GP_regression(X_train, y_train, X_test, num_fun, kernel_choice, kernel_parameter)
    print("posterior mean range: [%.4f, %.4f], max sd %.4f" % (mu.min(), mu.max(), np.nanmax(sd)))
