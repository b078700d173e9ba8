"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of the reference's exact-GP hot path.

This file is the *oracle*: a plain NumPy float64 restatement of the arithmetic of
happyjin/Gaussian_process for the path SURVEY.md section 8 names.  It is NOT part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the CPU arm being timed.  The product
package (``gaussian_process_b200``) never imports anything from ``oracle/`` and fails loudly when
its CUDA library is missing.

Pinning: the reference ships no tests or golden vectors.  This restatement is pinned by
(i) ``tests/test_oracle.py::test_oracle_matches_live_reference_*`` which run the *unmodified*
reference functions (through ``oracle/ref_loader.py``) against the functions here on seeded inputs
(build container only: /root/reference does not exist on the GPU box) and (ii) ``tests/golden/*.npz``
generated from the reference by ``oracle/gen_golden.py`` (committed; checked everywhere).

Every function cites the reference file:line it follows.  Linear algebra deliberately uses the same
NumPy entry points as the reference (``np.linalg.cholesky / solve / inv``, ``np.dot``) so that the
CPU timing of this port is the timing of the reference's path; the only deviation is that pairwise
squared distances are produced in row chunks (same formula, same summation order per element) so
that the N*D*N temporary of ``GP_regression.py:18`` does not have to exist at large N.
"""
from __future__ import annotations

import numpy as np
from scipy.special import expit

S_NOISE = 0.0005          # GP_regression.py:58,81,120; tune...:115,302; CO2...:139,191
S_NOISE_BO = 0.0001       # tune...:75; CO2...:160
JITTER = 1e-6             # GP_regression.py:154; tune...:98,159; CO2...:212
CO2_THETA_BOOK = np.array([66, 67, 2.4, 90, 1.3, .66, 1.2, .78, .18, 1.6, .19])  # CO2...:117


# --------------------------------------------------------------------------------------------
# A1/A2  covariance functions  (GP_regression.py:8-50)
# --------------------------------------------------------------------------------------------
def sqdist(a, b, chunk=None):
    """sum_d (a_id - b_jd)^2, GP_regression.py:18 (summation over d in index order).

    ``chunk`` rows of ``a`` are processed at a time; per element the arithmetic is identical to the
    reference's broadcast expression."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    n, m = a.shape[0], b.shape[0]
    if chunk is None:
        chunk = max(1, min(n, int(2 ** 27 // max(1, m * a.shape[1]))))
    out = np.empty((n, m))
    bt = b[:, :, None].T  # (1, D, m)
    for i0 in range(0, n, chunk):
        out[i0:i0 + chunk] = ((a[i0:i0 + chunk, :, None] - bt) ** 2).sum(1)
    return out


def rbf_kernel(a, b, sigma, l, chunk=None):
    """sigma^2 exp(-.5 (1/l^2) sqdist), GP_regression.py:8-19."""
    return (sigma ** 2) * np.exp(-.5 * (1 / (l ** 2)) * sqdist(a, b, chunk))


def lin_kernel(a, b, c):
    """(a - c) . (b^T - c), GP_regression.py:22-33."""
    return 0 + 1 * np.dot(a - c, b.T - c)


def per_kernel(a, b, parameters):
    """exp(-2 sin^2(pi |a-b| / p) / l^2) for 1-D inputs, GP_regression.py:36-50."""
    p, l = parameters
    r = np.absolute(np.tile(a, (1, len(b))) - np.tile(b.T, (len(a), 1)))
    return 1 * np.exp(-2 * (np.sin(np.pi * r / p)) ** 2 / l ** 2)


def kernel_by_choice(a, b, kernel_choice, parameter, sigma=1):
    """Dispatch of GP_regression.py:84-89 / :125-136."""
    if kernel_choice == 'rbf':
        return rbf_kernel(a, b, sigma, parameter)
    if kernel_choice == 'lin':
        return lin_kernel(a, b, parameter)
    if kernel_choice == 'per':
        return per_kernel(a, b, parameter)
    raise UnboundLocalError("kernel")  # the reference leaves `kernel` unbound


# --------------------------------------------------------------------------------------------
# A3  CO2 composite kernel  (CO2_example.py:9-94)
# --------------------------------------------------------------------------------------------
def co2_covariance(a, b, theta, chunk=None):
    """k1 + k2 + k3 + k4 of CO2_example.py:69-94 (theta = 11 hyper-parameters).

    The 1-D branch (:79) and the N-D branch (:86) produce the same numbers (both sum squared
    differences over d in order); delta (theta_11 term) is added iff the block is square (:60-63)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    t = np.asarray(theta, dtype=np.float64)
    d = sqdist(a, b, chunk)
    r = np.sqrt(d)
    k1 = (t[0] ** 2) * np.exp(-.5 * d / t[1] ** 2)                                  # :17
    k2 = t[2] ** 2 * np.exp(-.5 * d / t[3] ** 2 + -2 * ((np.sin(np.pi * r)) / t[4]) ** 2)  # :30-32
    k3 = t[5] ** 2 * (1.0 / np.power(1 + .5 * d / (t[7] * t[6] ** 2), t[7]))        # :44-46
    delta = np.eye(len(d)) if d.shape[0] == d.shape[1] else 0                        # :58-63
    k4 = t[8] ** 2 * np.exp(-.5 * d / t[9] ** 2) + t[10] ** 2 * delta               # :65-66
    return k1 + k2 + k3 + k4                                                         # :90-93


# --------------------------------------------------------------------------------------------
# A4-A7, A9  regression fit / predict / LML
# --------------------------------------------------------------------------------------------
def _fit_solve(K_train, y_train, s):
    """chol + two LU solves, GP_regression.py:138-140 (== tune...:127-129, :307-309)."""
    n = len(K_train)
    L = np.linalg.cholesky(K_train + s * np.eye(n))
    m = np.linalg.solve(L, y_train)
    alpha = np.linalg.solve(L.T, m)
    return L, alpha


def lml_from(L, y_train, alpha):
    """-.5 y^T alpha - sum log diag L - n/2 log 2 pi, tune...:141,312; CO2...:148."""
    n = len(y_train)
    return -.5 * np.dot(y_train.T, alpha) - np.log(np.diagonal(L)).sum(0) - n / 2.0 * np.log(2 * np.pi)


def _post_sample(K_ss, v, mu_post, num_fun):
    """chol(K_ss + 1e-6 I - v^T v) then mu + L_ z with z from the global NumPy RNG,
    GP_regression.py:154-155 (== tune...:98-99,159-160; CO2...:212-213)."""
    n = len(K_ss)
    L_ = np.linalg.cholesky(K_ss + JITTER * np.eye(n) - np.dot(v.T, v))
    return mu_post.reshape(-1, 1) + np.dot(L_, np.random.normal(size=(n, num_fun)))


def regression_prediction(X_train, X_test, y_train, kernel_choice, l, num_fun, s=S_NOISE, sigma=1):
    """GP_regression.py:109-156 -> (mu_post[n], stand_devi[n], f_post_fun[n,num_fun])."""
    K_train = kernel_by_choice(X_train, X_train, kernel_choice, l, sigma)
    K_s = kernel_by_choice(X_train, X_test, kernel_choice, l, sigma)
    K_ss = kernel_by_choice(X_test, X_test, kernel_choice, l, sigma)
    L, alpha = _fit_solve(K_train, y_train, s)
    mu_post = np.dot(K_s.T, alpha)                       # :143
    v = np.linalg.solve(L, K_s)                          # :144
    var_test = np.diag(K_ss) - np.sum(v ** 2, axis=0)    # :147
    stand_devi = np.sqrt(var_test)                       # :148
    return mu_post, stand_devi, _post_sample(K_ss, v, mu_post, num_fun)


def f_prior(X_test, mu_prior, kernel_choice, kernel_parameter, num_fun):
    """Prior draw chol(K + s I) z, GP_regression.py:71-92."""
    K = kernel_by_choice(X_test, X_test, kernel_choice, kernel_parameter, 1)
    B = np.linalg.cholesky(K + S_NOISE * np.eye(len(X_test)))
    return mu_prior + np.dot(B, np.random.normal(size=(len(X_test), num_fun)))


def rbf_lml(X_train, y_train, sigma, l, s=S_NOISE, chunk=None):
    """tune_hyperparms_regression.py:292-313 (X_test is unused there)."""
    K_train = rbf_kernel(X_train, X_train, sigma, l, chunk)
    L, alpha = _fit_solve(K_train, y_train, s)
    return lml_from(L, y_train, alpha)


def bo_posterior_1d(X_train, X_test, y_train):
    """tune_hyperparms_regression.py:67-101: GP over the 1-D hyper-parameter axis (sigma=l=1,
    s=1e-4) -> (mu, sd, f_post[n,1])."""
    K = rbf_kernel(X_train, X_train, 1, 1)
    K_s = rbf_kernel(X_train, X_test, 1, 1)
    K_ss = rbf_kernel(X_test, X_test, 1, 1)
    L, alpha = _fit_solve(K, y_train, S_NOISE_BO)
    mu_post = np.dot(K_s.T, alpha)
    v = np.linalg.solve(L, K_s)
    stand_devi = np.sqrt(np.diag(K_ss) - np.sum(v ** 2, axis=0))
    return mu_post, stand_devi, _post_sample(K_ss, v, mu_post, 1)


# --------------------------------------------------------------------------------------------
# A8  LML gradient / gradient-ascent step (tune_hyperparms_regression.py:31-64, :144)
# --------------------------------------------------------------------------------------------
def rbf_grad_l(X, sigma, l, alpha, K_y_inv, chunk=None):
    """.5 tr((alpha alpha^T - K_y^-1) dK/dl) exactly as tune...:54-57 (full N^3 GEMM + trace)."""
    d = sqdist(X, X, chunk)
    l_grad = sigma ** 2 * np.exp(-.5 * d / (l ** 2)) * (d / l ** 3)      # :54
    a = np.asarray(alpha).reshape(-1, 1)
    l_matrix = np.dot(np.dot(a, a.T) - K_y_inv, l_grad)                   # :55
    return .5 * np.diagonal(l_matrix).sum()                               # :56-57


def gradient_ascent_step(X, sigma, l, alpha, K_y_inv, step_size=0.01):
    """tune...:31-64 -> (sigma, l + 0.01 * grad)."""
    return sigma, l + step_size * rbf_grad_l(X, sigma, l, alpha, K_y_inv)


def rbf_fit_lml_grad(X_train, y_train, sigma, l, s=S_NOISE, chunk=None):
    """One iteration body of tune_hyperparms_first restricted to the train-side operations
    (tune...:123,127-129,141,144-145): returns (lml, dLML/dl, alpha).  This is the unit the
    headline metric times (fit + LML + gradient)."""
    K_train = rbf_kernel(X_train, X_train, sigma, l, chunk)
    L, alpha = _fit_solve(K_train, y_train, s)
    lml = lml_from(L, y_train, alpha)
    K_y_inv = np.dot(np.linalg.inv(L.T), np.linalg.inv(L))               # :144
    return lml, rbf_grad_l(X_train, sigma, l, alpha, K_y_inv, chunk), alpha


def rbf_fit_lml_grad_best_effort(X_train, y_train, sigma, l, s=S_NOISE, chunk=2048, timings=None):
    """Memory-feasible "best-effort CPU" variant of the same unit: the same quantities as
    `rbf_fit_lml_grad` (tune...:123-145) computed the LAPACK-aware way -- chunked kernel, dpotrf,
    triangular dpotrs, dpotri and the O(N^2) trace  .5 * sum((alpha alpha^T - K^-1) o dK/dl)  instead of
    the reference's LU solves, two dense inverses and the N^3 GEMM of tune...:55.  It is a timing baseline
    only (bench.py `cpu_best_effort`); parity is always taken against `rbf_fit_lml_grad`.  `timings`, when a
    dict, receives the seconds of the O(N^2 D) part ("n2") and of the O(N^3) part ("n3") so that a bounded
    sample can be scaled term by term."""
    import time
    from scipy.linalg import cho_factor, cho_solve, lapack
    n = len(X_train)
    t0 = time.perf_counter()
    d = sqdist(X_train, X_train, chunk)
    K = sigma ** 2 * np.exp(-.5 * d / (l ** 2))
    dK = K * (d / l ** 3)                                                 # tune...:54
    del d
    K[np.diag_indices(n)] += s
    t1 = time.perf_counter()
    c, low = cho_factor(K, lower=True, overwrite_a=True, check_finite=False)
    alpha = cho_solve((c, low), y_train, check_finite=False)
    lml = -.5 * np.dot(y_train, alpha) - np.log(np.diagonal(c)).sum() - n / 2.0 * np.log(2 * np.pi)
    Kinv, info = lapack.dpotri(c, lower=1, overwrite_c=1)
    if info != 0:
        raise np.linalg.LinAlgError("dpotri info=%d" % info)
    t2 = time.perf_counter()
    # only the lower triangle of Kinv is valid: tr(Kinv dK) = 2 * sum(tril(Kinv o dK), -1) + diag part
    prod = np.tril(Kinv) * dK
    tr_kinv = 2.0 * prod.sum() - np.diagonal(prod).sum()
    grad = .5 * (np.dot(alpha, np.dot(dK, alpha)) - tr_kinv)
    if timings is not None:
        timings["n3"] = t2 - t1
        timings["n2"] = (t1 - t0) + (time.perf_counter() - t2)
    return lml, grad, alpha


def tune_first(X_train, X_test, y_train, num_fun, sigma, l, max_iter=10000, tolerance=0.001):
    """tune_hyperparms_regression.py:104-162 without the prints: gradient ascent on l until
    |dLML| <= 1e-3 -> (mu, sd, f_post, lml, l, iterations)."""
    s = S_NOISE
    old = 0
    it = 0
    for i in range(max_iter):
        K_train = rbf_kernel(X_train, X_train, sigma, l)
        K_s = rbf_kernel(X_train, X_test, sigma, l)
        K_ss = rbf_kernel(X_test, X_test, sigma, l)
        L, alpha = _fit_solve(K_train, y_train, s)
        mu_post = np.dot(K_s.T, alpha)
        v = np.linalg.solve(L, K_s)
        stand_devi = np.sqrt(np.diag(K_ss) - np.sum(v ** 2, axis=0))
        lml = lml_from(L, y_train, alpha)
        K_y_inv = np.dot(np.linalg.inv(L.T), np.linalg.inv(L))
        sigma, l = gradient_ascent_step(X_train, sigma, l, alpha.reshape(-1, 1), K_y_inv)
        error = np.sqrt(np.sum((lml - old) ** 2))
        old = lml
        it = i + 1
        if error <= tolerance:
            break
    f_post = _post_sample(K_ss, v, mu_post, num_fun)
    return mu_post, stand_devi, f_post, lml, l, it


# --------------------------------------------------------------------------------------------
# CO2 composite-kernel path (CO2_example.py:131-214)
# --------------------------------------------------------------------------------------------
def _co2_fit(K_train, y_train, s):
    """chol, explicit inv(L), alpha = L^-T (L^-1 y), CO2...:143-145."""
    L = np.linalg.cholesky(K_train + s * np.eye(len(K_train)))
    L_inv = np.linalg.inv(L)
    alpha = np.dot(L_inv.T, np.dot(L_inv, y_train))
    return L, L_inv, alpha


def co2_lml(X_train, y_train, theta, s=S_NOISE, chunk=None):
    """CO2_example.py:131-149."""
    L, _, alpha = _co2_fit(co2_covariance(X_train, X_train, theta, chunk), y_train, s)
    return lml_from(L, y_train, alpha)


def co2_bo_posterior(theta_train, theta_test, y_train):
    """CO2_example.py:152-179: GP over the 11-D theta space with hyper-parameters theta_train[0]."""
    hp = theta_train[0]
    K = co2_covariance(theta_train, theta_train, hp)
    K_s = co2_covariance(theta_train, theta_test, hp)
    K_ss = co2_covariance(theta_test, theta_test, hp)
    L, L_inv, alpha = _co2_fit(K, y_train, S_NOISE_BO)
    mu_post = np.dot(K_s.T, alpha)
    v = np.dot(L_inv, K_s)
    return mu_post, np.sqrt(np.diag(K_ss) - np.sum(v ** 2, axis=0))


def co2_make_prediction(X_train, X_test, y_train, theta, s=S_NOISE):
    """CO2_example.py:182-214 -> (mu, sd, f_post[n,1])."""
    K_train = co2_covariance(X_train, X_train, theta)
    K_s = co2_covariance(X_train, X_test, theta)
    K_ss = co2_covariance(X_test, X_test, theta)
    L, L_inv, alpha = _co2_fit(K_train, y_train, s)
    mu_post = np.dot(K_s.T, alpha)
    v = np.dot(L_inv, K_s)
    stand_devi = np.sqrt(np.diag(K_ss) - np.sum(v ** 2, axis=0))
    return mu_post, stand_devi, _post_sample(K_ss, v, mu_post, 1)


def co2_dcov(a, theta):
    """dK/dtheta_j (j = 0..10) of the composite kernel on a square block -- SURVEY Appendix C,
    derived from CO2_example.py:17,30-32,44-46,65-66.  NOT in the reference (it tunes theta by
    Bayesian optimisation only): pinned by central differences of ``co2_lml`` in the tests."""
    t = np.asarray(theta, dtype=np.float64)
    d = sqdist(a, a)
    r = np.sqrt(d)
    e1 = np.exp(-.5 * d / t[1] ** 2)
    s2 = np.sin(np.pi * r) ** 2
    k2 = t[2] ** 2 * np.exp(-.5 * d / t[3] ** 2 - 2 * s2 / t[4] ** 2)
    u = 1 + .5 * d / (t[7] * t[6] ** 2)
    k3 = t[5] ** 2 * np.power(u, -t[7])
    e4 = np.exp(-.5 * d / t[9] ** 2)
    return [
        2 * t[0] * e1, t[0] ** 2 * e1 * d / t[1] ** 3,
        2 * k2 / t[2], k2 * d / t[3] ** 3, k2 * 4 * s2 / t[4] ** 3,
        2 * k3 / t[5], k3 * d / (t[6] ** 3 * u), k3 * (-np.log(u) + (u - 1) / u),
        2 * t[8] * e4, t[8] ** 2 * e4 * d / t[9] ** 3, 2 * t[10] * np.eye(len(d)),
    ]


def lml_grad_from(alpha, K_y_inv, dKs):
    """d LML / d theta_j = .5 sum_ik (alpha alpha^T - K_y^-1)_ik (dK_j)_ik (R&W eq. 5.9, the
    Hadamard-sum form of tune...:55-57)."""
    a = np.asarray(alpha).reshape(-1)
    return np.array([.5 * (a @ dK @ a - np.sum(K_y_inv * dK)) for dK in dKs])


def rbf_dcov(X, sigma, l):
    """[dK/dsigma, dK/dl] for the SE kernel: tune...:48 (commented out in the reference) and :54."""
    d = sqdist(X, X)
    e = np.exp(-.5 * d / (l ** 2))
    return [2 * sigma * e, sigma ** 2 * e * (d / l ** 3)]


# --------------------------------------------------------------------------------------------
# A10  binary Laplace  (GP_binary_classification.py:48-154)
# --------------------------------------------------------------------------------------------
def pi_function(f):
    """GP_binary...:48-54."""
    return expit(f)


def deriv_log_likelihood(y, f):
    """t - sigmoid(y f) with t=(y+1)/2, GP_binary...:66-74 (note: y=-1 gives -sigmoid(-f))."""
    return (y + 1) / 2 - pi_function(y * f)


def sec_deriv_log_likelihood(f):
    """-pi (1 - pi), GP_binary...:77-83."""
    return -pi_function(f) * (1 - pi_function(f))


def binary_training_reference(K, y_train, f_prior_, num_funs=1, tolerance=0.0001, max_iter=10000):
    """Reference-faithful GP_binary...:86-133: gradient and W are evaluated at ``f_prior`` on every
    iteration (never at f), so W, B, L are constant and the iterate converges linearly.

    Returns (W dense, L_inv dense, first_deri (N,1), f (N,num_funs), errors list)."""
    n = y_train.size
    W = np.zeros((n, n))
    f = np.zeros((n, num_funs))
    errors = []
    for _ in range(max_iter):
        first_deri = deriv_log_likelihood(y_train, f_prior_)                       # :104
        np.fill_diagonal(W, -sec_deriv_log_likelihood(f_prior_))                   # :105
        sW = np.sqrt(W)
        L = np.linalg.cholesky(np.eye(n) + np.dot(np.dot(sW, K), sW))              # :107
        L_inv = np.linalg.inv(L)                                                   # :108
        b = np.dot(W, f) + first_deri                                              # :109
        a = b - np.dot(sW, np.dot(L_inv.T, np.dot(L_inv, np.dot(np.dot(sW, K), b))))  # :110
        f_new = np.dot(K, a)                                                       # :111
        err = np.sqrt(np.sum((f_new - f) ** 2))                                    # :113
        errors.append(err)
        f = f_new
        if err <= tolerance:
            break
    return W, L_inv, first_deri, f, errors


def binary_training_newton(K, y_train, f0=None, tolerance=1e-10, max_iter=100):
    """Textbook mode (R&W Alg. 3.1): the body of GP_binary...:104-111 with ``f`` substituted for
    ``f_prior`` in :104-105, and the textbook gradient t - sigmoid(f) (t=(y+1)/2), using diagonal W
    (vector) instead of dense matrices.  Returns (f_hat (N,), w (N,), grad (N,), L, iterations)."""
    y = np.asarray(y_train, dtype=np.float64).reshape(-1)
    n = y.size
    f = np.zeros(n) if f0 is None else np.asarray(f0, dtype=np.float64).reshape(-1).copy()
    t = (y + 1) / 2
    it = 0
    for it in range(1, max_iter + 1):
        p = pi_function(f)
        g = t - p
        w = p * (1 - p)
        sw = np.sqrt(w)
        B = np.eye(n) + sw[:, None] * K * sw[None, :]
        L = np.linalg.cholesky(B)
        b = w * f + g
        c = np.linalg.solve(L, sw * np.dot(K, b))
        a = b - sw * np.linalg.solve(L.T, c)
        f_new = np.dot(K, a)
        err = np.sqrt(np.sum((f_new - f) ** 2))
        f = f_new
        if err <= tolerance:
            break
    p = pi_function(f)
    return f, p * (1 - p), t - p, L, it


def binary_predict_reference(x_star, X_train, L_inv, W, first_deri, kernel_parameter):
    """GP_binary...:136-154 for one or many test points: returns (f_star_mean, var_f_star, label).
    k* uses sigma=kernel_parameter, l=1 (:148-149)."""
    k_star = rbf_kernel(X_train, x_star, kernel_parameter, 1)
    f_star = np.dot(k_star.T, first_deri).reshape(-1)
    v = np.dot(L_inv, np.dot(np.sqrt(W), k_star))
    var = kernel_parameter ** 2 - np.sum(v * v, axis=0)
    label = np.where(pi_function(f_star) >= 0.5, 1, -1)
    return f_star, var, label


# --------------------------------------------------------------------------------------------
# A11  multiclass Laplace  (GP_multi_classification.py:26-197)
# --------------------------------------------------------------------------------------------
def softmax(X):
    """GP_multi...:26-33."""
    e_x = np.exp(X - np.max(X))
    return e_x / e_x.sum(axis=0)


def compute_pi(f, C, n, stride=60):
    """GP_multi...:36-63.  ``stride`` is the literal 60 of :55,:58; pi_vector is class-major with
    that stride, pi_matrix (Cn x n) is filled point-major (:59-61)."""
    pi_vector = np.zeros_like(f)
    pi_matrix = np.zeros((C * n, n))
    for i in range(n):
        triple = np.array([f[j * stride + i] for j in range(C)])
        sm = softmax(triple)
        for j in range(C):
            pi_vector[j * stride + i] = sm[j]
        pi_matrix[i * C:(i + 1) * C, i] = sm
    return pi_vector, pi_matrix


def multi_training_reference(K, y, C, n, tolerance=0.01, max_iter=10000, stride=60):
    """Reference-faithful GP_multi...:129-176 (s=3, step_size=1e-4): returns (pi_vector, f, errors)."""
    step_size = 0.0001
    s = 3
    f = np.zeros((C * n,))
    errors = []
    for _ in range(max_iter):
        pi_vector, pi_matrix = compute_pi(f, C, n, stride)                      # :147
        L = np.linalg.cholesky(s * np.eye(C * n) + K)                           # :148
        L_inv = np.linalg.inv(L)                                                # :149
        W = np.diag(pi_vector) - np.dot(pi_matrix, pi_matrix.T)                 # :150-152
        Kinv = np.dot(L_inv.T, L_inv)
        L2 = np.linalg.cholesky(s * np.eye(C * n) + Kinv + W)                   # :153-155
        L2_inv = np.linalg.inv(L2)                                              # :156
        rhs = np.dot((1 - step_size) * Kinv + W, f) + y + pi_vector            # :157
        f_new = np.dot(L2_inv, rhs)                                             # :158
        err = np.sqrt(np.sum((f_new - f) ** 2))
        errors.append(err)
        f = f_new
        if err <= tolerance:
            break
    return pi_vector, f, errors


def multi_predict_reference(x_star, X_train, C, y, pi_vector, kernel_parameter):
    """GP_multi...:179-197 for many test points: (f_star_mean [m,C], argmax [m])."""
    n = len(X_train)
    k_star = rbf_kernel(X_train, x_star, kernel_parameter, 1)       # (n, m)
    resid = (y - pi_vector).reshape(C, n)                            # class-major
    fm = np.dot(k_star.T, resid.T)                                   # (m, C)
    return fm, np.argmax(fm, axis=1)


def multi_training_newton(Ksub, y, C, n, tolerance=1e-8, max_iter=100):
    """Textbook mode, R&W Alg. 3.3, following the skeleton of the reference's (dead)
    ``model_training`` GP_multi...:92-101,107,113-117 with class-major Pi, R = stacked identities,
    compute_pi's 60 -> n and b = W f + y - pi.  ``Ksub`` is the single n x n block that the
    reference repeats C times on the diagonal (:233-238).
    Returns (pi (C,n), f (C,n), iterations)."""
    y = np.asarray(y, dtype=np.float64).reshape(C, n)
    f = np.zeros((C, n))
    it = 0
    for it in range(1, max_iter + 1):
        fm = f - f.max(axis=0, keepdims=True)
        p = np.exp(fm)
        p /= p.sum(axis=0, keepdims=True)                  # softmax over classes per point
        E = np.empty((C, n, n))
        for c in range(C):
            sd = np.sqrt(p[c])
            Lc = np.linalg.cholesky(np.eye(n) + sd[:, None] * Ksub * sd[None, :])   # :93
            Li = np.linalg.inv(Lc)                                                  # :94
            E[c] = sd[:, None] * np.dot(Li.T, Li) * sd[None, :]                     # :95
        M = np.linalg.cholesky(E.sum(axis=0))                                       # :107
        # W f = D f - Pi Pi^T f ; (Pi Pi^T f)_c,i = p_ci * sum_c' p_c'i f_c'i
        b = p * f - p * (p * f).sum(axis=0, keepdims=True) + y - p                  # :113
        cvec = np.einsum('cij,cj->ci', E, (Ksub @ b.T).T)                          # :114  E K b
        rsum = cvec.sum(axis=0)                                                     # R^T c
        z = np.linalg.solve(M.T, np.linalg.solve(M, rsum))                          # :116
        a = b - cvec + np.einsum('cij,j->ci', E, z)                                 # :116
        f_new = (Ksub @ a.T).T                                                      # :117
        err = np.sqrt(np.sum((f_new - f) ** 2))
        f = f_new
        if err <= tolerance:
            break
    fm = f - f.max(axis=0, keepdims=True)
    p = np.exp(fm)
    p /= p.sum(axis=0, keepdims=True)
    return p, f, it


# --------------------------------------------------------------------------------------------
# Synthetic inputs of SURVEY.md section 8(d): generated by the product's pure-NumPy module so that bench.py's GPU arm
# never has to import the oracle for its inputs; re-exported here for the tests.
# --------------------------------------------------------------------------------------------
from gaussian_process_b200.synthetic import synth_c1, synth_c2, synth_c3, synth_c4, synth_c5  # noqa: E402,F401
