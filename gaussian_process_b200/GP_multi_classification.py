"""Drop-in for the reference's ``GP_multi_classification`` module on the gpx B200 engine.

``model_training2`` reproduces the shipped iteration (GP_multi_classification.py:129-176, including the
literal stride 60 of ``compute_pi`` and its point-major Pi); ``model_training_newton`` is the textbook
softmax Laplace (R&W Alg. 3.3) with per-class factorisations that BASELINE.json's C=10 configuration
describes -- it takes the single n x n covariance block instead of the C-fold block-diagonal matrix.
"""
from __future__ import annotations

import numpy as np

from .engine import get_engine, padded
from .GP_regression import RBF_kernel  # noqa: F401
from .laplace import MultiLaplaceNewton, MultiLaplaceReference

VERBOSE = True
REFERENCE_STRIDE = 60   # the literal in GP_multi...:55,58


def softmax(X):
    """exp(X - max X) / column sums (GP_multi...:26-33); for a vector: softmax over its entries."""
    eng = get_engine()
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        C = X.shape[0]
        f = eng.to_device(X)
        pi = eng.empty(C)
        from ._lib import check
        eng._sync_stream()
        check(eng.lib.gpx_softmax_classes(eng.h, C, 1, 1, eng._p(f), eng._p(pi)), "gpx_softmax_classes")
        return eng.to_host(pi)
    # matrix input: the reference subtracts the global max and normalises each column
    C, n = X.shape
    f = eng.to_device(np.ascontiguousarray(X))
    pi = eng.empty(C, n)
    from ._lib import check
    eng._sync_stream()
    check(eng.lib.gpx_softmax_classes(eng.h, C, n, n, eng._p(f), eng._p(pi)), "gpx_softmax_classes")
    return eng.to_host(pi)


def compute_pi(f, C, n):
    """(pi_vector, pi_matrix) as GP_multi...:36-63: pi_vector uses index j*60+i (literal 60), pi_matrix
    (Cn x n) holds point i's class probabilities in rows i*C .. i*C+C-1 of column i."""
    eng = get_engine()
    f = np.asarray(f, dtype=np.float64)
    N = f.shape[0]
    if (C - 1) * REFERENCE_STRIDE + n > N:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % ((C - 1) * REFERENCE_STRIDE + n - 1, N))
    fd = eng.to_device(f)
    pi = eng.zeros(N)
    from ._lib import check
    eng._sync_stream()
    check(eng.lib.gpx_softmax_classes(eng.h, C, n, REFERENCE_STRIDE, eng._p(fd), eng._p(pi)), "gpx_softmax_classes")
    pi_vector = eng.to_host(pi)
    pi_matrix = np.zeros((C * n, n))
    cols = np.arange(n)
    for c in range(C):
        pi_matrix[cols * C + c, cols] = pi_vector[c * REFERENCE_STRIDE + cols]
    return pi_vector, pi_matrix


def model_training2(K, y, C, n):
    """Shipped multiclass iteration -> pi_vector (C*n,).  GP_multi...:129-176 (s = 3, tol = 1e-2)."""
    eng = get_engine()
    K = np.asarray(K, dtype=np.float64)
    N = C * n
    npad = padded(N)
    Kd = eng.zeros(npad, npad)
    Kd[:N, :N] = eng.to_device(K)
    model = MultiLaplaceReference(eng, Kd, C, n, stride=REFERENCE_STRIDE)

    def on_iter(j, err):
        if VERBOSE:
            print((N, N))
            print(repr(j + 1) + "th iteration, error:" + repr(float(err)))

    iters = model.fit(y, 0.01, 10000, on_iter)
    if VERBOSE and model.errors[-1] <= 0.01:
        print("The function has already converged after " + repr(iters) + " iterations!")
        print("The error is " + repr(float(model.errors[-1])))
        print("training end!")
    model_training2.last = model
    return eng.to_host(model.pi[:N])


def model_training_newton(K_sub, y, C, n, tolerance=1e-8, max_iter=100):
    """Textbook softmax Laplace (R&W Alg. 3.3): returns (pi (C*n,) class-major, f (C*n,))."""
    eng = get_engine()
    npad = padded(n)
    Kd = eng.zeros(npad, npad)
    Kd[:n, :n] = eng.to_device(np.asarray(K_sub, dtype=np.float64))
    model = MultiLaplaceNewton(eng, Kd, C, n)
    model.fit(y, tolerance, max_iter)
    model_training_newton.last = model
    return eng.to_host(model.pi).reshape(-1), eng.to_host(model.f).reshape(-1)


def predict_many(X_star, X_train, C, y, pi_vector, kernel_parameter):
    """f*_c = k*^T (y_c - pi_c) for every test point -> (f_mean (m,C), argmax (m,)).  GP_multi...:191-197."""
    eng = get_engine()
    Xd = eng.to_device(np.asarray(X_train, dtype=np.float64))
    Xs = eng.to_device(np.asarray(X_star, dtype=np.float64).reshape(-1, Xd.shape[1]))
    n, m = Xd.shape[0], Xs.shape[0]
    from ._lib import COV_SE, check
    Ks = eng.cov(COV_SE, Xd, Xs, [float(kernel_parameter), 1.0])
    resid = np.asarray(y, dtype=np.float64).reshape(C, n) - np.asarray(pi_vector, dtype=np.float64).reshape(C, n)
    out = np.empty((m, C))
    for c in range(C):
        r = eng.zeros(Ks.shape[0])
        r[:n] = eng.to_device(resid[c])
        mu = eng.empty(m)
        eng._sync_stream()
        check(eng.lib.gpx_predict_moments(eng.h, eng._p(Ks), None, n, m, Ks.stride(0), eng._p(r), None, eng._p(mu), None),
              "gpx_predict_moments")
        out[:, c] = eng.to_host(mu)
    return out, np.argmax(out, axis=1)


def prediction(x_star, y_star_true, X_train, C, y, pi_vector, kernel_parameter):
    """argmax_c f*_c == y_star_true  (GP_multi...:179-197)."""
    _, am = predict_many(x_star, X_train, C, y, pi_vector, kernel_parameter)
    return bool(am[0] == y_star_true)


def dataset_generator():
    """Three Gaussian blobs (GP_multi...:200-211); host-side."""
    from sklearn.datasets import make_blobs
    return make_blobs(n_features=2, centers=3)


if __name__ == "__main__":
    from scipy.linalg import block_diag
    from sklearn.model_selection import train_test_split
    X, y = dataset_generator()
    X_train, X_test, y_train, y_test = train_test_split(X, y, test_size=.4, random_state=42)
    num_train, num_classes = len(X_train), np.size(np.unique(y))
    K_sub = RBF_kernel(X_train, X_train, 1, 1)
    K = block_diag(*([K_sub] * num_classes))
    y_targets = np.zeros((num_classes * num_train,))
    y_targets[y_train * 60 + np.arange(num_train)] = 1
    pi_vector = model_training2(K, y_targets, num_classes, num_train)
    _, am = predict_many(X_test, X_train, num_classes, y_targets, pi_vector, 1)
    print("classification right rate is: %0.2f percent" % (np.mean(am == y_test) * 100))
