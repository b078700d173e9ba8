"""Drop-in for the reference's ``GP_binary_classification`` module on the gpx B200 engine.

``model_training`` / ``prediction`` keep the reference's signatures and (quirky) arithmetic
(GP_binary_classification.py:86-154): W and the gradient are evaluated at ``f_prior`` on every
iteration, so B = I + W^1/2 K W^1/2 is factored once on the device and each iteration is two
symmetric matvecs plus two triangular solves.  ``mode="newton"`` runs the textbook Newton iteration
(R&W Alg. 3.1) instead -- the variant BASELINE.json's N=16384 configuration describes.
"""
from __future__ import annotations

import numpy as np

from . import GP_regression as _gpr
from .engine import get_engine, padded
from .GP_regression import RBF_kernel, f_prior  # noqa: F401  (re-exported like the reference's imports)
from .laplace import BinaryLaplace

TOLERANCE = 0.0001   # GP_binary...:98
VERBOSE = True       # the reference prints every iteration (:101,116-121)

_last_model = None   # device state of the most recent model_training call (reused by prediction)


def _elementwise(op: int, x):
    eng = get_engine()
    arr = np.asarray(x, dtype=np.float64)
    flat = eng.to_device(arr.reshape(-1))
    out = eng.vec_op(op, flat.numel(), eng.empty(flat.numel()), x=flat)
    return eng.to_host(out).reshape(arr.shape)


def dataset_generator():
    """make_moons(noise=0.3), labels in {-1,+1}, standardised (GP_binary...:13-32); host-side."""
    from sklearn.datasets import make_moons
    from sklearn.preprocessing import StandardScaler
    X, y = make_moons(noise=0.3, random_state=0)
    y[y == 0] = -1
    return StandardScaler().fit_transform(X), y


def pi_function(f):
    """Logistic sigmoid (GP_binary...:48-54)."""
    return _elementwise(9, f)


def label_function(f_star):
    """+1 if sigmoid(f*) >= 0.5 else -1 (GP_binary...:35-45)."""
    return 1 if float(np.asarray(pi_function(f_star)).reshape(-1)[0]) >= 0.5 else -1


def log_likelihood(z):
    """-log(1 + exp(-z)) (GP_binary...:57-63)."""
    return _elementwise(8, z)


def deriv_log_likelihood(y, f):
    """t - sigmoid(y f), t = (y+1)/2 -- the shipped form, including its y=-1 sign (GP_binary...:66-74)."""
    eng = get_engine()
    f = np.asarray(f, dtype=np.float64)
    yb = np.broadcast_to(np.asarray(y, dtype=np.float64), f.shape) if np.ndim(y) == 0 or np.shape(y) != f.shape else np.asarray(y, dtype=np.float64)
    yd, fd = eng.to_device(np.ascontiguousarray(yb).reshape(-1)), eng.to_device(f.reshape(-1))
    g = eng.empty(fd.numel())
    from ._lib import check
    eng._sync_stream()
    check(eng.lib.gpx_logistic_terms(eng.h, 0, fd.numel(), eng._p(yd), eng._p(fd), eng._p(g), None, None), "gpx_logistic_terms")
    return eng.to_host(g).reshape(f.shape)


def sec_deriv_log_likelihood(f):
    """-sigmoid(f)(1 - sigmoid(f)) (GP_binary...:77-83)."""
    eng = get_engine()
    f = np.asarray(f, dtype=np.float64)
    fd = eng.to_device(f.reshape(-1))
    w = eng.empty(fd.numel())
    from ._lib import check
    eng._sync_stream()
    check(eng.lib.gpx_logistic_terms(eng.h, 1, fd.numel(), eng._p(fd), eng._p(fd), None, eng._p(w), None), "gpx_logistic_terms")
    return -eng.to_host(w).reshape(f.shape)


def _device_K(K):
    eng = get_engine()
    K = np.asarray(K, dtype=np.float64)
    n = K.shape[0]
    Kd = eng.zeros(padded(n), padded(n))
    Kd[:n, :n] = eng.to_device(K)
    return Kd, n


def model_training(K, y_train, f_prior, num_funs, mode="reference"):
    """Laplace mode finding -> (W dense (N,N), L_inv dense (N,N), first_deri (N,1)).

    mode="reference": GP_binary...:86-133 as shipped.  mode="newton": textbook Newton (W, gradient at f)."""
    global _last_model
    eng = get_engine()
    Kd, n = _device_K(K)
    model = BinaryLaplace(eng, Kd, n)
    if VERBOSE:
        print("training model!")

    def on_iter(i, err):
        if VERBOSE:
            print((n, num_funs))
            print(repr(i + 1) + "th iteration, error:" + repr(float(err)))

    if mode == "reference":
        iters = model.fit_reference(y_train, f_prior, TOLERANCE, 10000, on_iter)
    elif mode == "newton":
        iters = model.fit_newton(y_train, None, TOLERANCE, 10000, on_iter)
    else:
        raise ValueError("mode must be 'reference' or 'newton'")
    if VERBOSE and model.errors[-1] <= TOLERANCE:
        print("The function has already converged after " + repr(iters) + " iterations!")
        print("The error is " + repr(float(model.errors[-1])))
        print("training end!")
    W = np.zeros((n, n))
    np.fill_diagonal(W, eng.to_host(model.w[:n]))
    L_inv = model.L_inverse_host()
    first_deri = eng.to_host(model.g[:n]).reshape(-1, 1)
    _last_model = dict(model=model, W=W, L_inv=L_inv, first_deri=first_deri, X=None, Xd=None)
    return W, L_inv, first_deri


def _model_for(X_train, L_inv, W, first_deri):
    """Device state matching the (L_inv, W, first_deri) triple: reuse the last training run when the caller
    passes its outputs back (the reference's own calling pattern), otherwise upload."""
    eng = get_engine()
    lm = _last_model
    if lm is not None and lm["L_inv"] is L_inv and lm["W"] is W and lm["first_deri"] is first_deri:
        model = lm["model"]
    else:
        n = np.asarray(W).shape[0]
        npad = padded(n)
        model = BinaryLaplace(eng, eng.zeros(npad, npad), n)     # K itself is not needed to predict
        model.w = eng.zeros(npad)
        model.w[:n] = eng.to_device(np.diag(np.asarray(W, dtype=np.float64)).copy())
        model.sw = eng.vec_op(10, n, eng.zeros(npad), x=model.w)
        model.g = eng.zeros(npad)
        model.g[:n] = eng.to_device(np.asarray(first_deri, dtype=np.float64).reshape(-1))
        model.Linv = eng.zeros(npad, npad)                       # explicit inverse supplied by the caller
        model.Linv[:n, :n] = eng.to_device(np.asarray(L_inv, dtype=np.float64))
    return model


def prediction(x_star, y_star_true, X_train, L_inv, W, first_deri, kernel_parameter):
    """Is the MAP label of x_star equal to y_star_true?  (GP_binary...:136-154; sigma=kernel_parameter, l=1.)"""
    f_mean, _var, labels = predict_many(x_star, X_train, L_inv, W, first_deri, kernel_parameter)
    return bool(labels[0] == y_star_true)


def predict_many(X_star, X_train, L_inv, W, first_deri, kernel_parameter):
    """Batched form of ``prediction``: (f*_mean[m], var_f*[m], labels[m] in {-1,+1})."""
    eng = get_engine()
    model = _model_for(X_train, L_inv, W, first_deri)
    Xd = eng.to_device(np.asarray(X_train, dtype=np.float64))
    Xs = eng.to_device(np.asarray(X_star, dtype=np.float64).reshape(-1, Xd.shape[1]))
    f_mean, var = model.predict(Xd, Xs, float(kernel_parameter), 1.0)
    labels = np.where(pi_function(f_mean) >= 0.5, 1, -1)
    return f_mean, var, labels


if __name__ == "__main__":
    from sklearn.model_selection import train_test_split
    X, y = dataset_generator()
    X_train, X_test, y_train, y_test = train_test_split(X, y, test_size=.4, random_state=42)
    K_train = RBF_kernel(X_train, X_train, 1, l=1)
    lin = np.linspace
    X_sampling = np.concatenate((lin(X[:, 0].min(), X[:, 0].max(), len(X_train)).reshape(-1, 1),
                                 lin(X[:, 1].min(), X[:, 1].max(), len(X_train)).reshape(-1, 1)), axis=1)
    fp = f_prior(X_sampling, np.zeros((len(X_train), 1)), 'rbf', len(X_train), 1)
    W, L_inv, first_deri = model_training(K_train, y_train.reshape(-1, 1), fp, 1)
    _, _, labels = predict_many(X_test, X_train, L_inv, W, first_deri, 1)
    print("classification right rate is: %0.2f" % (np.mean(labels == y_test) * 100))
