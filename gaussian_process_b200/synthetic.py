"""Seeded synthetic inputs of the BASELINE.json configs C1-C5.

Shared by the drop-in modules' ``__main__`` demos, ``bench.py`` and the tests; pure NumPy, no CUDA.  The reference
reads its data from matplotlib clicks / ``fetch_mldata`` (dead upstream API); these generators replace them with
data of the same shape.
"""
from __future__ import annotations

import numpy as np

def synth_c1(N=5, n=100):
    rs = np.random.RandomState(1234)
    X = rs.uniform(-5, 5, (N, 1))
    y = np.sin(0.9 * X).ravel() + np.sqrt(5e-4) * rs.randn(N)
    return X, y, np.linspace(-5, 5, n)[:, None]


def synth_c2(N=8192, n_test=240):
    t = 1958 + np.arange(N) / 12.0
    y = (315 + 1.3 * (t - 1958) + 0.012 * (t - 1958) ** 2 + 3 * np.sin(2 * np.pi * t)
         + 0.8 * np.sin(4 * np.pi * t) + 0.3 * np.random.RandomState(0).randn(N))
    y = y - y.mean()
    X = t[:, None]
    Xs = np.arange(X.max() // 1 + 1, X.max() // 1 + 21, 1. / 12)[:n_test, None]
    return X, y, Xs


def synth_c3(N=16384, D=8):
    rs = np.random.RandomState(5)
    X = rs.randn(N, D)
    y = np.where(X[:, 0] * X[:, 1] > 0, 1.0, -1.0).reshape(-1, 1)
    f_prior_ = 0.5 * rs.randn(N, 1)
    return X, y, f_prior_


def synth_c4(n=8192, C=10, D=16, n_test=2048):
    centres = 3 * np.random.RandomState(0).randn(C, D)
    labels = np.arange(n) % C
    X = centres[labels] + np.random.RandomState(1).randn(n, D)
    y = np.zeros(C * n)
    y[labels * n + np.arange(n)] = 1
    tl = np.arange(n_test) % C
    Xt = centres[tl] + np.random.RandomState(2).randn(n_test, D)
    return X, labels, y, Xt, tl


def synth_c5(N=65536, D=16):
    rs = np.random.RandomState(2024)
    X = rs.randn(N, D)
    y = np.sin(X.sum(1)) + 0.05 * rs.randn(N)
    return X, y
