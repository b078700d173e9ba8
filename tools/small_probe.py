"""Where the time of the one-launch small posterior goes: raw C-ABI call vs Engine wrapper vs drop-in function, and
the kernel's share (difference between a 1x1 problem and the real one)."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from gaussian_process_b200 import GP_regression as G
from gaussian_process_b200 import get_engine
from gaussian_process_b200._lib import COV_SE


def best(fn, reps=200, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[0] * 1e6, ts[len(ts) // 2] * 1e6


def main():
    eng = get_engine()
    rs = np.random.RandomState(0)
    for N, n, nf in ((1, 1, 1), (5, 100, 0), (5, 100, 10), (64, 100, 10), (128, 128, 10)):
        X = rs.uniform(-5, 5, (N, 1))
        y = np.sin(X).ravel()
        Xs = np.linspace(-5, 5, n).reshape(-1, 1)
        Z = rs.randn(n, max(nf, 1))
        th = (ctypes.c_double * 2)(1.0, 1.0)
        mu, var, fp = np.empty(n), np.empty(n), np.empty((n, max(nf, 1)))
        lml = ctypes.c_double()
        P = lambda a: ctypes.c_void_p(a.ctypes.data)
        args = (eng.h, COV_SE, P(X), N, 1, P(y), P(Xs), n, th, 2, 5e-4, 1e-6, P(Z) if nf else None, nf, P(mu), P(var),
                P(fp) if nf else None, ctypes.byref(lml))
        raw = best(lambda: eng.lib.gpx_gp_small_posterior_host(*args))
        wrap = best(lambda: eng.small_posterior(COV_SE, X, y, Xs, [1.0, 1.0], 5e-4, 1e-6, Z if nf else None))
        line = "N=%3d n=%3d nf=%2d  raw C call %6.1f us (median %6.1f)   Engine.small_posterior %6.1f (%6.1f)" % (
            N, n, nf, raw[0], raw[1], wrap[0], wrap[1])
        if nf:
            drop = best(lambda: G.prediction(X, Xs, y, 'rbf', 1, nf))
            line += "   GP_regression.prediction %6.1f (%6.1f)" % drop
        print(line)
    X, y, Xs = rs.uniform(-5, 5, (5, 1)), rs.randn(5), np.linspace(-5, 5, 100).reshape(-1, 1)
    print("host pieces (us): get_state %.1f, normal(100,10) %.1f" % (
        best(lambda: np.random.get_state())[0], best(lambda: np.random.normal(size=(100, 10)))[0]))


if __name__ == "__main__":
    main()
