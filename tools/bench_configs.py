"""Per-config timings for BASELINE.json configs C1-C4: GPU seconds through the drop-in API next to
the oracle port of the reference path on the host cores (bounded sample, N^3-scaled where the full size is
infeasible on the CPU).  One JSON line per config.  C5 is bench.py."""
import argparse
import contextlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import gp_oracle as O
from gaussian_process_b200 import get_engine
from gaussian_process_b200._lib import COV_CO2, COV_SE


def gpu_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


def cpu_time(fn, reps=1):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,n1,a10,a11,c2,c3,c4")
    ap.add_argument("--cpu-n", type=int, default=2048)
    a = ap.parse_args()
    eng = get_engine()
    cores = os.cpu_count()
    want = a.configs.split(",")

    if "c1" in want:
        from gaussian_process_b200 import GP_regression as G
        X, y, Xs = O.synth_c1(5, 100)
        np.random.seed(0)
        G.FUSED_SMALL_PATH = False
        tg_tiled, _ = gpu_time(lambda: G.prediction(X, Xs, y, 'rbf', 1, 10), reps=20, warm=3)
        G.FUSED_SMALL_PATH = True
        np.random.seed(0)
        tg, (mu, sd, fp) = gpu_time(lambda: G.prediction(X, Xs, y, 'rbf', 1, 10), reps=50, warm=3)
        np.random.seed(0)
        mu, sd, fp = G.prediction(X, Xs, y, 'rbf', 1, 10)
        np.random.seed(0)
        tc, _ = cpu_time(lambda: O.regression_prediction(X, Xs, y, 'rbf', 1, 10), reps=50)
        np.random.seed(0)
        mu_o, sd_o, fp_o = O.regression_prediction(X, Xs, y, 'rbf', 1, 10)
        print(json.dumps({"config": "C1 GP_regression.prediction N=5 n=100 D=1 (as shipped)", "gpu_s": tg, "gpu_s_tiled_path": tg_tiled,
                          "cpu_s": tc, "cpu_cores": cores,
                          "parity": {"mu": rel(mu, mu_o), "var": rel(sd ** 2, sd_o ** 2), "f_post": rel(fp, fp_o)},
                          "note": "fused path: one H2D copy, ONE kernel launch (csrc/small.cu), one D2H copy; tiled path: ~60 launches "
                                  "and 6 host synchronisations for a 128-padded problem"}))

    if "n1" in want:
        # the gradient-ascent driver at an as-shipped size (N=12 train, n=100 test, 33 iterations)
        from gaussian_process_b200 import GP_regression as G
        from gaussian_process_b200 import tune_hyperparms_regression as T
        X, y, Xs = O.synth_c1(12, 100)
        out = {}
        for name, fused in (("gpu_s_tiled_path", False), ("gpu_s", True)):
            G.FUSED_SMALL_PATH = fused
            np.random.seed(0)
            out[name], res = gpu_time(lambda: quiet(T.tune_hyperparms_first, X, Xs, y, 1, 1.0, np.array([0.6])), reps=5 if fused else 2)
        np.random.seed(0)
        tc, ref = cpu_time(lambda: O.tune_first(X, Xs, y, 1, 1.0, np.array([0.6])), reps=5)
        print(json.dumps({"config": "N1 tune_hyperparms_first N=12 n=100 (gradient ascent on l, %d iterations)" % ref[5], **out,
                          "cpu_s": tc, "cpu_cores": cores,
                          "parity": {"lml": rel(res[3], ref[3]), "mu": rel(res[0], ref[0])},
                          "note": "fused path: the whole ascent loop is ONE kernel launch (gp_small_grad_kernel, all state in shared "
                                  "memory) + the one-launch posterior; tiled path: one fused fit+grad call (~100 launches) per iteration"}))

    if "a10" in want:
        # binary Laplace at the as-shipped size (KA4: N=128, reference-faithful mode, 138 iterations)
        from gaussian_process_b200 import GP_binary_classification as B
        g = np.load(os.path.join(ROOT, "tests", "golden", "ka4_binary.npz"))
        X, y, fpr = g["X"], g["y"], g["f_prior"]
        K = O.rbf_kernel(X, X, 1, 1)
        tg, (W, L_inv, fd) = gpu_time(lambda: quiet(B.model_training, K, y, fpr, 1), reps=3)
        tc, ref = cpu_time(lambda: quiet(O.binary_training_reference, K, y, fpr, 1), reps=3)
        print(json.dumps({"config": "A10 GP_binary_classification.model_training N=128 (as shipped, reference-faithful, 138 iterations)",
                          "gpu_s": tg, "cpu_s": tc, "cpu_cores": cores, "parity": {"L_inv": rel(L_inv, g["L_inv"]), "first_deri": rel(fd, g["first_deri"])}}))

    if "a11" in want:
        # multiclass Laplace at the as-shipped size (KA5: C=3, n=60, reference-faithful mode, 18 iterations)
        from scipy.linalg import block_diag
        from gaussian_process_b200 import GP_multi_classification as M
        g = np.load(os.path.join(ROOT, "tests", "golden", "ka5_multi.npz"))
        Ks = O.rbf_kernel(g["Xtr"], g["Xtr"], 1, 1)
        Kb = block_diag(Ks, Ks, Ks)
        tg, pi = gpu_time(lambda: quiet(M.model_training2, Kb, g["y_targets"], 3, 60), reps=3)
        tc, ref = cpu_time(lambda: quiet(O.multi_training_reference, Kb, g["y_targets"], 3, 60), reps=3)
        print(json.dumps({"config": "A11 GP_multi_classification.model_training2 C=3 n=60 (as shipped, reference-faithful, 18 iterations)",
                          "gpu_s": tg, "cpu_s": tc, "cpu_cores": cores, "parity": {"pi_vector": rel(pi, g["pi_vector"])}}))

    if "c2" in want:
        from gaussian_process_b200 import CO2_example as C2
        th = O.CO2_THETA_BOOK
        X, y, Xs = O.synth_c2(8192)
        tg_lml, lml = gpu_time(lambda: C2.compute_mar_likelihood(X, y, th))
        tg_pred, (mu, sd, fp) = gpu_time(lambda: C2.make_prediction(X, Xs, y, th))
        tg_grad, (lml2, grad) = gpu_time(lambda: C2.compute_mar_likelihood_gradient(X, y, th))
        Xc, yc, Xsc = O.synth_c2(a.cpu_n)
        tc_lml, lml_c = cpu_time(lambda: O.co2_lml(Xc, yc, th))
        tc_pred, (mu_c, sd_c, _) = cpu_time(lambda: O.co2_make_prediction(Xc, Xsc, yc, th))
        lml_g_small = C2.compute_mar_likelihood(Xc, yc, th)
        mu_g, sd_g, _ = C2.make_prediction(Xc, Xsc, yc, th)
        s3 = (8192 / a.cpu_n) ** 3
        print(json.dumps({"config": "C2 CO2 composite kernel N=8192 D=1 (11 theta), 240 test points",
                          "gpu_s": {"compute_mar_likelihood": tg_lml, "make_prediction": tg_pred, "lml_plus_grad_11_theta": tg_grad},
                          "cpu_s_scaled": {"compute_mar_likelihood": tc_lml * s3, "make_prediction": tc_pred * s3},
                          "cpu_sample": "oracle port at N=%d: lml %.2f s, predict %.2f s; scaled by (8192/%d)^3" % (a.cpu_n, tc_lml, tc_pred, a.cpu_n),
                          "cpu_cores": cores, "lml_8192": float(lml), "potrf_alg_tflops": (8192 ** 3 / 3) / tg_lml / 1e12,
                          "parity_at_cpu_n": {"lml": rel(lml_g_small, lml_c), "mu": rel(mu_g, mu_c), "var": rel(sd_g ** 2, sd_c ** 2)}}))

    if "c3" in want:
        from gaussian_process_b200.laplace import BinaryLaplace
        N, D = 16384, 8
        X, y, fpr = O.synth_c3(N, D)
        Xd = eng.to_device(X)
        Kd = eng.cov(COV_SE, Xd, Xd, [1.0, 1.0], same_x=True)
        m = BinaryLaplace(eng, Kd, N)
        t0 = time.perf_counter()
        it = m.fit_newton(y, tolerance=1e-6)
        torch.cuda.synchronize()
        t_newton = time.perf_counter() - t0
        m2 = BinaryLaplace(eng, Kd, N)
        t0 = time.perf_counter()
        it_ref = m2.fit_reference(y, fpr, tolerance=1e-4, max_iter=2000)
        torch.cuda.synchronize()
        t_ref = time.perf_counter() - t0
        n_c = a.cpu_n
        Xc, yc, fc = O.synth_c3(n_c, D)
        Kc = O.rbf_kernel(Xc, Xc, 1, 1)
        tc, (f_o, w_o, g_o, L_o, it_o) = cpu_time(lambda: O.binary_training_newton(Kc, yc, tolerance=1e-6))
        Kdc = eng.cov(COV_SE, eng.to_device(Xc), eng.to_device(Xc), [1.0, 1.0], same_x=True)
        mc = BinaryLaplace(eng, Kdc, n_c)
        mc.fit_newton(yc, tolerance=1e-6)
        print(json.dumps({"config": "C3 binary Laplace N=16384 D=8 SE kernel",
                          "gpu_s": {"newton_total": t_newton, "newton_iterations": it, "per_newton_iteration": t_newton / it,
                                    "reference_faithful_total": t_ref, "reference_faithful_iterations": it_ref},
                          "cpu_s_scaled": {"per_newton_iteration": tc / it_o * (N / n_c) ** 3},
                          "cpu_sample": "oracle textbook Newton at N=%d: %.2f s for %d iterations; per-iteration scaled by (%d/%d)^3" % (n_c, tc, it_o, N, n_c),
                          "cpu_cores": cores, "potrf_alg_tflops_per_iter": (N ** 3 / 3) / (t_newton / it) / 1e12,
                          "parity_at_cpu_n": {"f_hat": rel(eng.to_host(mc.f[:n_c]), f_o)}}))

    if "c4" in want:
        from gaussian_process_b200.laplace import MultiLaplaceNewton
        n, C, D = 8192, 10, 16
        X, labels, y, Xt, tl = O.synth_c4(n, C, D, 2048)
        Xd = eng.to_device(X)
        Kd = eng.cov(COV_SE, Xd, Xd, [1.0, 1.0], same_x=True)
        model = MultiLaplaceNewton(eng, Kd, C, n)
        model.fit(y, tolerance=1e-6, max_iter=1)          # warm-up: allocations, extra handles / streams
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        it = model.fit(y, tolerance=1e-6, max_iter=30)
        torch.cuda.synchronize()
        t_fit = time.perf_counter() - t0
        t0 = time.perf_counter()
        fm = model.predict(Xd, eng.to_device(Xt), y)
        torch.cuda.synchronize()
        t_pred = time.perf_counter() - t0
        acc = float(np.mean(np.argmax(fm, axis=1) == tl))
        n_c = 512
        Xc, lc, yc, _, _ = O.synth_c4(n_c, C, D, 16)
        Kc = O.rbf_kernel(Xc, Xc, 1, 1)
        tc, (p_o, f_o, it_o) = cpu_time(lambda: O.multi_training_newton(Kc, yc, C, n_c, tolerance=1e-6))
        Kdc = eng.cov(COV_SE, eng.to_device(Xc), eng.to_device(Xc), [1.0, 1.0], same_x=True)
        mc = MultiLaplaceNewton(eng, Kdc, C, n_c)
        mc.fit(yc, tolerance=1e-6)
        print(json.dumps({"config": "C4 multiclass softmax Laplace C=10 n=8192 D=16 (per-class factorisations), 2048 test points",
                          "gpu_s": {"fit_total": t_fit, "iterations": it, "per_iteration": t_fit / it, "predict_2048": t_pred},
                          "test_accuracy": acc,
                          "cpu_s_scaled": {"per_iteration": tc / it_o * (n / n_c) ** 3},
                          "cpu_sample": "oracle Alg-3.3 at n=%d C=10: %.2f s for %d iterations; per-iteration scaled by (%d/%d)^3" % (n_c, tc, it_o, n, n_c),
                          "cpu_cores": cores, "alg_tflops_per_iter": (C * n ** 3 + n ** 3 / 3) / (t_fit / it) / 1e12,
                          "parity_at_cpu_n": {"f_hat": rel(eng.to_host(mc.f), f_o)}}))


if __name__ == "__main__":
    main()
