"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (``python -m oracle.gen_golden``); needs /root/reference.  The
reference functions are loaded through ``oracle/ref_loader.py`` (in-memory py2->py3 shim, files
untouched) and executed on the seeded inputs of SURVEY.md Appendix B / section 8(d).  Inputs and
outputs are stored so the GPU box (which has no /root/reference) can check the CUDA path and the
oracle port against the reference's own numbers.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import gp_oracle as O  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402


def _errs(txt):
    """Parse the per-iteration errors the reference prints ("<i>th iteration, error:<repr>")."""
    vals = []
    for line in txt.splitlines():
        if "th iteration, error:" in line:
            tok = line.split("error:")[1].strip()
            tok = tok.replace("np.float64(", "").rstrip(")")
            vals.append(float(tok))
    return vals


def _quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn(*a, **k)
    return out, buf.getvalue()


def ka6_co2_bo() -> None:
    """KA6: the CO2 Bayesian-optimisation driver end to end (CO2_example.py:330-379) on a 60-point Mauna-Loa-shaped
    series: 4 acquisition labels x 10 iterations x 500 candidates, `random` and NumPy RNGs seeded.  Stores the returned
    theta, every printed per-iteration best LML and the book LML."""
    import random
    R = load_reference()
    C2 = R["CO2_example"]

    # The driver predates NumPy 2 / Python 3 in two library calls (np.delete with an EMPTY float index array from
    # overlap(), random.sample on an ndarray).  The reference file stays untouched; its module namespace gets proxies
    # that restore the old library semantics for exactly those two calls.
    class _NpCompat:
        def __getattr__(self, k):
            return getattr(np, k)

        @staticmethod
        def delete(arr, obj, axis=None):
            obj = np.asarray(obj)
            if obj.dtype.kind == "f":
                obj = obj.astype(np.int64)
            return np.delete(arr, obj, axis)

    class _RandomCompat:
        def __getattr__(self, k):
            return getattr(random, k)

        @staticmethod
        def sample(pop, k):
            return random.sample(list(pop), k)

    C2.np, C2.random = _NpCompat(), _RandomCompat()
    X, y, Xs = O.synth_c2(60, 24)
    random.seed(42)
    np.random.seed(42)
    th, txt = _quiet(C2.tune_hyperparameters_BO, X, Xs, y)
    lines = txt.splitlines()
    best = [float(lines[i + 1]) for i, ln in enumerate(lines) if ln.endswith("th iteration!")]
    book = float(lines[-1])
    np.savez(os.path.join(GOLD, "ka6_co2_bo.npz"), N=60, theta=np.asarray(th), best_lml=np.array(best), book_lml=np.float64(book),
             nlines=len(lines))
    print("ka6: theta", th, "best", best[-3:], "book", book)


def main() -> None:
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "ka6":
        return ka6_co2_bo()
    R = load_reference()
    G, T, C2 = R["GP_regression"], R["tune_hyperparms_regression"], R["CO2_example"]
    B, M = R["GP_binary_classification"], R["GP_multi_classification"]
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="gpx_golden_"))
    try:
        # ---- KA1: regression, rbf / lin / per, N=64 and the shipped N=5 -------------------
        out = {}
        for tag, N in (("n64", 64), ("n5", 5)):
            X, y, Xs = O.synth_c1(N, 100)
            out[tag + "_X"], out[tag + "_y"], out[tag + "_Xs"] = X, y, Xs
            for kc, par in (("rbf", 1), ("lin", 0.5), ("per", [2.0, 1.5])):
                if kc == "lin":
                    continue  # rank-1 K + 5e-4 I: posterior sampling chol is not PD in the reference
                np.random.seed(7)
                mu, sd, fp = G.prediction(X, Xs, y, kc, par, 10)
                out["%s_%s_mu" % (tag, kc)] = mu
                out["%s_%s_sd" % (tag, kc)] = sd
                out["%s_%s_fpost" % (tag, kc)] = fp
            out[tag + "_lml"] = np.float64(T.compute_mar_likelihood(X, Xs, y, 1, 1))
            out[tag + "_K_rbf"] = G.RBF_kernel(X, Xs, 1.3, 0.7)
            out[tag + "_K_lin"] = G.lin_kernel(X, Xs, 0.5)
            out[tag + "_K_per"] = G.per_kernel(X, Xs, [2.0, 1.5])
            np.random.seed(11)
            out[tag + "_fprior"] = G.f_prior(Xs, np.zeros((100, 1)), "rbf", 1, 3)
        np.savez(os.path.join(GOLD, "ka1_regression.npz"), **out)

        # ---- KA2: LML + dLML/dl, D=16 (the C5 generator at N=512) ------------------------
        X, y = O.synth_c5(512, 16)
        sigma, l, s = 1.0, 4.0, 5e-4
        K = G.RBF_kernel(X, X, sigma, l)
        L = np.linalg.cholesky(K + s * np.eye(512))
        alpha = np.linalg.solve(L.T, np.linalg.solve(L, y))
        K_y_inv = np.dot(np.linalg.inv(L.T), np.linalg.inv(L))
        _, l_new = T.gradient_ascent(X, X, sigma, l, alpha.reshape(-1, 1), K_y_inv)
        lml = T.compute_mar_likelihood(X, None, y, sigma, l)
        np.savez(os.path.join(GOLD, "ka2_lml_grad.npz"), N=512, D=16, sigma=sigma, l=l, s=s,
                 lml=np.float64(lml), dlml_dl=np.float64((l_new - l) / 0.01), alpha=alpha,
                 logdiag_sum=np.log(np.diagonal(L)).sum(), Kinv_trace=np.trace(K_y_inv))
        # tune_hyperparms_first on a small 1-D problem (full loop incl. prints)
        Xt, yt, Xst = O.synth_c1(8, 100)
        np.random.seed(3)
        (mu, sd, fp, opt), txt = _quiet(T.tune_hyperparms_first, Xt, Xst, yt, 2, 1, np.array([1.7]))
        np.savez(os.path.join(GOLD, "ka2_tune_first.npz"), mu=mu, sd=sd, fpost=fp, lml=np.float64(opt),
                 stdout=np.array(txt))
        # BO posterior over the l axis (tune...:67-101)
        lt = np.array([0.5, 2.0, 3.5]).reshape(-1, 1)
        ltest = np.linspace(0.01, 5, 50).reshape(-1, 1)
        yl = np.array([-3.0, -1.0, -2.5])
        np.random.seed(5)
        mu, sd, fp = T.bayesian_opt(lt, ltest, yl)
        np.savez(os.path.join(GOLD, "ka2_bo.npz"), lt=lt, ltest=ltest, yl=yl, mu=mu, sd=sd, fpost=fp)

        # ---- KA3: CO2 composite -----------------------------------------------------------
        out = {}
        theta = O.CO2_THETA_BOOK
        for N in (468, 2048):
            X, y, Xs = O.synth_c2(N)
            out["lml_%d" % N] = np.float64(C2.compute_mar_likelihood(X, y, theta))
        X, y, Xs = O.synth_c2(468)
        Kc = C2.covariance_function(X, X, theta)
        out["K_probe"] = np.array([Kc[0, 0], Kc[0, 1], Kc[0, 12], Kc[3, 400]])
        out["K_468"] = Kc
        out["Ks_468"] = C2.covariance_function(X, Xs, theta)
        np.random.seed(9)
        mu, sd, fp = C2.make_prediction(X, Xs, y, theta)
        out["mu"], out["sd"], out["fpost"] = mu, sd, fp
        th_tr = C2.init_hyperms(5, 11)
        rs = np.random.RandomState(3)
        th_te = theta * (0.5 + rs.rand(40, 11))
        ylml = rs.randn(5) * 10 - 300
        mu_bo, sd_bo = C2.bayesian_opt(th_tr, th_te, ylml)
        out["bo_theta_train"], out["bo_theta_test"], out["bo_y"] = th_tr, th_te, ylml
        out["bo_mu"], out["bo_sd"] = mu_bo, sd_bo
        np.savez_compressed(os.path.join(GOLD, "ka3_co2.npz"), **out)

        # ---- KA4: binary Laplace ------------------------------------------------------------
        rs = np.random.RandomState(5)
        X = rs.randn(128, 2)
        y = np.where(X[:, 0] * X[:, 1] > 0, 1, -1).reshape(-1, 1)
        K = G.RBF_kernel(X, X, 1, 1)
        fpr = 0.5 * rs.randn(128, 1)
        (W, L_inv, g), txt = _quiet(B.model_training, K, y, fpr, 1)
        errs = _errs(txt)
        Xq = np.array([[0.3, -0.2], [1.0, 1.0], [-0.7, 0.4], [-1.2, -0.3]])
        fbar, var, ok = [], [], []
        for q in Xq:
            xs = q.reshape(-1, 2)
            ks = G.RBF_kernel(X, xs, 1, 1)
            fbar.append(float(np.dot(ks.T, g)))
            v = np.dot(L_inv, np.dot(np.sqrt(W), ks))
            var.append(float(G.RBF_kernel(xs, xs, 1, 1) - np.dot(v.T, v)))
            ok.append(bool(B.prediction(xs, 1, X, L_inv, W, g, 1)))
        np.savez_compressed(os.path.join(GOLD, "ka4_binary.npz"), X=X, y=y, f_prior=fpr, Wdiag=np.diag(W).copy(),
                            L_inv=L_inv, first_deri=g, errors=np.array(errs), Xq=Xq, fbar=np.array(fbar),
                            var=np.array(var), is_plus=np.array(ok))

        # ---- KA5: multiclass Laplace --------------------------------------------------------
        from scipy.linalg import block_diag
        from sklearn.datasets import make_blobs
        Xa, ya = make_blobs(n_samples=100, n_features=2, centers=3, random_state=0)
        Xtr, ytr, Xte, yte = Xa[:60], ya[:60], Xa[60:], ya[60:]
        Ks = G.RBF_kernel(Xtr, Xtr, 1, 1)
        Kb = block_diag(Ks, Ks, Ks)
        yt = np.zeros(180)
        yt[ytr * 60 + np.arange(60)] = 1
        pi, txt = _quiet(M.model_training2, Kb, yt, 3, 60)
        errs = _errs(txt)
        hit = [bool(M.prediction(Xte[i].reshape(-1, 2), yte[i], Xtr, 3, yt, pi, 1)) for i in range(40)]
        rs = np.random.RandomState(1)
        fprobe = rs.randn(180)
        pv, pm = M.compute_pi(fprobe, 3, 60)
        np.savez_compressed(os.path.join(GOLD, "ka5_multi.npz"), Xtr=Xtr, ytr=ytr, Xte=Xte, yte=yte, y_targets=yt,
                            pi_vector=pi, errors=np.array(errs), hits=np.array(hit), fprobe=fprobe,
                            pi_probe=pv, pim_probe=pm)
    finally:
        os.chdir(cwd)
    ka6_co2_bo()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
