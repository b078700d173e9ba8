#!/usr/bin/env python
"""bench.py -- benchmarks of the exact-GP hot path on the BASELINE.json configs.

    python bench.py [--config c5|c4|c3|c2|c1] [--gpus N --steps K --warmup W] [--impl gpx|reference]

Default (= the headline metric): config C5, one gradient-ascent iteration body of tune_hyperparms_regression
(tune...:123,127-129,141,144-145): SE covariance build (+ s I) -> Cholesky -> alpha -> LML -> K^-1 -> dLML/dtheta at
N=65536, D=16, float64, synthetic data.  The other configs are selected with --config:
    c1  GP_regression.prediction as shipped (N=5, n=100, 10 posterior draws)          step = one call
    c2  CO2 composite kernel N=8192: compute_mar_likelihood + make_prediction(240)     step = both calls
    c3  binary Laplace N=16384 D=8, textbook Newton on B = I + W^1/2 K W^1/2           step = one Newton iteration
    c4  multiclass softmax Laplace C=10, n=8192, D=16 (classes sharded over ranks)     step = one Alg-3.3 iteration
    c2p CO2 N=8192: posterior mean / variance of 2048 test points from the factor of the distributed fit, test points
        sharded over the ranks                                             step = one prediction of all points

Every line: `value` = seconds per step with the inputs resident in HBM (CUDA events on the launching stream, max over
ranks); `e2e` = the same work through the public host-buffer API (H2D of the inputs and D2H of the results inside the
timed region); `roofline` for the dominant kernel class; `cpu_baseline` = the oracle port of the reference path timed on
the host cores on a bounded sample (rank 0, N=1 only); `clocks` sampled with nvidia-smi during the timed region.
`--impl reference` times only that CPU arm (all host threads) for the same config / metric.
"""
from __future__ import annotations

import argparse
import contextlib
import ctypes
import io
import json
import os
import subprocess
import sys
import tempfile
import time

# The CPU arm must use every host core.  torchrun exports OMP_NUM_THREADS=1 and OpenBLAS sizes its thread pool when it
# is loaded (raising the limit later through threadpoolctl is reported but not honoured), so the variables are set
# BEFORE numpy is imported, on the rank that times the CPU legs.
if ("reference" in sys.argv or os.environ.get("WORLD_SIZE", "1") == "1") and os.environ.get("RANK", "0") == "0":
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S_NOISE = 5e-4
SIGMA, ELL = 1.0, 4.0
METRICS = {
    "c5": "gp_fit_lml_grad_seconds_n65536_fp64",
    "c4": "multiclass_laplace_seconds_per_iteration_c10_n8192_fp64",
    "c3": "binary_laplace_seconds_per_newton_iteration_n16384_fp64",
    "c2": "co2_lml_plus_prediction_seconds_n8192_fp64",
    "c1": "gp_regression_prediction_seconds_n5_fp64",
    "c2p": "co2_prediction_2048_test_points_seconds_n8192_fp64",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpx", choices=["gpx", "reference"])
    ap.add_argument("--config", default="c5", choices=sorted(METRICS))
    ap.add_argument("--npoints", dest="n", type=int, default=0, help="training points (0 = the BASELINE size of the config)")
    ap.add_argument("--dim", dest="d", type=int, default=0)
    ap.add_argument("--cpu-sample-n", type=int, default=0, help="size of the bounded CPU sample (0 = per-config default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--block", dest="nb", type=int, default=256, help="block-cyclic block width of the multi-GPU path")
    ap.add_argument("--mg", action="store_true", help="use the block-cyclic multi-GPU driver even at world size 1")
    ap.add_argument("--dump-launches", default="", help="write a per-GEMM-launch CSV (phase, shape, ms) of the timed region")
    ap.add_argument("--parity-n", type=int, default=4096, help="size of the in-run multi-GPU parity check (0 = off)")
    return ap.parse_args()


DEFAULT_N = {"c5": (65536, 16), "c4": (8192, 16), "c3": (16384, 8), "c2": (8192, 1), "c1": (5, 1), "c2p": (8192, 1)}
DEFAULT_CPU_N = {"c5": 4096, "c4": 512, "c3": 2048, "c2": 2048, "c1": 5, "c2p": 2048}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_ms: int = 200):
        self.path = tempfile.mktemp(prefix="gpx_clocks_", suffix=".csv")
        self.proc = None
        self.gpu_index = gpu_index
        self.period_ms = period_ms

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, smax, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return out


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
def all_host_threads():
    """(context manager, BLAS thread count actually in use).  The pool size was fixed at import time from the environment
    set at the top of this file (torchrun's OMP_NUM_THREADS=1 silently halved the CPU arm at N > 1 in round 1); the
    count reported is the one the BLAS library itself states."""
    ncpu = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        ctx = threadpool_limits(limits=ncpu)
        n = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        return ctx, int(max(n)) if n else ncpu
    except Exception:
        return contextlib.nullcontext(), int(os.environ.get("OMP_NUM_THREADS", ncpu))


def cpu_sample(config: str, n_full: int, d: int, n_sample: int):
    """One bounded sample of the reference path for `config` on the host: returns (seconds scaled to the full size,
    seconds measured, description).  The oracle uses the reference's own NumPy entry points (LU solves, dense inverses,
    N^3 GEMMs), so its timing is the timing of the reference's path."""
    from oracle import gp_oracle as O
    from gaussian_process_b200 import synthetic as S
    if config == "c5":
        X, y = S.synth_c5(n_sample, d)
        t0 = time.perf_counter()
        O.rbf_fit_lml_grad(X, y, SIGMA, ELL, S_NOISE)
        t = time.perf_counter() - t0
        r = (n_full / n_sample) ** 3
        return t * r, t, ("oracle port of tune_hyperparms_regression.py:123-145 (NumPy %s) at N=%d D=%d: %.2f s; scaled by (%d/%d)^3"
                          % (np.__version__, n_sample, d, t, n_full, n_sample))
    if config == "c2":
        X, y, Xs = S.synth_c2(n_sample, 240)
        t0 = time.perf_counter()
        O.co2_lml(X, y, O.CO2_THETA_BOOK)
        np.random.seed(0)
        O.co2_make_prediction(X, Xs, y, O.CO2_THETA_BOOK)
        t = time.perf_counter() - t0
        r = (n_full / n_sample) ** 3
        return t * r, t, ("oracle port of CO2_example.py:131-149 + :182-214 at N=%d, 240 test points: %.2f s; scaled by (%d/%d)^3"
                          % (n_sample, t, n_full, n_sample))
    if config == "c3":
        X, y, _ = S.synth_c3(n_sample, d)
        K = O.rbf_kernel(X, X, 1, 1)
        t0 = time.perf_counter()
        res = O.binary_training_newton(K, y, tolerance=1e-6)
        t = (time.perf_counter() - t0) / res[4]
        r = (n_full / n_sample) ** 3
        return t * r, t, ("oracle textbook Newton (GP_binary_classification.py:104-111 per iteration) at N=%d D=%d: %.2f s per "
                          "iteration over %d iterations; scaled by (%d/%d)^3" % (n_sample, d, t, res[4], n_full, n_sample))
    if config == "c4":
        X, labels, y, _, _ = S.synth_c4(n_sample, 10, d, 16)
        K = O.rbf_kernel(X, X, 1, 1)
        t0 = time.perf_counter()
        p, f, it = O.multi_training_newton(K, y, 10, n_sample, tolerance=1e-6)
        t = (time.perf_counter() - t0) / it
        r = (n_full / n_sample) ** 3
        return t * r, t, ("oracle Alg-3.3 (GP_multi_classification.py:66-126 per iteration) at C=10 n=%d: %.2f s per iteration "
                          "over %d iterations; scaled by (%d/%d)^3" % (n_sample, t, it, n_full, n_sample))
    if config == "c2p":
        X, y, _ = S.synth_c2(n_sample, 240)
        Xs = X[-1, 0] + (1 + np.arange(2048))[:, None] / 12.0
        K = O.co2_covariance(X, X, O.CO2_THETA_BOOK) + O.S_NOISE * np.eye(n_sample)
        L = np.linalg.cholesky(K)
        alpha = np.linalg.solve(L.T, np.linalg.solve(L, y))
        t0 = time.perf_counter()                       # CO2_example.py:196-208 given the factor
        Ks = O.co2_covariance(X, Xs, O.CO2_THETA_BOOK)
        mu = Ks.T @ alpha
        v = np.linalg.solve(L, Ks)
        var = np.diag(O.co2_covariance(Xs, Xs, O.CO2_THETA_BOOK)) - np.sum(v ** 2, axis=0)
        t = time.perf_counter() - t0
        r = (n_full / n_sample) ** 3                   # np.linalg.solve(L, K_s) LU-factors L: the N^3 term dominates
        return t * r, t, ("oracle port of CO2_example.py:196-208 (K_s, mean, LU solve(L, K_s), variance) for 2048 test points at "
                          "N=%d: %.2f s; scaled by (%d/%d)^3 (the reference's LU solve dominates)" % (n_sample, t, n_full, n_sample))
    X, y, Xs = S.synth_c1(n_full, 100)
    reps = 200
    np.random.seed(0)
    t0 = time.perf_counter()
    for _ in range(reps):
        O.regression_prediction(X, Xs, y, 'rbf', 1, 10)
    t = (time.perf_counter() - t0) / reps
    return t, t, "oracle port of GP_regression.py:109-156 at the full size N=%d n=100 (mean of %d calls): %.3f ms" % (n_full, reps, t * 1e3)


def cpu_best_effort_sample(n_sample: int, d: int, n_full: int):
    """SURVEY 8d "best-effort CPU" for C5: the same LML + gradient through chunked kernel + dpotrf/dpotrs/dpotri and the
    O(N^2) trace, so that the speed-up is not only quoted against the reference's LU-solve / dense-inverse path."""
    from oracle import gp_oracle as O
    from gaussian_process_b200 import synthetic as S
    X, y = S.synth_c5(n_sample, d)
    tm = {}
    t0 = time.perf_counter()
    lml, grad, _ = O.rbf_fit_lml_grad_best_effort(X, y, SIGMA, ELL, S_NOISE, timings=tm)
    t = time.perf_counter() - t0
    r = n_full / n_sample
    return {"value": tm["n3"] * r ** 3 + tm["n2"] * r ** 2, "unit": "s", "cores": os.cpu_count(), "kind": "port-best-effort",
            "sample": "chunked kernel + scipy cho_factor/cho_solve/dpotri + O(N^2) trace at N=%d D=%d took %.2f s "
                      "(N^3 part %.2f s, N^2 part %.2f s); parts scaled by (%d/%d)^3 and ^2"
                      % (n_sample, d, t, tm["n3"], tm["n2"], n_full, n_sample)}


def workload_name(config, n, d):
    return {
        "c5": "C5 tune_hyperparms_regression LML+grad N=%d D=%d (SE kernel, s=5e-4)" % (n, d),
        "c4": "C4 GP_multi_classification softmax Laplace C=10 n=%d D=%d, one Alg-3.3 iteration" % (n, d),
        "c3": "C3 GP_binary_classification Laplace N=%d D=%d SE kernel, one Newton iteration" % (n, d),
        "c2": "C2 CO2_example composite kernel N=%d: compute_mar_likelihood + make_prediction(240 test points)" % n,
        "c1": "C1 GP_regression.prediction as shipped N=%d n=100 D=1, 10 posterior draws" % n,
        "c2p": "C2 CO2_example posterior mean + variance of 2048 test points at N=%d (factor resident, test points sharded)" % n,
    }[config]


def run_reference(args, n, d):
    """CPU arm: the oracle port of the reference path (the reference itself is five Python-2 scripts with no installable
    package; oracle/gp_oracle.py restates its arithmetic with the same NumPy entry points and is pinned to it by
    tests/test_oracle.py).  Rank 0 only; all host threads; each step = one bounded sample, N^3-scaled."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n_sample = args.cpu_sample_n or DEFAULT_CPU_N[args.config]
    ctx, cores = all_host_threads()
    vals, raw, desc = [], [], ""
    with ctx:
        for i in range(args.warmup + args.steps):
            v, t, desc = cpu_sample(args.config, n, d, n_sample)
            if i >= args.warmup:
                vals.append(v)
                raw.append(t)
    val = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRICS[args.config], "value": val, "unit": "s", "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": val * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, n, d)},
            "cpu_baseline": {"value": val, "unit": "s", "cores": cores, "kind": "port",
                             "sample": desc + "; mean measured sample %.3f s over %d steps" % (float(np.mean(raw)), len(raw))},
            "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm: per-config workloads
class Workload:
    """step(): one pass with inputs resident in HBM (enqueue only); host_step(): the same through the public host-buffer
    API; alg_flops: algorithmic flops of one step; finish(): dict of result scalars for the JSON line."""
    scaling = "strong"
    h2d = d2h = 0
    timing_handles = ()

    def finish(self):
        return {}


def make_c5(args, eng, n, D, world, rank, torch):
    from gaussian_process_b200 import padded, synthetic as S
    from gaussian_process_b200._lib import COV_SE, check
    lib = eng.lib
    w = Workload()
    X, y = S.synth_c5(n, D)
    theta = np.array([SIGMA, ELL])
    thp = theta.ctypes.data_as(ctypes.c_void_p)
    Xd, yd = eng.to_device(X), eng.to_device(y)
    use_mg = world > 1 or args.mg
    out = eng.empty(16)
    gradp = ctypes.c_void_p(out.data_ptr() + 24)
    w.parity = None
    if use_mg:
        if world > 1:
            eng.mg_init()
        if args.parity_n and world > 1:
            w.parity = mg_parity_check(args, eng, world, torch, COV_SE, thp, D)
        npad = int(lib.gpx_mg_padded_dim(n, args.nb, world))
        ws = eng.empty(int(lib.gpx_mg_workspace_elems(n, args.nb, world)))
        alpha = eng.empty(npad)

        def step():
            eng._sync_stream()
            check(lib.gpx_mg_fit_grad(eng.h, COV_SE, eng._p(Xd), n, D, thp, 2, S_NOISE, eng._p(yd), args.nb, eng._p(ws),
                                      eng._p(alpha), eng._p(out), gradp, 1), "gpx_mg_fit_grad")
        w.release = lambda: None
    else:
        npad = padded(n)
        bufs = {"A": eng.empty(npad, npad), "Kinv": eng.empty(npad, npad)}
        dinv = eng.empty(npad // 128, 128, 128)
        alpha = eng.empty(npad)

        def step():
            A, Kinv = bufs["A"], bufs["Kinv"]
            eng._sync_stream()
            check(lib.gpx_gp_fit_grad(eng.h, COV_SE, eng._p(Xd), n, D, thp, 2, S_NOISE, eng._p(yd), eng._p(A), npad, A.stride(0),
                                      eng._p(dinv), eng._p(Kinv), eng._p(alpha), eng._p(out), gradp), "gpx_gp_fit_grad")

        def release():
            bufs.clear()
            torch.cuda.empty_cache()
        w.release = release
    w.step = step
    Xh, yh = torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory()
    lml_h, grad_h = ctypes.c_double(), (ctypes.c_double * 2)()
    if use_mg:
        def host_step():   # host X, y -> device every step; LML + gradient read back every step
            xd = Xh.to(eng.device, non_blocking=True)
            yd2 = yh.to(eng.device, non_blocking=True)
            eng._sync_stream()
            check(lib.gpx_mg_fit_grad(eng.h, COV_SE, eng._p(xd), n, D, thp, 2, S_NOISE, eng._p(yd2), args.nb, eng._p(ws),
                                      eng._p(alpha), eng._p(out), gradp, 1), "gpx_mg_fit_grad")
            lml_h.value = float(out.cpu()[0])
    else:
        Xn, yn = Xh.numpy(), yh.numpy()

        def host_step():
            check(lib.gpx_host_lml(eng.h, COV_SE, Xn.ctypes.data_as(ctypes.c_void_p), n, D, thp, 2, S_NOISE,
                                   yn.ctypes.data_as(ctypes.c_void_p), ctypes.byref(lml_h), grad_h), "gpx_host_lml")
    w.host_step = host_step
    w.h2d, w.d2h = int((n * D + n) * 8), int((3 + 2) * 8)
    w.alg_flops = float(padded(n)) ** 3            # potrf N^3/3 + trtri N^3/3 + lauum N^3/3
    w.flops_note = "N^3 = potrf + trtri + lauum, N^3/3 each"
    w.parallelism = ("single GPU" if not use_mg else
                     "1-D block-cyclic block columns (nb=%d, boustrophedon block->rank map) over %d GPU(s): NCCL panel broadcasts, grouped K=1024 trailing updates, K^-1 by two local prefix TRSMs (no all-gather)" % (args.nb, world))
    w.l2 = "inputs (2 x %.1f GB matrices) larger than L2; no flush needed" % (float(padded(n)) ** 2 * 8 / 1e9)
    w.timing_handles = (eng,)
    w.exclusive_kernel_time = not use_mg     # the recursive single-GPU path runs its big GEMMs on one stream

    def finish():
        res = eng.to_host(out)
        r = {"lml": float(res[0]), "dlml_dsigma": float(res[3]), "dlml_dl": float(res[4])}
        if w.parity is not None:
            r["parity"] = w.parity
        return r
    w.finish = finish
    return w


def mg_parity_check(args, eng, world, torch, COV_SE, thp, D):
    """In-run multi-GPU parity (the 2-rank pytest is skipped on a 1-GPU test box): the distributed fit + LML + gradient
    at a reduced N against this rank's own single-GPU gpx_gp_fit_grad on the same inputs; max error over ranks."""
    import torch.distributed as dist
    from gaussian_process_b200 import padded, synthetic as S
    from gaussian_process_b200._lib import check
    lib = eng.lib
    n = args.parity_n
    X, y = S.synth_c5(n, D)
    Xd, yd = eng.to_device(X), eng.to_device(y)
    npm = int(lib.gpx_mg_padded_dim(n, args.nb, world))
    ws = eng.empty(int(lib.gpx_mg_workspace_elems(n, args.nb, world)))
    a_mg, o_mg = eng.empty(npm), eng.empty(16)
    eng._sync_stream()
    check(lib.gpx_mg_fit_grad(eng.h, COV_SE, eng._p(Xd), n, D, thp, 2, S_NOISE, eng._p(yd), args.nb, eng._p(ws), eng._p(a_mg),
                              eng._p(o_mg), ctypes.c_void_p(o_mg.data_ptr() + 24), 1), "gpx_mg_fit_grad(parity)")
    npad = padded(n)
    A, Kinv, dinv = eng.empty(npad, npad), eng.empty(npad, npad), eng.empty(npad // 128, 128, 128)
    a_1, o_1 = eng.empty(npad), eng.empty(16)
    check(lib.gpx_gp_fit_grad(eng.h, COV_SE, eng._p(Xd), n, D, thp, 2, S_NOISE, eng._p(yd), eng._p(A), npad, npad, eng._p(dinv),
                              eng._p(Kinv), eng._p(a_1), eng._p(o_1), ctypes.c_void_p(o_1.data_ptr() + 24)), "gpx_gp_fit_grad(parity)")
    om, o1 = eng.to_host(o_mg), eng.to_host(o_1)
    am, a1 = eng.to_host(a_mg[:n]), eng.to_host(a_1[:n])
    errs = torch.tensor([abs(om[0] - o1[0]) / abs(o1[0]), float(np.max(np.abs(om[3:5] - o1[3:5]) / np.abs(o1[3:5]))),
                         float(np.max(np.abs(am - a1)) / np.max(np.abs(a1)))], dtype=torch.float64, device=eng.device)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    e = [float(v) for v in errs.cpu()]
    ok = e[0] < 1e-8 and e[1] < 1e-7 and e[2] < 1e-7
    del ws, A, Kinv
    torch.cuda.empty_cache()
    if not ok:
        raise AssertionError("multi-GPU parity failed: lml %.2e grad %.2e alpha %.2e" % tuple(e))
    return {"n": n, "world": world, "against": "single-GPU gpx_gp_fit_grad on every rank", "lml_rel": e[0], "grad_rel": e[1],
            "alpha_rel": e[2], "tolerance": {"lml": 1e-8, "grad": 1e-7, "alpha": 1e-7}, "ok": ok}


def make_c2(args, eng, n, D, world, rank, torch):
    from gaussian_process_b200 import CO2_example as C2, padded, synthetic as S
    from gaussian_process_b200._lib import COV_CO2
    from gaussian_process_b200 import GP_regression as G
    w = Workload()
    X, y, Xs = S.synth_c2(n, 240)
    th = C2.HYPERMS_BOOK.astype(np.float64)
    Xd, yd, Xsd = eng.to_device(X), eng.to_device(y), eng.to_device(Xs)
    res = {}

    def step():   # device-resident inputs: fit (LML) + a second fit and the 240-point posterior, as the two reference calls do
        fit = eng.fit(COV_CO2, Xd, yd, th, S_NOISE)
        res["lml"] = fit.lml
        fit2 = eng.fit(COV_CO2, Xd, yd, th, S_NOISE)
        mu, var, V = eng.predict(fit2, Xsd, want_v=True)
        Lp, _ = eng.posterior_sample_factor(COV_CO2, th, Xsd, V, 1e-6)
        res["mu0"] = mu
    w.step = step

    def host_step():
        res["lml_host"] = float(C2.compute_mar_likelihood(X, y, th))
        np.random.seed(0)
        mu, sd, fp = C2.make_prediction(X, Xs, y, th)
        res["mu_host"] = float(mu[0])
    w.host_step = host_step
    w.release = lambda: None
    m = Xs.shape[0]
    w.h2d, w.d2h = int(2 * (n * D + n) * 8 + m * D * 8), int((1 + 3 * m) * 8)
    npad = float(padded(n))
    w.alg_flops = 2 * npad ** 3 / 3 + 2 * npad * npad * padded(m)
    w.flops_note = "two Cholesky factorisations (N^3/3 each: the reference factors in compute_mar_likelihood and again in make_prediction) + the 240-column triangular solve"
    w.parallelism = "single GPU" if world == 1 else "%d independent replicas (the N=8192 factorisation does not shard profitably)" % world
    w.l2 = "K (%.2f GB) larger than L2; no flush needed" % (npad * npad * 8 / 1e9)
    w.timing_handles = (eng,)
    w.exclusive_kernel_time = False
    w.scaling = "weak" if world > 1 else "strong"
    w.finish = lambda: {"lml": float(res.get("lml", float("nan"))), "lml_e2e": res.get("lml_host")}
    return w


def make_c2p(args, eng, n, D, world, rank, torch):
    from gaussian_process_b200 import CO2_example as C2, padded, parallel as P, synthetic as S
    from gaussian_process_b200._lib import COV_CO2
    w = Workload()
    X, y, _ = S.synth_c2(n, 240)
    m = 2048
    Xs = X[-1, 0] + (1 + np.arange(m))[:, None] / 12.0
    th = C2.HYPERMS_BOOK.astype(np.float64)
    if world > 1:
        eng.mg_init()
    fit = eng.mg_fit(COV_CO2, X, y, th, S_NOISE, nb=args.nb)      # distributed fit: the factor is replicated on every rank
    lo, hi = P.shard_range(m, rank, world)
    Xs_loc = eng.to_device(Xs[lo:hi])
    res = {}

    def step():   # this rank's shard of test points, from the replicated factor (no refit, no communication)
        res["mu"], res["var"], _ = eng.predict(fit, Xs_loc, m_total=m, row0=lo)
    w.step = step

    def host_step():   # host test points in, all-gathered (mu, var) out on the host
        res["mu_h"], res["var_h"] = eng.mg_predict(fit, Xs)
    w.host_step = host_step
    w.release = lambda: None
    w.h2d, w.d2h = int(m * D * 8), int(2 * m * 8)
    npad = float(padded(n))
    w.alg_flops = npad * npad * padded(m) + 2 * npad * padded(m)
    w.flops_note = "triangular solve L^-1 K_s for 2048 columns (N^2 m) + the two column reductions"
    w.parallelism = "single GPU" if world == 1 else "%d ranks, %d test points each, factor replicated by the distributed fit" % (world, hi - lo)
    w.l2 = "factor (%.2f GB) larger than L2; no flush needed" % (npad * npad * 8 / 1e9)
    w.timing_handles = (eng,)
    w.exclusive_kernel_time = False
    w.finish = lambda: {"lml": fit.lml, "mu0": float(res["mu"][0].item()) if hi > lo else None}
    return w


def make_c3(args, eng, n, D, world, rank, torch):
    from gaussian_process_b200 import padded, synthetic as S
    from gaussian_process_b200 import GP_binary_classification as B
    from gaussian_process_b200._lib import COV_SE
    from gaussian_process_b200.laplace import BinaryLaplace
    w = Workload()
    X, y, fpr = S.synth_c3(n, D)
    distributed = world > 1 or args.mg
    if distributed:
        # B = I + W^1/2 K W^1/2 factored by the block-cyclic multi-GPU Cholesky (gpx_mg_laplace_binary_step)
        from gaussian_process_b200.distributed import BinaryLaplaceDistributed
        model = BinaryLaplaceDistributed(eng, X, 1.0, 1.0, nb=args.nb)
        npad_m = model.npad
        yd = eng.zeros(npad_m)
        yd[:n] = eng.to_device(y.reshape(-1))
        st = {"f": eng.zeros(npad_m), "fn": eng.zeros(npad_m), "it": 0}
        err_t = model.err

        def step():
            eng._sync_stream()
            model.step(yd, st["f"], st["fn"])
            st["f"], st["fn"] = st["fn"], st["f"]
            st["it"] += 1
        w.host_step = None
    else:
        Xd = eng.to_device(X)
        Kd = eng.cov(COV_SE, Xd, Xd, [1.0, 1.0], same_x=True)
        model = BinaryLaplace(eng, Kd, n)
        yd = model._pad_vec(y)
        st = {"f": eng.zeros(model.npad), "fn": eng.zeros(model.npad), "it": 0}
        model._workspace()
        err_t = model._err

        def step():   # one Newton iteration from the current iterate (the first steps are the real Newton trajectory from f = 0)
            eng._sync_stream()
            model.newton_step(yd, st["f"], st["fn"])
            st["f"], st["fn"] = st["fn"], st["f"]
            st["it"] += 1
        Kh = {}

        def host_step():   # public API: host K, y in; (W, L_inv, gradient) out -- whole fit, reported per Newton iteration
            if "K" not in Kh:
                Kh["K"] = eng.to_host(Kd[:n, :n])
            t0 = time.perf_counter()
            Wm, Linv, g = quiet(B.model_training, Kh["K"], y, fpr, 1, mode="newton")
            st["e2e_iters"] = len(B._last_model["model"].errors)
            st["e2e_total"] = time.perf_counter() - t0
        w.host_step = host_step
    w.step = step
    w.e2e_per_iteration = True
    w.release = lambda: None
    npad = float(padded(n))
    w.h2d, w.d2h = int((n * n + 2 * n) * 8), int((n * n + 2 * n) * 8)
    w.alg_flops = npad ** 3 / 3
    w.flops_note = "one Cholesky of B = I + W^1/2 K W^1/2 (N^3/3) per Newton iteration"
    w.parallelism = ("single GPU" if not distributed else
                     "B built and factored block-cyclically (nb=%d) over %d GPU(s), NCCL panel broadcast; mat-vecs and solves "
                     "replicated" % (args.nb, world))
    w.l2 = "K and B (%.1f GB each) larger than L2; no flush needed" % (npad * npad * 8 / 1e9)
    w.timing_handles = (eng,)
    w.exclusive_kernel_time = False

    def finish():
        return {"newton_iterations_run": st["it"], "last_error": float(err_t[0].item()),
                "e2e_note": ("whole model_training(mode='newton') call incl. 2.1 GB H2D of K and D2H of inv(L): %.3f s for %d iterations"
                             % (st.get("e2e_total", float("nan")), st.get("e2e_iters", 0))) if not distributed else None}
    w.finish = finish
    w.state = st
    return w


def make_c4(args, eng, n, D, world, rank, torch):
    from gaussian_process_b200 import padded, parallel as P, synthetic as S
    from gaussian_process_b200 import GP_multi_classification as M
    from gaussian_process_b200._lib import COV_SE
    from gaussian_process_b200.laplace import MultiLaplaceNewton
    C = 10
    w = Workload()
    X, labels, y, Xt, tl = S.synth_c4(n, C, D, 2048)
    Xd = eng.to_device(X)
    Kd = eng.cov(COV_SE, Xd, Xd, [1.0, 1.0], same_x=True)
    if world > 1:
        eng.mg_init()
    model = MultiLaplaceNewton(eng, Kd, C, n, classes=P.shard_classes(C, rank, world))
    model._prepare()
    yd = eng.to_device(y.reshape(C, n))
    st = {"f": eng.zeros(C, n), "fn": eng.zeros(C, n), "pi": eng.zeros(C, n), "it": 0}

    def step():
        eng._sync_stream()
        model.step(yd, st["f"], st["fn"], st["pi"])
        st["f"], st["fn"] = st["fn"], st["f"]
        st["it"] += 1
    w.step = step
    Kh = {}

    def host_step():   # public API (single process): host K_sub, y in; pi, f out -- whole fit, reported per iteration
        if "K" not in Kh:
            Kh["K"] = eng.to_host(Kd[:n, :n])
        t0 = time.perf_counter()
        pi, f = M.model_training_newton(Kh["K"], y, C, n, tolerance=1e-6, max_iter=30)
        st["e2e_iters"] = len(M.model_training_newton.last.errors)
        st["e2e_total"] = time.perf_counter() - t0
    w.host_step = host_step if world == 1 else None
    w.e2e_per_iteration = True
    w.release = lambda: None
    npad = float(padded(n))
    w.h2d, w.d2h = int((n * n + C * n) * 8), int(2 * C * n * 8)
    w.alg_flops = C * npad ** 3 + npad ** 3 / 3
    w.flops_note = "per class potrf + trtri + lauum of B_c (N^3) x 10 classes + chol(sum_c E_c) (N^3/3)"
    w.parallelism = ("single GPU, %d streams" % model.nstreams if world == 1 else
                     "classes sharded over %d ranks (class c on rank c mod P; sum_c E_c, R^T c, f all-reduced with NCCL)" % world)
    w.l2 = "every B_c (%.2f GB) larger than L2; no flush needed" % (npad * npad * 8 / 1e9)
    w.timing_handles = tuple([eng] + [ln["eng"] for ln in model._lanes])
    w.exclusive_kernel_time = False

    def finish():
        return {"iterations_run": st["it"], "last_error": float(model._err[0].item()), "classes_on_rank0": model.classes,
                "e2e_note": ("whole model_training_newton call incl. H2D of K_sub: %.3f s for %d iterations"
                             % (st.get("e2e_total", float("nan")), st.get("e2e_iters", 0))) if world == 1 else None}
    w.finish = finish
    w.state = st
    return w


def make_c1(args, eng, n, D, world, rank, torch):
    from gaussian_process_b200 import GP_regression as G, synthetic as S
    from gaussian_process_b200._lib import COV_SE
    w = Workload()
    X, y, Xs = S.synth_c1(n, 100)
    Z = np.random.RandomState(0).normal(size=(100, 10))
    res = {}

    def step():   # raw C-ABI call of the fused one-launch posterior (host pointers by design: 1.7 KB in, 9.6 KB out)
        res["out"] = eng.small_posterior(COV_SE, X, y, Xs, [1.0, 1.0], S_NOISE, 1e-6, Z)
    w.step = step

    def host_step():
        np.random.seed(0)
        res["mu"], res["sd"], res["fp"] = G.prediction(X, Xs, y, 'rbf', 1, 10)
    w.host_step = host_step
    w.release = lambda: None
    w.h2d, w.d2h = int((n * D + n + 100 * D + 100 * 10) * 8), int((100 * 2 + 100 * 10 + 1) * 8)
    w.alg_flops = 0.0
    w.bytes = float(w.h2d + w.d2h)
    w.flops_note = "latency-bound: one CTA, serial chain of 100 pivots; algorithmic bytes = inputs + outputs"
    w.parallelism = "single GPU, one thread block" if world == 1 else "%d independent replicas" % world
    w.l2 = "problem fits in shared memory; nothing to flush"
    w.timing_handles = ()
    w.exclusive_kernel_time = True
    w.scaling = "weak" if world > 1 else "strong"
    w.finish = lambda: {"lml": float(res["out"][3])}
    return w


MAKERS = {"c5": make_c5, "c4": make_c4, "c3": make_c3, "c2": make_c2, "c1": make_c1, "c2p": make_c2p}


def run_gpx(args, n, D):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from gaussian_process_b200 import get_engine
    from gaussian_process_b200._lib import check

    eng = get_engine(local_rank)
    lib = eng.lib
    cfg = args.config
    w = MAKERS[cfg](args, eng, n, D, world, rank, torch)
    steps, warmup = args.steps, args.warmup
    if cfg == "c1":           # a 0.1 ms step: repeat it so that the timed region is long enough for the clock sampler
        inner = 500
    else:
        inner = 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_samples = [eng.fp64_peak(True, 8192)[0]]      # cold (burst) sample
    for _ in range(warmup):
        for _ in range(inner):
            w.step()
    torch.cuda.synchronize()
    peak_samples.append(eng.fp64_peak(True, 8192)[0])  # warm sample
    # ---- timed region: K steps, CUDA events on the launching stream, clocks sampled meanwhile
    sampler = ClockSampler(local_rank, 100 if cfg != "c5" else 200)
    barrier()
    sampler.start()
    launches0 = sum(e.launches() for e in (w.timing_handles or (eng,)))
    for e in w.timing_handles:
        check(lib.gpx_timing_enable(e.h, 1), "gpx_timing_enable")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        for _ in range(inner):
            w.step()
    e1.record()
    barrier()
    if cfg != "c5":           # short timed regions: keep the sampler alive long enough for a few samples under load
        if world > 1:         # a fixed count on every rank (the steps contain collectives)
            for _ in range(8):
                w.step()
        else:
            t_end = time.time() + 1.0
            while time.time() < t_end:
                w.step()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    gemm_ms = gemm_launches = gemm_flops_exec = 0.0
    phases = None
    for e in w.timing_handles:
        tbuf = (ctypes.c_double * 16)()
        check(lib.gpx_timing_collect(e.h, tbuf, 16), "gpx_timing_collect")
        if args.dump_launches and e is eng:
            check(lib.gpx_timing_dump(e.h, (args.dump_launches + ".rank%d.csv" % rank).encode()), "gpx_timing_dump")
        check(lib.gpx_timing_enable(e.h, 0), "gpx_timing_enable")
        gemm_ms += tbuf[0] / steps
        gemm_launches += tbuf[1] / steps
        gemm_flops_exec += tbuf[2] / steps
        if phases is None:
            phases = {k: tbuf[3 + i] / steps for i, k in enumerate(["cov_build", "potrf", "solves_lml", "trtri", "lauum", "gradient"])}
    launches = sum(e.launches() for e in (w.timing_handles or (eng,))) - launches0
    peak_samples.append(eng.fp64_peak(True, 8192)[0])  # right after the timed region
    dmma_peak = max(peak_samples)
    dfma_peak, _ = eng.fp64_peak(False, 8192)
    if world > 1:
        t = torch.tensor([ms_total], device=eng.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    sec_per_step = ms_total / 1e3 / steps / inner
    results = w.finish()

    # ---- e2e: host buffers through the public API (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e and w.host_step is not None:
        w.release()
        w.host_step()                               # warm-up (allocations)
        barrier()
        e0.record()
        t0 = time.perf_counter()
        nrep = steps if not getattr(w, "e2e_per_iteration", False) else 1
        for _ in range(nrep * inner):
            w.host_step()
        e1.record()
        barrier()
        ms_h = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3 if cfg == "c1" else 0.0)
        if world > 1:
            t = torch.tensor([ms_h], device=eng.device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_h = float(t.item())
        if getattr(w, "e2e_per_iteration", False):
            iters = max(1, w.state.get("e2e_iters", 1))
            e2e = {"value": ms_h / 1e3 / iters, "unit": "s", "h2d_bytes_per_step": int(w.h2d / iters),
                   "d2h_bytes_per_step": int(w.d2h / iters), "iterations": iters, "total_s": ms_h / 1e3}
        else:
            e2e = {"value": ms_h / 1e3 / nrep / inner, "unit": "s", "h2d_bytes_per_step": w.h2d, "d2h_bytes_per_step": w.d2h}
        results.update(w.finish())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline
    if cfg == "c1":
        from_peaks = measured_hbm_peak()
        ach = w.bytes / sec_per_step / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": from_peaks[0], "unit": "GB/s", "frac": ach / from_peaks[0],
                    "traffic": None, "kernel": "gp_small_kernel<SE> (one 1024-thread CTA)", "peak_source": from_peaks[1],
                    "note": w.flops_note}
    else:
        # algorithmic flops over the WALL-CLOCK step time of this rank's share: conservative (it charges the non-GEMM phases
        # to the kernel) and exclusive (the per-launch event pairs overlap when the look-ahead factorisation runs GEMMs on
        # three streams, so their sum can exceed the wall clock); the event-pair sum is reported next to it
        ach = (w.alg_flops / world) / sec_per_step / 1e12
        per = "GPU (rank 0): algorithmic flops / wall-clock step time"
        ach_sum = (w.alg_flops / world) / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        roofline = {"bound": "tensor", "achieved": ach, "peak": dmma_peak, "unit": "TFLOP/s",
                    "frac": ach / dmma_peak if dmma_peak else None, "traffic": None,
                    "kernel": "dgemm_dmma_tma_kernel / dgemm_dmma_kernel (FP64 DMMA.8x8x4)", "alg_flops_per_step": w.alg_flops,
                    "alg_flops_note": w.flops_note, "per": per, "achieved_by_kernel_event_sum": ach_sum,
                    "kernel_ms_per_step_summed": gemm_ms,
                    "kernel_launches_per_step": gemm_launches, "kernel_flops_executed_per_step": gemm_flops_exec,
                    "peak_source": "measured in this run: register-resident DMMA.8x8x4 issue loop on 148 SMs, max of the samples "
                                   "taken before warm-up / after warm-up / after the timed region (MEASURED_PEAKS.json has no "
                                   "FP64 figure; 128 flop/clk/SM x 148 SMs x 1.965 GHz = 37.2); DFMA loop = %.1f TF" % dfma_peak,
                    "peak_samples": peak_samples}
    cpu = cpu_best = None
    if not args.no_cpu_baseline and world == 1:      # the CPU legs are timed on rank 0 at N=1 only
        n_sample = args.cpu_sample_n or DEFAULT_CPU_N[cfg]
        ctx, cores = all_host_threads()
        with ctx:
            v, t, desc = cpu_sample(cfg, n, D, n_sample)
            cpu = {"value": v, "unit": "s", "cores": cores, "kind": "port", "sample": desc}
            if cfg == "c5":
                cpu_best = cpu_best_effort_sample(n_sample, D, n)
    line = {
        "metric": METRICS[cfg], "value": sec_per_step, "unit": "s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": sec_per_step * 1e3, "higher_is_better": False, "scaling": w.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg, n, D), "l2": w.l2, "parallelism": w.parallelism},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    if cpu_best:
        line["cpu_best_effort"] = cpu_best
    if phases and cfg == "c5":
        line["phases_ms"] = phases
        npad = float(int(lib.gpx_padded_dim(n)))
        line["potrf_tflops"] = (npad ** 3 / 3) / (phases["potrf"] * 1e-3) / 1e12 if phases["potrf"] > 0 else 0.0
    if inner > 1:
        line["config"]["calls_per_step_loop"] = inner
    line.update(results)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measured_hbm_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        return 6457.0, "B200_PROFILING.md fallback (no MEASURED_PEAKS.json on this box)"


def main():
    args = parse()
    n, d = DEFAULT_N[args.config]
    n, d = (args.n or n), (args.d or d)
    if args.impl == "reference":
        run_reference(args, n, d)
    else:
        run_gpx(args, n, d)


if __name__ == "__main__":
    main()
