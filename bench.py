#!/usr/bin/env python
"""bench.py -- headline benchmark of the exact-GP hot path (BASELINE.json metric).

Workload (config C5): one gradient-ascent iteration body of
tune_hyperparms_regression.tune_hyperparms_first restricted to the train-side operations
(tune...:123,127-129,141,144-145): SE covariance build (+ s I) -> Cholesky -> alpha -> LML -> K^-1 ->
dLML/dtheta, N=65536, D=16, float64, synthetic data (RandomState(2024)).

A "step" is one such pass.  `value` = seconds per step with X, y resident in HBM (CUDA events, max over
ranks); `e2e` = the same through the host-buffer C-ABI call gpx_host_lml (H2D of X, y and D2H of LML +
gradient inside the timed region).  `--impl reference` times the oracle port of the reference's NumPy
path on the host cores on a bounded sample and scales it by N^3.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gp_fit_lml_grad_seconds_n65536_fp64"
S_NOISE = 5e-4
SIGMA, ELL = 1.0, 4.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpx", choices=["gpx", "reference"])
    ap.add_argument("--npoints", dest="n", type=int, default=65536, help="training points (BASELINE config: 65536)")
    ap.add_argument("--dim", dest="d", type=int, default=16)
    ap.add_argument("--cpu-sample-n", type=int, default=4096, help="size of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--block", dest="nb", type=int, default=256, help="block-cyclic block width of the multi-GPU path")
    ap.add_argument("--mg", action="store_true", help="use the block-cyclic multi-GPU driver even at world size 1")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="gpx_clocks_", suffix=".csv")
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.fh, stderr=subprocess.DEVNULL)
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


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(n_sample: int, d: int, n_full: int, repeats: int = 1):
    """Time the oracle port of the reference path (NumPy/OpenBLAS, all host threads) at n_sample and
    scale to n_full by N^3 (the path is 8 N^3-dominated: LU solves, inv, GEMMs)."""
    from oracle import gp_oracle as O
    X, y = O.synth_c5(n_sample, d)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        lml, grad, _ = O.rbf_fit_lml_grad(X, y, SIGMA, ELL, S_NOISE)
        best = min(best, time.perf_counter() - t0)
    scaled = best * (n_full / n_sample) ** 3
    return best, scaled, float(lml), float(grad)


def cpu_best_effort_sample(n_sample: int, d: int, n_full: int):
    """SURVEY 8d "best-effort CPU": the same LML + gradient through chunked kernel + dpotrf/dpotrs/dpotri and the
    O(N^2) trace (oracle.rbf_fit_lml_grad_best_effort), so that the speed-up is not only quoted against the
    reference's LU-solve / dense-inverse / N^3-GEMM path.  The O(N^3) and O(N^2 D) parts are scaled separately."""
    from oracle import gp_oracle as O
    X, y = O.synth_c5(n_sample, d)
    tm = {}
    t0 = time.perf_counter()
    lml, grad, _ = O.rbf_fit_lml_grad_best_effort(X, y, SIGMA, ELL, S_NOISE, timings=tm)
    t = time.perf_counter() - t0
    r = n_full / n_sample
    scaled = tm["n3"] * r ** 3 + tm["n2"] * r ** 2
    return {"value": scaled, "unit": "s", "cores": host_threads(), "kind": "port-best-effort",
            "sample": "chunked kernel + scipy cho_factor/cho_solve/dpotri + O(N^2) trace at N=%d D=%d took %.2f s "
                      "(N^3 part %.2f s, N^2 part %.2f s); parts scaled by (%d/%d)^3 and ^2"
                      % (n_sample, d, t, tm["n3"], tm["n2"], n_full, n_sample),
            "lml_sample": float(lml), "dlml_dl_sample": float(grad)}


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = []
    for i in range(args.warmup + args.steps):
        t, scaled, lml, grad = cpu_reference_sample(args.cpu_sample_n, args.d, args.n)
        if i >= args.warmup:
            times.append(scaled)
        if sum(times) * (args.cpu_sample_n / args.n) ** 3 > 150:   # keep the whole run within a few minutes
            break
    val = float(np.mean(times))
    cores = host_threads()
    sample = ("oracle port of tune_hyperparms_regression.py:123-145 (NumPy %s) at N=%d D=%d measured %.2f s/step, "
              "scaled by (%d/%d)^3" % (np.__version__, args.cpu_sample_n, args.d, val * (args.cpu_sample_n / args.n) ** 3,
                                       args.n, args.cpu_sample_n))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "s", "n_gpus": args.gpus, "steps": len(times),
            "warmup": args.warmup, "ms_per_step": val * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C5 tune_hyperparms_regression LML+grad N=%d D=%d (SE kernel, s=5e-4)" % (args.n, args.d)},
            "cpu_baseline": {"value": val, "unit": "s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpx(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from gaussian_process_b200 import get_engine, padded
    from gaussian_process_b200._lib import COV_SE, check
    from oracle import gp_oracle as O  # synthetic-input generator + cpu_baseline leg only

    eng = get_engine(local_rank)
    lib = eng.lib
    n, D = args.n, args.d
    npad = padded(n)
    X, y = O.synth_c5(n, D)
    theta = np.array([SIGMA, ELL])
    thp = theta.ctypes.data_as(ctypes.c_void_p)

    Xd, yd = eng.to_device(X), eng.to_device(y)
    use_mg = world > 1 or args.mg
    out = eng.empty(16)
    gradp = ctypes.c_void_p(out.data_ptr() + 24)
    if use_mg:
        if world > 1:
            eng.mg_init()
        npad = int(lib.gpx_mg_padded_dim(n, args.nb, world))
        ws = eng.empty(int(lib.gpx_mg_workspace_elems(n, args.nb, world)))
        alpha = eng.empty(npad)
        A = Kinv = None

        def step():
            eng._sync_stream()
            check(lib.gpx_mg_fit_grad(eng.h, COV_SE, eng._p(Xd), n, D, thp, 2, S_NOISE, eng._p(yd), args.nb, eng._p(ws),
                                      eng._p(alpha), eng._p(out), gradp, 1), "gpx_mg_fit_grad")
    else:
        A = eng.empty(npad, npad)
        Kinv = eng.empty(npad, npad)
        dinv = eng.empty(npad // 128, 128, 128)
        alpha = eng.empty(npad)

        def step():
            eng._sync_stream()
            check(lib.gpx_gp_fit_grad(eng.h, COV_SE, eng._p(Xd), n, D, thp, 2, S_NOISE, eng._p(yd), eng._p(A), npad, A.stride(0),
                                      eng._p(dinv), eng._p(Kinv), eng._p(alpha), eng._p(out), gradp), "gpx_gp_fit_grad")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_samples = [eng.fp64_peak(True, 8192)[0]]      # cold (burst) sample
    for _ in range(args.warmup):
        step()
    peak_samples.append(eng.fp64_peak(True, 8192)[0])  # warm sample
    # ---- timed region: K steps, CUDA events on the launching stream, clocks sampled meanwhile
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = eng.launches()
    check(lib.gpx_timing_enable(eng.h, 1), "gpx_timing_enable")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    tbuf = (ctypes.c_double * 16)()
    check(lib.gpx_timing_collect(eng.h, tbuf, 16), "gpx_timing_collect")
    check(lib.gpx_timing_enable(eng.h, 0), "gpx_timing_enable")
    launches = eng.launches() - launches0
    # measured FP64 peaks (MEASURED_PEAKS.json has no FP64 figure): register-resident issue loops, taken right after
    # the timed region while the GPU is at its loaded clocks
    peak_samples.append(eng.fp64_peak(True, 8192)[0])  # right after the timed region
    dmma_peak = max(peak_samples)
    dfma_peak, _ = eng.fp64_peak(False, 8192)
    if world > 1:
        t = torch.tensor([ms_total], device=eng.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    sec_per_step = ms_total / 1e3 / args.steps
    res = eng.to_host(out)
    lml, grad = float(res[0]), res[3:5].copy()

    gemm_ms, gemm_launches, gemm_flops_exec = tbuf[0] / args.steps, tbuf[1] / args.steps, tbuf[2] / args.steps
    phases = {k: tbuf[3 + i] / args.steps for i, k in enumerate(["cov_build", "potrf", "solves_lml", "trtri", "lauum", "gradient"])}
    alg_flops = float(padded(n)) ** 3            # potrf N^3/3 + trtri N^3/3 + lauum N^3/3
    # per-GPU roofline: this rank's share of the algorithmic flops over this rank's DMMA-GEMM kernel time
    achieved = (alg_flops / world) / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    npad = float(padded(n))
    potrf_tflops = (float(npad) ** 3 / 3) / (phases["potrf"] * 1e-3) / 1e12 if phases["potrf"] > 0 else 0.0

    # ---- e2e: host buffers through the C ABI (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        del A, Kinv
        torch.cuda.empty_cache()
        Xh = torch.from_numpy(X).pin_memory()
        yh = torch.from_numpy(y).pin_memory()
        lml_h = ctypes.c_double()
        grad_h = (ctypes.c_double * 2)()

        if use_mg:
            def host_step():   # host X, y -> device every step; LML + gradient read back every step
                xd = Xh.to(eng.device, non_blocking=True)
                yd2 = yh.to(eng.device, non_blocking=True)
                eng._sync_stream()
                check(lib.gpx_mg_fit_grad(eng.h, COV_SE, eng._p(xd), n, D, thp, 2, S_NOISE, eng._p(yd2), args.nb, eng._p(ws),
                                          eng._p(alpha), eng._p(out), gradp, 1), "gpx_mg_fit_grad")
                o = out.cpu()
                lml_h.value = float(o[0])
        else:
            Xn, yn = Xh.numpy(), yh.numpy()

            def host_step():
                check(lib.gpx_host_lml(eng.h, COV_SE, Xn.ctypes.data_as(ctypes.c_void_p), n, D, thp, 2, S_NOISE,
                                       yn.ctypes.data_as(ctypes.c_void_p), ctypes.byref(lml_h), grad_h), "gpx_host_lml")

        host_step()                               # warm-up (allocates the handle's scratch)
        barrier()
        e0.record()
        for _ in range(args.steps):
            host_step()
        e1.record()
        barrier()
        ms_h = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_h], device=eng.device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_h = float(t.item())
        e2e = {"value": ms_h / 1e3 / args.steps, "unit": "s", "h2d_bytes_per_step": int((n * D + n) * 8),
               "d2h_bytes_per_step": int((3 + 2) * 8), "lml": lml_h.value}

    if rank != 0:
        return
    cpu = cpu_best = None
    if not args.no_cpu_baseline and world == 1:      # the CPU legs are timed on rank 0 at N=1 only
        t, scaled, lml_c, grad_c = cpu_reference_sample(args.cpu_sample_n, D, n)
        cpu = {"value": scaled, "unit": "s", "cores": host_threads(), "kind": "port",
               "sample": "oracle port (NumPy/OpenBLAS) one LML+grad iteration at N=%d D=%d took %.2f s; scaled by (%d/%d)^3"
                         % (args.cpu_sample_n, D, t, n, args.cpu_sample_n)}
        cpu_best = cpu_best_effort_sample(args.cpu_sample_n, D, n)
    line = {
        "metric": METRIC, "value": sec_per_step, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec_per_step * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C5 tune_hyperparms_regression LML+grad N=%d D=%d (SE kernel, s=5e-4)" % (n, D),
                   "l2": "inputs (2 x %.1f GB matrices) larger than L2; no flush needed" % (npad * npad * 8 / 1e9),
                   "parallelism": ("single GPU" if not use_mg else
                                   "1-D block-cyclic block columns (nb=%d) over %d GPU(s), NCCL panel broadcast + all-gather" % (args.nb, world))},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s",
                     "frac": achieved / dmma_peak if dmma_peak else None, "traffic": None,
                     "kernel": "dgemm_dmma_kernel (FP64 DMMA.8x8x4)", "alg_flops_per_step": alg_flops, "per": "GPU (rank 0)",
                     "kernel_ms_per_step": gemm_ms, "kernel_launches_per_step": gemm_launches,
                     "kernel_flops_executed_per_step": gemm_flops_exec,
                     "peak_source": "measured in this run: register-resident DMMA.8x8x4 issue loop on 148 SMs, max of the "
                                    "samples taken before warm-up / after warm-up / after the timed region (MEASURED_PEAKS.json "
                                    "has no FP64 figure; 128 flop/clk/SM x 148 SMs x 1.965 GHz = 37.2); DFMA loop = %.1f TF" % dfma_peak,
                     "peak_samples": peak_samples},
        "cpu_baseline": cpu,
        "cpu_best_effort": cpu_best,
        "potrf_tflops": potrf_tflops,
        "phases_ms": phases,
        "lml": lml, "dlml_dsigma": float(grad[0]), "dlml_dl": float(grad[1]),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpx(args)


if __name__ == "__main__":
    main()
