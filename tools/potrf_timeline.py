"""Per-launch timeline of one gpx_potrf at mid N (CUDA-event pairs around every GEMM launch, gpx_timing_dump)."""
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes

import torch

from gaussian_process_b200 import get_engine, synthetic as S
from gaussian_process_b200._lib import COV_SE, check

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
eng = get_engine()
X, y = S.synth_c5(N, 16)
Xd = eng.to_device(X)
K = eng.cov(COV_SE, Xd, Xd, [1.0, 4.0], diag_add=5e-4, same_x=True, lower=True)
A = K.clone()
dinv = eng.empty(N // 128, 128, 128)
for _ in range(2):
    A.copy_(K)
    check(eng.lib.gpx_potrf_async(eng.h, eng._p(A), N, N, eng._p(dinv)), "potrf")
torch.cuda.synchronize()
A.copy_(K)
check(eng.lib.gpx_timing_enable(eng.h, 1), "t")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
check(eng.lib.gpx_potrf_async(eng.h, eng._p(A), N, N, eng._p(dinv)), "potrf")
e1.record()
torch.cuda.synchronize()
buf = (ctypes.c_double * 16)()
check(eng.lib.gpx_timing_collect(eng.h, buf, 16), "c")
path = "/tmp/potrf_tl.csv"
check(eng.lib.gpx_timing_dump(eng.h, path.encode()), "d")
rows = list(csv.DictReader(open(path)))
print("N=%d potrf %.3f ms with timing events; %d GEMM launches, GEMM event sum %.2f ms, leaf sum %.2f ms (%d)" % (N, e0.elapsed_time(e1), len(rows), buf[0], buf[11], buf[12]))
t0 = float(rows[0]["start_ms"])
big = [r for r in rows if int(r["K"]) > 128]
print("grouped launches (K > 128): start  dur   M N K  TF")
for r in big[:40]:
    ms = float(r["ms"])
    print("  %7.3f %6.3f  %5s %5s %4s  %.1f" % (float(r["start_ms"]) - t0, ms, r["M"], r["N"], r["K"], float(r["flops"]) / ms / 1e9))
small = [r for r in rows if int(r["K"]) == 128 and int(r["M"]) == 128 and int(r["N"]) == 128]
ts = [float(r["start_ms"]) - t0 for r in small]
print("diag-block updates (chain): n=%d, first 24 start times:" % len(ts), " ".join("%.3f" % t for t in ts[:24]))
if len(ts) > 2:
    import numpy as np
    d = np.diff(ts)
    print("chain step (start-to-start of consecutive diag updates): median %.1f us, mean %.1f us, max %.1f us" % (np.median(d) * 1e3, d.mean() * 1e3, d.max() * 1e3))
