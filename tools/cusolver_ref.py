"""Reference point only (never on the product path): cuSOLVER potrf / cuBLAS DGEMM through torch vs gpx_potrf / gpx_gemm."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussian_process_b200 import get_engine
from gaussian_process_b200._lib import COV_SE
from oracle import gp_oracle as O
eng = get_engine()
torch.backends.cuda.preferred_linalg_library("cusolver")
def t_ms(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for N in [int(a) for a in sys.argv[1:]] or [8192, 16384, 32768]:
    X, y = O.synth_c5(N, 16)
    Xd = eng.to_device(X)
    K = eng.cov(COV_SE, Xd, Xd, [1.0, 4.0], diag_add=5e-4, same_x=True)
    A = K.clone()
    def ours():
        A.copy_(K); eng.potrf(A)
    def copy_only():
        A.copy_(K)
    def cus():
        torch.linalg.cholesky(K)
    tc = t_ms(copy_only); to = t_ms(ours) - tc; tr = t_ms(cus)
    fl = N ** 3 / 3
    print("N=%d potrf: gpx %.2f ms (%.1f TF) | cuSOLVER (torch.linalg.cholesky) %.2f ms (%.1f TF)" % (N, to, fl / to / 1e9, tr, fl / tr / 1e9))
M = 8192
a = torch.randn(M, M, device="cuda", dtype=torch.float64); c = torch.zeros(M, M, device="cuda", dtype=torch.float64)
tg = t_ms(lambda: eng.gemm(a, a, c, True, True, M, M, M, alpha=1.0, beta=0.0)); tb = t_ms(lambda: torch.matmul(a, a.T, out=c))
print("DGEMM 8192^3: gpx %.2f ms (%.1f TF) | cuBLAS %.2f ms (%.1f TF)" % (tg, 2 * M ** 3 / tg / 1e9, tb, 2 * M ** 3 / tb / 1e9))
