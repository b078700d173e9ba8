"""Stand-alone probe of the DMMA GEMM kernel: C -= A A^T (both operands k-major), the Cholesky trailing update shape."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussian_process_b200 import get_engine

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
mode = sys.argv[4] if len(sys.argv) > 4 else "kk"
eng = get_engine()
g = torch.Generator(device="cuda").manual_seed(0)
akm, bkm = mode[0] == "k", mode[1] == "k"
A = torch.randn(M if akm else K, K if akm else M, device="cuda", dtype=torch.float64, generator=g)
B = torch.randn(M if bkm else K, K if bkm else M, device="cuda", dtype=torch.float64, generator=g)
C = torch.zeros(M, M, device="cuda", dtype=torch.float64)
for _ in range(2):
    eng.gemm(A, B, C, akm, bkm, M, M, K, alpha=-1.0, beta=1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    eng.gemm(A, B, C, akm, bkm, M, M, K, alpha=-1.0, beta=1.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("gemm %s M=N=%d K=%d: %.3f ms  %.2f TFLOP/s" % (mode, M, K, ms, 2.0 * M * M * K / ms / 1e9))
