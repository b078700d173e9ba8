"""Times one DMMA GEMM launch shape with explicit leading dimensions (operands are views into large buffers, as in the
factorisation drivers):  python tools/gemm_shape_probe.py M N K lda ldb ldc [mode=kk] [reps=5]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes

import torch

from gaussian_process_b200 import get_engine
from gaussian_process_b200._lib import check


def run(eng, M, N, K, lda, ldb, ldc, mode="kk", reps=5, bufs=None):
    akm, bkm = mode[0] == "k", mode[1] == "k"
    rows_a = M if akm else K
    rows_b = N if bkm else K
    need = max(rows_a * lda, rows_b * ldb)
    if bufs is None or bufs[0].numel() < need or bufs[1].numel() < M * ldc:
        bufs = (torch.randn(need, device="cuda", dtype=torch.float64), torch.zeros(M * ldc, device="cuda", dtype=torch.float64))
    A, C = bufs
    Bp = A.data_ptr() + (8 * 128 * lda if rows_b + 128 <= need // max(lda, ldb) else 0)   # B a few rows below A (same buffer)
    call = lambda: check(eng.lib.gpx_gemm(eng.h, int(akm), int(bkm), M, N, K, -1.0, ctypes.c_void_p(A.data_ptr()), lda,
                                          ctypes.c_void_p(Bp), ldb, 1.0, ctypes.c_void_p(C.data_ptr()), ldc), "gpx_gemm")
    eng._sync_stream()
    call(); call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("gemm %s M=%d N=%d K=%d lda=%d ldb=%d ldc=%d: %.3f ms  %.2f TFLOP/s" % (mode, M, N, K, lda, ldb, ldc, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
    return bufs


if __name__ == "__main__":
    eng = get_engine()
    if len(sys.argv) > 6:
        a = [int(v) for v in sys.argv[1:7]]
        run(eng, *a, mode=sys.argv[7] if len(sys.argv) > 7 else "kk", reps=int(sys.argv[8]) if len(sys.argv) > 8 else 5)
    else:   # the trailing-update shapes of the distributed Cholesky (P = 2 and 8 at N = 65536) against the recursive path's
        bufs = None
        for (M, N, K, lda, ldb, ldc) in [
            (32768, 32768, 32768, 65536, 65536, 65536),   # top-level SYRK of the recursive single-GPU path
            (32768, 32768, 1024, 65536, 65536, 65536),
            (63488, 8192, 1024, 65536, 65536, 32768),     # P=2 deferred chunk, K=1024
            (63488, 8192, 512, 65536, 65536, 32768),
            (63488, 8192, 256, 65536, 65536, 32768),      # ungrouped (round-1) update
            (63488, 2048, 1024, 65536, 65536, 8192),      # P=8 deferred chunk
            (63488, 2048, 256, 65536, 65536, 8192),
            (63488, 8192, 1024, 66048, 66048, 33024),     # same with leading dimensions that are not powers of two
            (63488, 8192, 1024, 1024, 1024, 8192),        # packed operands (ld = K)
            (32768, 256, 256, 65536, 65536, 8192),        # eager column update
        ]:
            bufs = run(eng, M, N, K, lda, ldb, ldc, bufs=bufs)
