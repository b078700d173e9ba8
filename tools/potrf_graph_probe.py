"""gpx_potrf eager vs replayed from a CUDA graph (host enqueue cost of the look-ahead factorisation) at mid N."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gaussian_process_b200 import get_engine
from gaussian_process_b200._lib import COV_SE, check
from gaussian_process_b200 import synthetic as S

eng = get_engine()
for N in [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384]:
    X, y = S.synth_c5(N, 16)
    Xd = eng.to_device(X)
    K = eng.cov(COV_SE, Xd, Xd, [1.0, 4.0], diag_add=5e-4, same_x=True, lower=True)
    A = K.clone()
    dinv = eng.empty(N // 128, 128, 128)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        eng._sync_stream()

        def run():
            A.copy_(K)
            check(eng.lib.gpx_potrf_async(eng.h, eng._p(A), N, N, eng._p(dinv)), "potrf")

        def time_it(fn, reps=10):
            fn(); fn()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                A.copy_(K)
                e0.record(st)
                fn()
                e1.record(st)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            return best

        def eager():
            check(eng.lib.gpx_potrf_async(eng.h, eng._p(A), N, N, eng._p(dinv)), "potrf")

        t_e = time_it(eager)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st, capture_error_mode="relaxed"):
            eager()
        t_g = time_it(g.replay)
    fl = N ** 3 / 3
    print("N=%d potrf eager %.3f ms (%.1f TF) | graph replay %.3f ms (%.1f TF)" % (N, t_e, fl / t_e / 1e9, t_g, fl / t_g / 1e9), flush=True)
