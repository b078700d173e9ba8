import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussian_process_b200 import get_engine
from gaussian_process_b200._lib import COV_SE
from oracle import gp_oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
eng = get_engine()
X, y = O.synth_c5(N, 16)
Xd = eng.to_device(X)
A = eng.cov(COV_SE, Xd, Xd, [1.0, 4.0], diag_add=5e-4, same_x=True)
dinv = eng.potrf(A)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
x = eng.to_device(y)
print("N=%d block_inverses %.3f ms" % (N, t(lambda: eng.block_inverses(A, dinv))))
D, bs = eng.block_inverses(A, dinv)
print("bs=%d trsv (leaf chain) %.3f ms | trsv_big %.3f ms" % (bs, t(lambda: eng.trsv(A, dinv, x)), t(lambda: eng.trsv_big(A, D, bs, x))))
B = eng.zeros(N, 256)
print("trsm 256 cols (leaf chain) %.3f ms | trsm_big %.3f ms" % (t(lambda: eng.trsm(A, dinv, B)), t(lambda: eng.trsm_big(A, D, bs, B))))
