import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussian_process_b200 import get_engine, synthetic as S
from gaussian_process_b200._lib import COV_SE
eng = get_engine()
X, y = S.synth_c5(128, 16)
K = eng.cov(COV_SE, eng.to_device(X), eng.to_device(X), [1.0, 4.0], diag_add=5e-4, same_x=True)
for _ in range(3):
    A = K.clone(); eng.potrf(A)
out = (ctypes.c_longlong * 4)()
eng.lib.gpx_debug_leaf_cycles(out)
print("leaf cycles load/factor/inverse/store:", list(out), "total us at 1.965GHz: %.1f" % (sum(out) / 1965.0))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
A = K.clone(); dinv = eng.empty(1, 128, 128)
torch.cuda.synchronize()
e0.record()
for _ in range(50):
    eng.lib.gpx_potrf_async(eng.h, eng._p(A), 128, 128, eng._p(dinv))
e1.record(); torch.cuda.synchronize()
print("leaf launch-to-launch %.1f us" % (e0.elapsed_time(e1) * 1000 / 50))
