"""Phase / kernel-class breakdown of one fit(+grad) at a given N (CUDA events inside libgpx)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussian_process_b200 import get_engine
from gaussian_process_b200._lib import COV_SE, check
from oracle import gp_oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
grad = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = get_engine()
X, y = O.synth_c5(N, 16)
for _ in range(2): eng.fit(COV_SE, X, y, [1.0, 4.0], 5e-4, with_grad=bool(grad))
check(eng.lib.gpx_timing_enable(eng.h, 1), "t")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); l0 = eng.launches()
fit = eng.fit(COV_SE, X, y, [1.0, 4.0], 5e-4, with_grad=bool(grad))
e1.record(); torch.cuda.synchronize()
buf = (ctypes.c_double * 16)()
check(eng.lib.gpx_timing_collect(eng.h, buf, 16), "c")
names = ["cov", "potrf", "solve+lml", "trtri", "lauum", "grad"]
print("N=%d total %.2f ms, launches %d | gemm %.2f ms (%d launches, %.1f TF executed) | leaf %.2f ms (%d launches, %.1f us each)" % (
    N, e0.elapsed_time(e1), eng.launches() - l0, buf[0], buf[1], buf[2] / max(buf[0], 1e-9) / 1e9, buf[11], buf[12], 1e3 * buf[11] / max(buf[12], 1)))
print("phases ms:", {n: round(buf[3 + i], 3) for i, n in enumerate(names)})
