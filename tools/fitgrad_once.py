"""One warm-up + one fit + LML + gradient at a given N (the command of the ncu captures of the non-GEMM kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gaussian_process_b200 import get_engine, synthetic as S
from gaussian_process_b200._lib import COV_CO2, COV_SE

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
kind = sys.argv[2] if len(sys.argv) > 2 else "se"
eng = get_engine()
if kind == "se":
    X, y = S.synth_c5(N, 16)
    args = (COV_SE, X, y, [1.0, 4.0], 5e-4)
else:
    X, y, _ = S.synth_c2(N)
    args = (COV_CO2, X, y, [66, 67, 2.4, 90, 1.3, .66, 1.2, .78, .18, 1.6, .19], 5e-4)
for _ in range(2):
    fit = eng.fit(*args, with_grad=True)
torch.cuda.synchronize()
print("N=%d %s lml=%.6f grad=%s" % (N, kind, fit.lml, fit.grad[:2]))
