"""Small-shape tour of every hand-written kernel class for compute-sanitizer (memcheck / racecheck): TMA-fed DMMA GEMM
(mbarrier ring), cp.async GEMM, potrf leaf, look-ahead potrf, solves, inverse, covariance / gradient, fused small-problem
kernels, Laplace steps, the emulated multi-rank driver.  Results are checked loosely against NumPy so that a sanitizer-clean
run is also a correct run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gaussian_process_b200 import get_engine, synthetic as S
from gaussian_process_b200._lib import COV_CO2, COV_SE

eng = get_engine()
rs = np.random.RandomState(0)
# GEMM, all four layouts (TMA kernel) + batched (cp.async kernel) through potrf / trtri
a = eng.to_device(rs.randn(256, 128))
for akm, bkm in ((1, 1), (1, 0), (0, 1), (0, 0)):
    A = a if akm else a.t().contiguous()
    B = a if bkm else a.t().contiguous()
    C = eng.zeros(256, 256)
    eng.gemm(A, B, C, bool(akm), bool(bkm), 256, 256, 128)
    assert np.allclose(eng.to_host(C), eng.to_host(a) @ eng.to_host(a).T, atol=1e-10)
# fit + gradient at N=1300 (look-ahead potrf on 3 streams, level-batched trtri, lauum, TRSVs with block inverses, gradient)
X, y = S.synth_c5(1300, 16)
fit = eng.fit(COV_SE, X, y, [1.0, 4.0], 5e-4, with_grad=True)
fit2 = eng.fit(COV_SE, X, y, [1.0, 4.0], 5e-4)
mu, var, _ = eng.predict(fit2, X[:200] + 0.01)
# CO2 kernel, all 11 derivatives
Xc, yc, Xs = S.synth_c2(300)
fc = eng.fit(COV_CO2, Xc, yc, [66, 67, 2.4, 90, 1.3, .66, 1.2, .78, .18, 1.6, .19], 5e-4, with_grad=True)
# fused small-problem kernels
Xa, ya, Xsa = S.synth_c1(12, 100)
eng.small_posterior(COV_SE, Xa, ya, Xsa, [1.0, 1.0], 5e-4, 1e-6, rs.randn(100, 3))
eng.small_lml_grad(COV_SE, Xa, ya, [1.0, 1.0], 5e-4)
eng.small_ascent(Xa, ya, 1.0, 0.6, 5e-4, 0.01, 1e-3, 50)
# device-resident optimiser (CUDA-graph replay)
eng.ascend(COV_SE, X[:400], y[:400], [1.0, 2.0], [1, 1], 5e-4, 1e-4, 1e-3, 4)
# Laplace steps
from gaussian_process_b200.laplace import BinaryLaplace, MultiLaplaceNewton
Xb, yb, fpr = S.synth_c3(500, 8)
Kb = eng.cov(COV_SE, eng.to_device(Xb), eng.to_device(Xb), [1.0, 1.0], same_x=True)
mb = BinaryLaplace(eng, Kb, 500)
mb.fit_newton(yb, tolerance=1e-8, max_iter=4)
BinaryLaplace(eng, Kb, 500).fit_reference(yb, fpr, tolerance=1e-3, max_iter=20)
Xm, lab, ym, _, _ = S.synth_c4(200, 4, 6, 8)
Km = eng.cov(COV_SE, eng.to_device(Xm), eng.to_device(Xm), [1.0, 1.0], same_x=True)
MultiLaplaceNewton(eng, Km, 4, 200, concurrency=2).fit(ym, tolerance=1e-6, max_iter=3)
# emulated 3-rank block-cyclic driver
eng.mg_emulate_fit_grad(3, COV_SE, X[:900], y[:900], [1.0, 4.0], 5e-4, nb=128)
torch.cuda.synchronize()
print("SANITIZE_TOUR_OK lml=%.6f" % fit.lml)
