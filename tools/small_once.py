"""One fused small-posterior call at the as-shipped size (for ncu captures of gp_small_kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from gaussian_process_b200 import get_engine
from gaussian_process_b200._lib import COV_SE

eng = get_engine()
rs = np.random.RandomState(0)
N, n, nf = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (5, 100, 10)))
X = rs.uniform(-5, 5, (N, 1))
y = np.sin(X).ravel()
Xs = np.linspace(-5, 5, n).reshape(-1, 1)
Z = rs.randn(n, nf)
for _ in range(3):
    mu, var, fp, lml = eng.small_posterior(COV_SE, X, y, Xs, [1.0, 1.0], 5e-4, 1e-6, Z)
print("ok", float(mu[0]), lml)
