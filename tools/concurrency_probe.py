"""Do independent factorisations issued through different handles / streams overlap?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussian_process_b200 import get_engine
from gaussian_process_b200.engine import new_engine
from gaussian_process_b200._lib import COV_SE
from oracle import gp_oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
eng = get_engine()
X, y = O.synth_c5(N, 16)
Xd = eng.to_device(X)
Kmat = eng.cov(COV_SE, Xd, Xd, [1.0, 4.0], diag_add=5e-4, same_x=True)
lanes = []
for k in range(K):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        lanes.append((new_engine(0), st, Kmat.clone()))
torch.cuda.synchronize()
def run(concurrent):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k, (e, st, A) in enumerate(lanes):
        s = st if concurrent else lanes[0][1]
        ee = e if concurrent else lanes[0][0]
        with torch.cuda.stream(s):
            A.copy_(Kmat); d = ee.potrf_async(A)
    th = time.perf_counter() - t0
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, th * 1e3
for _ in range(2): run(True); run(False)
print("N=%d K=%d sequential (one stream): %.2f ms (host enqueue %.2f ms)" % ((N, K) + run(False)))
print("N=%d K=%d concurrent (K streams) : %.2f ms (host enqueue %.2f ms)" % ((N, K) + run(True)))
