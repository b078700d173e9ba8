import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussian_process_b200 import get_engine
from oracle import gp_oracle as O
eng = get_engine()
def peak(tag):
    print(tag, "DMMA %.2f TF" % eng.fp64_peak(True, 8192)[0], flush=True)
peak("start")
time.sleep(2.0); peak("after sleep 2s")
X, y = O.synth_c5(65536, 16); peak("after synth 65536")
a = torch.empty(int(30e9 // 8), dtype=torch.float64, device="cuda"); peak("after 30GB alloc")
b = torch.empty(int(60e9 // 8), dtype=torch.float64, device="cuda"); peak("after +60GB alloc")
del a, b; torch.cuda.empty_cache(); peak("after free")
A = torch.randn(4096, 4096, device="cuda", dtype=torch.float64); C = torch.zeros(4096, 4096, device="cuda", dtype=torch.float64)
eng.gemm(A, A, C, True, True, 4096, 4096, 4096); torch.cuda.synchronize(); peak("after gemm")
Xd = eng.to_device(X); peak("after to_device")
