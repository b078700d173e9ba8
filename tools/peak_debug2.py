import ctypes, os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussian_process_b200 import get_engine
eng = get_engine()
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,pstate", "--format=csv,noheader", "-lms", "100"], stdout=subprocess.PIPE, text=True)
t00 = time.time()
def peak(tag, iters=8192):
    tf = ctypes.c_double(); ms = ctypes.c_double()
    eng.lib.gpx_bench_fp64_peak(eng.h, 1, iters, ctypes.byref(tf), ctypes.byref(ms))
    print("%.2fs %s DMMA %.2f TF (%.2f ms)" % (time.time() - t00, tag, tf.value, ms.value), flush=True)
peak("start")
x = np.random.RandomState(0).randn(65536 * 16 * 4); y = np.sin(x).sum()
peak("after numpy work")
peak("again")
time.sleep(0.5)
x = np.random.RandomState(0).randn(65536 * 16 * 4); y = np.sin(x).sum()
for i in range(6): peak("rep%d" % i, 2048)
A = torch.randn(8192, 8192, device="cuda", dtype=torch.float64); C = torch.zeros(8192, 8192, device="cuda", dtype=torch.float64)
for _ in range(60): eng.gemm(A, A, C, True, True, 8192, 8192, 8192)
torch.cuda.synchronize()
for i in range(6): peak("postgemm%d" % i)
p.terminate(); print(p.stdout.read())
