"""Sustained vs burst FP64 DMMA peak: run the register-resident DMMA loop for several seconds and sample clocks/power."""
import ctypes, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_b200 import get_engine
eng = get_engine()
q = "clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu"
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + q, "--format=csv,noheader", "-lms", "250"], stdout=subprocess.PIPE, text=True)
t0 = time.time()
res = []
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dur = float(sys.argv[2]) if len(sys.argv) > 2 else 6.0
while time.time() - t0 < dur:
    tf = ctypes.c_double(); ms = ctypes.c_double()
    eng.lib.gpx_bench_fp64_peak(eng.h, mode, 8192, ctypes.byref(tf), ctypes.byref(ms))
    res.append((round(time.time() - t0, 2), round(tf.value, 2)))
p.terminate()
out = p.stdout.read().strip().splitlines()
print("mode", mode, "TF over time:", res[::max(1, len(res)//12)])
print("smi samples:", out[::max(1, len(out)//10)])
