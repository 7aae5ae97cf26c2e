"""tuning aid: timeline of CTA 0 / thread 0 of ppo_fb_tc_kernel (library built with SAT_NVCC_DEFINES=SAT_TC_TRACE)"""
import os, sys, ctypes as C, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
exec(open(os.path.join(ROOT, "tools", "profile_ppo_fused.py")).read().split('print("ok")')[0])
from ppo_rl_satellite_b200 import _lib as L
lib = L.load()
buf = (C.c_ulonglong * 256)()
lib.sat_debug_fb_tc_trace.argtypes = [C.c_void_p]
assert lib.sat_debug_fb_tc_trace(buf) == 0
t = np.array(buf, dtype=np.int64)
names = ["start", "l1_full", "h1 chunks done", "h1 stored", "l2_full", "head sums (bar1)", "loss (bar2)", "dz2 chunks done", "dz2 stored", "next x", "pre-l3", "l3_full", "dz1 stored"]
for tile in range(4):
    v = (t[tile * 16:tile * 16 + 13] - t[0]) / 1e3
    print("tile", tile, " ".join(f"{n}={x:.1f}" for n, x in zip(names, v)))
