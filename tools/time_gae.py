"""times / profiles the GAE kernel at the config-5 shard (T=2048, N=8192) and at (256, 65536)"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for T, N in ((2048, 8192), (256, 65536)):
    r = torch.randn((T, N), device="cuda"); v = torch.randn((T + 1, N), device="cuda")
    d = (torch.rand((T, N), device="cuda") < 0.02).to(torch.uint8)
    a = torch.empty_like(r); vt = torch.empty_like(r)
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.gae_time_major(r, v, d, adv=a, v_target=vt); e1.record(); e1.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1))
    ms = float(np.mean(ts))
    print(f"T={T} N={N}: {ms*1e3:.1f} us  {17.0*T*N/ms/1e6:.0f} GB/s")
