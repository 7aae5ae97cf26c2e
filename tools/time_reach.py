"""time the batched reachable-domain sweep (201 x 201 directions per state)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppo_rl_satellite_b200 import engine as E
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "elements_golden.npz"))
el = torch.from_numpy(g["elements_live"][:800]).cuda(); dm = torch.from_numpy(g["fuel"][:800]).cuda()
for n in (1, 64, 800):
    E.reachable_domain(el[:n], dm[:n]); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): hi, lo, v = E.reachable_domain(el[:n], dm[:n])
    b.record(); torch.cuda.synchronize()
    print(f"states {n}: {a.elapsed_time(b) / 5:.3f} ms per sweep, valid directions {int(v.sum())}")
