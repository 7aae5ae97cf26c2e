import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng
import bench
n = 65536
actor = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 0))
obs = torch.randn((n, 18), device="cuda")
act = torch.empty((n, 3), device="cuda"); logp = torch.empty_like(act)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(25):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); actor.sample(obs=obs, seed=1, step=i, act=act, logp=logp); b.record(); b.synchronize()
    if i >= 5: ts.append(a.elapsed_time(b))
t = np.mean(ts)
print(f"actor: {t*1e3:.1f} us  {141824*n/t/1e9:.1f} TFLOP/s; peaks fp32 {eng.measure_vector_peak('fp32'):.1f} fp32x2 {eng.measure_vector_peak('fp32x2'):.1f} fp64 {eng.measure_vector_peak('fp64'):.1f}")
