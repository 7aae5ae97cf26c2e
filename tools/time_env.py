"""times the rk4-mode env step (kernel A + kernel B) alone, CUDA events, L2 flushed between launches"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng
n = int(os.environ.get("SAT_PROFILE_ENVS", "65536"))
env = eng.EnvBatch(n, mode="rk4", substeps=100, h=1.0, d_capture=20000.0, max_episode_steps=1000)
rng = np.random.default_rng(1234)
env.set_state(np.array([200000.0, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3.0, (n, 3)),
              np.array([18000.0, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3.0, (n, 3)))
g = torch.Generator(device="cuda").manual_seed(99)
pa = torch.rand((8, n, 3), generator=g, device="cuda") * 4 - 2
ea = torch.rand((8, n, 3), generator=g, device="cuda") * 4 - 2
obs = torch.empty((n, 18), dtype=torch.float32, device="cuda")
st, rs = eng.RunningStats(18), eng.RunningStats(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(25):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env.step(pa[i % 8], ea[i % 8], obs_f32=obs, obs_stats=st, ret_stats=rs); b.record(); b.synchronize()
    if i >= 5:
        ts.append(a.elapsed_time(b))
print(f"MINB={os.environ.get('SAT_FINISH_MINB','1')} envs={n} env step: mean {np.mean(ts)*1e3:.1f} us  min {np.min(ts)*1e3:.1f} us  -> {n/np.mean(ts)/1e-3:.3e} env-steps/s")
