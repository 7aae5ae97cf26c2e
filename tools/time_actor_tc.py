"""times the actor sampling kernel at 65 536 rows: FFMA2 path vs the tensor-core path (one network, and the pursuer + evader
pair in one launch)"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng
import bench
n = int(os.environ.get("SAT_PROFILE_ENVS", "65536"))
actor = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 0))
other = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 1))
x = torch.randn((n, 18), device="cuda")
act = torch.empty((n, 3), device="cuda"); lp = torch.empty_like(act); act2 = torch.empty_like(act); lp2 = torch.empty_like(act)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, label):
    ts = []
    for i in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i); e1.record(); e1.synchronize()
        if i >= 4: ts.append(e0.elapsed_time(e1))
    print(f"{label}: {np.mean(ts)*1e3:.1f} us (min {np.min(ts)*1e3:.1f}) for {n} rows", flush=True)
for tc in (False, True):
    timed(lambda i: actor.sample(obs=x, seed=1, step=i, act=act, logp=lp, tc=tc), f"one actor  tc={tc}")
    timed(lambda i: actor.sample_pair(other, obs=x, seed=1, step=2 * i, other_step=2 * i + 1, act=act, logp=lp, other_act=act2, other_logp=lp2, tc=tc), f"actor pair tc={tc}")
