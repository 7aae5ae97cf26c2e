"""single actor launch vs the paired launch (both networks in one grid), per env count"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppo_rl_satellite_b200 import engine as eng
import bench
a = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 0))
b = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 1))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn):
    ts = []
    for i in range(25):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i); e1.record(); e1.synchronize()
        if i >= 5: ts.append(e0.elapsed_time(e1))
    return np.mean(ts) * 1e3
for n in (4096, 8192, 16384, 32768, 65536):
    obs = torch.randn((n, 18), device="cuda")
    A = [torch.empty((n, 3), device="cuda") for _ in range(4)]
    one = t(lambda i: a.sample(obs=obs, seed=1, step=i, act=A[0], logp=A[1]))
    two = t(lambda i: (a.sample(obs=obs, seed=1, step=i, act=A[0], logp=A[1]), b.sample(obs=obs, seed=1, step=i, act=A[2], logp=A[3])))
    pair = t(lambda i: a.sample_pair(b, obs=obs, seed=1, step=i, other_step=i, act=A[0], logp=A[1], other_act=A[2], other_logp=A[3]))
    print(f"rows {n:6d}: one launch {one:6.1f} us, two launches {two:6.1f} us, paired launch {pair:6.1f} us")
