"""times the tensor-core actor pair on the env-state path (observation rebuilt + normalised in the kernel, obs_out written): what
the rollout / config-3 step launches"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng
import bench
n = int(os.environ.get("SAT_PROFILE_ENVS", "65536"))
env = eng.EnvBatch(n, mode="rk4", substeps=100, h=1.0, d_capture=20000.0, max_episode_steps=1000)
rng = np.random.default_rng(1)
env.set_state(np.array([2e5, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3, (n, 3)), np.array([1.8e4, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3, (n, 3)))
st = eng.RunningStats(18); st.update_normalize(env.observe())
a = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 0))
b = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 1))
o = torch.empty((n, 18), device="cuda"); A = [torch.empty((n, 3), device="cuda") for _ in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for label, kw in (("state path + obs_out", dict(env=env, obs_stats=st, obs_out=o)), ("state path", dict(env=env, obs_stats=st)), ("obs given", dict(obs=o))):
    ts = []
    for i in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a.sample_pair(b, seed=1, step=2 * i, other_step=2 * i + 1, act=A[0], logp=A[1], other_act=A[2], other_logp=A[3], **kw); e1.record(); e1.synchronize()
        if i >= 4: ts.append(e0.elapsed_time(e1))
    print(f"actor pair, {label}: {np.mean(ts)*1e3:.1f} us (min {np.min(ts)*1e3:.1f})", flush=True)
