"""Short driver for ncu: a few fused rollout steps (2 actor samples + env step in rk4 mode) and one K1 launch.
Same kernels, shapes and launch parameters as bench.py's timed region (65 536 envs, S = 100)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng
import bench

n = int(os.environ.get("SAT_PROFILE_ENVS", "65536")); S = 100
env = eng.EnvBatch(n, mode="rk4", substeps=S, h=1.0, d_capture=20000.0, max_episode_steps=1000)
rng = np.random.default_rng(1234)
env.set_state(np.array([200000.0, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3.0, (n, 3)),
              np.array([18000.0, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3.0, (n, 3)))
pursuer = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 0))
evader = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 1))
obs_stats, ret_stats = eng.RunningStats(18), eng.RunningStats(1)
obs_stats.update_normalize(env.observe())
act = torch.empty((n, 3), dtype=torch.float32, device="cuda"); logp = torch.empty_like(act)
eact = torch.empty_like(act); elogp = torch.empty_like(act)
obs = torch.empty((n, 18), dtype=torch.float32, device="cuda")
std = torch.zeros(1, dtype=torch.float64, device="cuda")
for t in range(int(os.environ.get("SAT_PROFILE_STEPS", "4"))):
    pursuer.sample(env=env, obs_stats=obs_stats, seed=11, step=t, act=act, logp=logp, obs_out=obs)
    evader.sample(env=env, obs_stats=obs_stats, seed=12, step=t, act=eact, logp=elogp)
    env.step(act, eact, obs_f32=obs, obs_stats=obs_stats, ret_stats=ret_stats, ret_std_out=std)
x, _ = eng.alloc_soa(6, 1 << 20, torch.float64, "cuda")
ang = torch.rand(1 << 20, device="cuda", dtype=torch.float64) * 6.283185307179586
x[0], x[1], x[2] = 7000 * torch.cos(ang), 7000 * torch.sin(ang), 100 * torch.randn(1 << 20, device="cuda", dtype=torch.float64)
x[3], x[4], x[5] = -7.5 * torch.sin(ang), 7.5 * torch.cos(ang), 0.1 * torch.randn(1 << 20, device="cuda", dtype=torch.float64)
eng.rk4_propagate(x, 1.0, 100)
torch.cuda.synchronize()
print("profile_step done", float(env.reward.mean()))
