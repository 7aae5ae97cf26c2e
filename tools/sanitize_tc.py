"""small invocations of the tensor-core kernels (small-case smoke run; compute-sanitizer memcheck where the pool allows it): actor (one network / pair, obs and env-state
paths, ragged sizes) and the fused PPO step (actor + critic, tensor-core forward/backward + dW2) at ragged minibatch sizes"""
import os, sys, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ppo_rl_satellite_b200 import engine as eng
import bench
rng = np.random.default_rng(0)
a = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 0))
b = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 1))
for n in (1, 127, 129, 300):
    env = eng.EnvBatch(n, mode="cw", d_capture=20000.0, max_episode_steps=10)
    env.set_state(np.array([2e5, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3, (n, 3)), np.array([1.8e4, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3, (n, 3)))
    st = eng.RunningStats(18); st.update_normalize(env.observe())
    o = torch.empty((n, 18), device="cuda")
    a.sample(env=env, obs_stats=st, seed=1, step=2, obs_out=o, tc=True)
    a.sample(obs=o, seed=1, step=2, tc=True, mean_out=torch.empty((n, 3), device="cuda"), eps_out=torch.empty((n, 3), device="cuda"))
    a.sample_pair(b, env=env, obs_stats=st, seed=1, step=2, other_step=3, obs_out=o, tc=True)
    a.sample_pair(b, obs=o, seed=1, step=2, other_step=3, tc=True)
import test_gpu_ppo_fused as T
for mb in (37, 128, 300):
    args = T._args(use_tanh=1)
    fused, eager = T._pair(args)
    s, act, logp, adv, vt = T._data(eager, 400)
    index = torch.randperm(400, device="cuda")[:mb].contiguous()
    f = fused._fused_for(mb)
    na, nc = f["nets"]
    na.actor_grad(s, act, logp, adv.reshape(-1), index.data_ptr(), mb, 0.1, 0.01)
    nc.critic_grad(s, vt.reshape(-1), index.data_ptr(), mb)
torch.cuda.synchronize()
print("sanitize_tc done")
