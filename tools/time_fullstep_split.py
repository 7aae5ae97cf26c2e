"""full fused rollout step (2 actor kernels + env step) as ONE batch on one stream vs TWO independent half-batches on two
streams (FP32 actor work of one half overlaps the FP64 env work of the other)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng
import bench
N = 65536
def make(n, seed):
    env = eng.EnvBatch(n, mode="rk4", substeps=100, h=1.0, d_capture=20000.0, max_episode_steps=1000)
    rng = np.random.default_rng(seed)
    env.set_state(np.array([200000.0, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3.0, (n, 3)),
                  np.array([18000.0, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3.0, (n, 3)))
    st, rs = eng.RunningStats(18), eng.RunningStats(1)
    st.update_normalize(env.observe())
    bufs = dict(act=torch.empty((n, 3), device="cuda"), logp=torch.empty((n, 3), device="cuda"), eact=torch.empty((n, 3), device="cuda"),
                elogp=torch.empty((n, 3), device="cuda"), obs=torch.empty((n, 18), device="cuda"), std=torch.zeros(1, dtype=torch.float64, device="cuda"))
    return env, st, rs, bufs
P = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 0))
E = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 1))
def step(part, t, off):
    env, st, rs, b = part
    P.sample(env=env, obs_stats=st, seed=11, step=t, row_offset=off, act=b["act"], logp=b["logp"], obs_out=b["obs"])
    E.sample(env=env, obs_stats=st, seed=12, step=t, row_offset=off, act=b["eact"], logp=b["elogp"])
    env.step(b["act"], b["eact"], obs_stats=st, ret_stats=rs, ret_std_out=b["std"])
K = 20
for R in (1, 2, 4):
    parts = [make(N // R, 7 + r) for r in range(R)]
    streams = [torch.cuda.Stream() for _ in range(R)]
    def run(k0):
        cur = torch.cuda.current_stream()
        for s in streams: s.wait_stream(cur)
        for t in range(K):
            for r in range(R):
                with torch.cuda.stream(streams[r]):
                    step(parts[r], k0 + t, r * (N // R))
        for s in streams: cur.wait_stream(s)
    run(0); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(K); b.record(); b.synchronize()
    ms = a.elapsed_time(b) / K
    print(f"independent parts={R}: {ms*1e3:.0f} us per step of all {N} envs -> {N/ms/1e-3:.3e} env-steps/s")
