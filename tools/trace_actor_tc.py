"""tuning aid: timeline of CTA 0 of actor_tc_kernel (library built with SAT_NVCC_DEFINES=SAT_TC_TRACE)"""
import os, sys, ctypes as C, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng, _lib as L
import bench
n = 65536
actor = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 0))
x = torch.randn((n, 18), device="cuda")
if os.environ.get("SAT_TRACE_STATE", "0") == "1":
    env = eng.EnvBatch(n, mode="rk4", substeps=100, h=1.0, d_capture=20000.0, max_episode_steps=1000)
    rng = np.random.default_rng(1)
    env.set_state(np.array([2e5, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3, (n, 3)), np.array([1.8e4, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3, (n, 3)))
    st = eng.RunningStats(18); st.update_normalize(env.observe())
    kw = dict(env=env, obs_stats=st, obs_out=x)
else:
    kw = dict(obs=x)
for i in range(5):
    actor.sample(seed=1, step=i, tc=True, **kw)
torch.cuda.synchronize()
lib = L.load()
buf = (C.c_ulonglong * 512)()
lib.sat_debug_actor_tc_trace.argtypes = [C.c_void_p]
assert lib.sat_debug_actor_tc_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(2, 256)
t0 = t[1, 0]
mma = (t[0] - t0) / 1e3
cmp_ = (t[1] - t0) / 1e3
print("MMA lane: per chunk (before a_full wait, after wait, after issue) us")
for tile in range(4):
    for c in range(9):
        k = (tile * 9 + c) * 3
        print(f"  t{tile} c{c}: {mma[k]:8.2f} {mma[k+1]:8.2f} {mma[k+2]:8.2f}")
names = ["pre-l1_full", "l1_full", "R done"] + [f"chunk{k} arrived" for k in range(1, 9)] + ["pre-l2_full", "l2_full", "acc_free", "tile end"]
print("compute thread 0:")
for tile in range(4):
    print("  tile", tile, " ".join(f"{names[j]}={cmp_[tile*16+j]:.2f}" for j in range(15)), f"X stage free={cmp_[tile*16+15]:.2f}")
