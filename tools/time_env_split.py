"""does splitting the device-resident env step into R env ranges on two streams help? (desynchronised phases)"""
import os, sys, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng, _lib as L
n = 65536
env = eng.EnvBatch(n, mode="rk4", substeps=100, h=1.0, d_capture=20000.0, max_episode_steps=1000)
rng = np.random.default_rng(1234)
env.set_state(np.array([200000.0, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3.0, (n, 3)),
              np.array([18000.0, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3.0, (n, 3)))
g = torch.Generator(device="cuda").manual_seed(99)
pa = torch.rand((8, n, 3), generator=g, device="cuda") * 4 - 2
ea = torch.rand((8, n, 3), generator=g, device="cuda") * 4 - 2
obs = torch.empty((n, 18), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
lib = L.load()
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
env.params.action_dtype = L.ACT_F32
for R in (1, 2, 3, 4):
    per = -(-n // R // 64) * 64
    subs = []
    lo = 0
    while lo < n:
        hi = min(n, lo + per)
        st = L.SatEnvState(env._state_buf.data_ptr() + lo * 8, env._istate_buf.data_ptr() + lo * 4, hi - lo, env.ld)
        ws = torch.zeros(lib.sat_workspace_bytes(hi - lo), dtype=torch.uint8, device="cuda")
        subs.append((lo, hi, st, ws)); lo = hi
    ts = []
    for i in range(25):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        a.record()
        for s in streams: s.wait_stream(cur)
        for k, (lo, hi, st, ws) in enumerate(subs):
            s = streams[k % 2]
            L.check(lib.sat_env_step(C.byref(st), pa[i % 8, lo:hi].data_ptr(), ea[i % 8, lo:hi].data_ptr(), None, obs[lo:hi].data_ptr(), None, None,
                                     env.reward[lo:hi].data_ptr(), env.done[lo:hi].data_ptr(), None, None, None, ws.data_ptr(), C.byref(env.params), s.cuda_stream))
        for s in streams: cur.wait_stream(s)
        b.record(); b.synchronize()
        if i >= 5: ts.append(a.elapsed_time(b))
    print(f"ranges={R}: env step {np.mean(ts)*1e3:.1f} us (min {np.min(ts)*1e3:.1f})")
