import os, sys, time, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng, _lib as L
n = 65536
env = eng.EnvBatch(n, mode="rk4", substeps=100, h=1.0, d_capture=20000.0, max_episode_steps=1000)
pa, ea, obs, rew, done = env.host_buffers()
hb = env._host
rng = np.random.default_rng(0)
pa[...] = rng.uniform(-2, 2, (n, 3)); ea[...] = rng.uniform(-2, 2, (n, 3))
lib = L.load(); env.params.action_dtype = L.ACT_F32
streams = [torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()]
def run(bounds):
    subs = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        st = L.SatEnvState(env._state_buf.data_ptr() + lo * 8, env._istate_buf.data_ptr() + lo * 4, hi - lo, env.ld)
        ws = torch.zeros(lib.sat_workspace_bytes(hi - lo), dtype=torch.uint8, device="cuda")
        subs.append((lo, hi, st, ws))
    def step():
        for k, (lo, hi, st, ws) in enumerate(subs):
            s = streams[k % len(streams)]
            L.check(lib.sat_env_step(C.byref(st), hb["pa"].data_ptr() + lo * 12, hb["ea"].data_ptr() + lo * 12, None, hb["obs"].data_ptr() + lo * 72, None, None,
                                     hb["rew"].data_ptr() + lo * 8, hb["done"].data_ptr() + lo, None, None, None, ws.data_ptr(), C.byref(env.params), s.cuda_stream))
        for s in streams: s.synchronize()
    for _ in range(3): step()
    t0 = time.perf_counter()
    for _ in range(20): step()
    return (time.perf_counter() - t0) / 20
for name, b in (("1", [0, n]), ("50/50", [0, n // 2, n]), ("62/38", [0, 40960, n]), ("75/25", [0, 49152, n]), ("50/25/25", [0, 32768, 49152, n]), ("56/31/13", [0, 36864, 57344, n])):
    dt = run(b)
    print(f"ranges {name}: {dt*1e6:.0f} us/step -> {n/dt:.3e}")
