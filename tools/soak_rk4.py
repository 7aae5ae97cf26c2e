"""rk4-mode soak (the north-star production mode): N envs x T env steps x S RK4+J2 substeps on the GPU against the CPU oracle
(reference step logic composed with the script's RungeKutta). Reports the worst relative state error in the inertial frame,
done mismatches and reward differences on envs whose danger-zone history agrees. Appends to gpurun_out/soak_parity.txt."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
S = int(sys.argv[3]) if len(sys.argv) > 3 else 100
kw = dict(d_capture=20000.0, max_episode_steps=12)
env = eng.EnvBatch(n, mode="rk4", substeps=S, h=1.0, auto_reset=True, **kw)
orc = O.BatchEnv(n, nthreads=os.cpu_count(), **kw)
rng = np.random.default_rng(42)
obs = torch.empty((n, 18), dtype=torch.float64, device="cuda")
R = np.array([27098000.0, 32306000.0, 0.0]); V = np.array([-2350.0, 1970.0, 0.0])
diverged = np.zeros(n, dtype=bool)
worst_r = worst_v = worst_rew = 0.0
done_mismatch = dones = 0
t0 = time.time()
for t in range(T):
    pa = rng.uniform(-2, 2, (n, 3)).astype(np.float32); ea = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    r, d = env.step(torch.from_numpy(pa).cuda(), torch.from_numpy(ea).cuda(), obs_f64=obs)
    o_obs, o_r, o_d = orc.step_rk4(pa.astype(np.float64), ea.astype(np.float64), h=1.0, substeps=S)
    dz_same = env.dangerous_zone.cpu().numpy() == orc.aux()[3]
    ok = ~diverged
    dn = d.cpu().numpy()
    done_mismatch += int((dn[ok] != o_d[ok]).sum()); dones += int(dn.sum())
    got = obs.cpu().numpy()
    for lo in (6, 12):
        dr = np.linalg.norm(got[ok, lo:lo + 3] - o_obs[ok, lo:lo + 3], axis=1) / np.linalg.norm(o_obs[ok, lo:lo + 3] + R, axis=1)
        dv = np.linalg.norm(got[ok, lo + 3:lo + 6] - o_obs[ok, lo + 3:lo + 6], axis=1) / np.linalg.norm(o_obs[ok, lo + 3:lo + 6] + V, axis=1)
        worst_r, worst_v = max(worst_r, float(dr.max())), max(worst_v, float(dv.max()))
    both = ok & dz_same
    worst_rew = max(worst_rew, float(np.abs(r.cpu().numpy()[both] - o_r[both]).max()))
    diverged |= ~dz_same
line = (f"rk4 mode: {n} envs x {T} steps x {S} RK4+J2 substeps ({2 * n * T * S} RK4 steps), {dones} episode ends: worst relative "
        f"error position {worst_r:.2e}, velocity {worst_v:.2e} (bar 1e-9); done mismatches {done_mismatch}; worst |reward difference| "
        f"{worst_rew:.2e} on envs with the same danger-zone history; envs excluded after a danger-zone flip: {int(diverged.sum())}; "
        f"{time.time() - t0:.0f} s")
print(line, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "soak_rk4.txt"), "w").write(line + "\n")
