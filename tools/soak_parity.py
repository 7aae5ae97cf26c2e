"""Large-sample parity soak (cw mode, the env as shipped): N envs x T steps on the GPU against the CPU oracle with the same
random fp32 actions. Reports the number of envs whose integer danger-zone count ever differed from the oracle (target 0) and
checks that every other env is bit-identical in observation, reward and done. Every differing evaluation is re-run through
sat_danger_zone_count / the oracle's debug entry and dumped. Writes gpurun_out/soak_parity.txt (+ soak_flips.npz)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
T = int(sys.argv[2]) if len(sys.argv) > 2 else 40
g = np.load(os.path.join(ROOT, "tests/golden/env_golden.npz"))
out = []
Rcw = np.array([27098000.0, 32306000.0, 0.0]); Vcw = np.array([-2350.0, 1970.0, 0.0])
flip_states, flip_fuel = [], []
for flag in (0, 1):
    kw = dict(d_capture=181200.0, max_episode_steps=40, flag=flag)
    env = eng.EnvBatch(n, mode="cw", auto_reset=True, stm=g["stm100_columns"], **kw)
    orc = O.BatchEnv(n, nthreads=os.cpu_count(), M=g["stm100_columns"], **kw)
    rng = np.random.default_rng(1000 + flag)
    obs = torch.empty((n, 18), dtype=torch.float64, device="cuda")
    flipped = np.zeros(n, dtype=bool)
    evals = dones = bad = 0
    t0 = time.time()
    for t in range(T):
        pa = rng.uniform(-2, 2, (n, 3)).astype(np.float32); ea = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
        r, d = env.step(torch.from_numpy(pa).cuda(), torch.from_numpy(ea).cuda(), obs_f64=obs)
        o_obs, o_r, o_d = orc.step(pa.astype(np.float64), ea.astype(np.float64))
        dz, o_dz = env.dangerous_zone.cpu().numpy(), orc.aux()[3]
        dn = d.cpu().numpy().astype(bool)
        evals += int((~dn).sum()); dones += int(dn.sum())
        new = (dz != o_dz) & ~flipped & ~dn
        if new.any():
            o = o_obs[new]
            flip_states.append(np.concatenate([Rcw + o[:, 6:9], Vcw + o[:, 9:12], Rcw + o[:, 12:15], Vcw + o[:, 15:18]], axis=1))
            flip_fuel.append(env.fuel_c.cpu().numpy()[new])
        flipped |= dz != o_dz
        ok = ~flipped
        same = (np.array_equal(dn[ok], o_d[ok].astype(bool)) and np.array_equal(r.cpu().numpy()[ok], o_r[ok])
                and np.array_equal(obs.cpu().numpy()[ok], o_obs[ok]))
        bad += 0 if same else 1
    line = (f"flag {flag}: {n} envs x {T} steps, {evals} danger-zone evaluations, {dones} episode ends; envs whose count ever "
            f"differed from the oracle: {int(flipped.sum())} ({flipped.sum() / max(1, evals):.2e} per evaluation); steps on which a "
            f"non-flipped env differed in obs/reward/done: {bad}; {time.time() - t0:.0f} s")
    print(line, flush=True); out.append(line)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
if flip_states:
    bs, bf = np.concatenate(flip_states), np.concatenate(flip_fuel)
    cnt, dbg = eng.danger_zone_count(torch.from_numpy(bs).cuda(), torch.from_numpy(bf).cuda(), debug=True)
    odbg = [O.danger_zone_debug(bs[i, 0:3], bs[i, 3:6], bs[i, 6:9], bs[i, 9:12], bf[i]) for i in range(len(bs))]
    np.savez(os.path.join(ROOT, "gpurun_out", "soak_flips.npz"), states=bs, fuel=bf, dev_cnt=cnt.cpu().numpy(), dev_dbg=dbg.cpu().numpy(),
             orc_cnt=np.array([c for c, _ in odbg]), orc_dbg=np.array([d for _, d in odbg]))
    np.set_printoptions(precision=17, linewidth=220)
    for i in range(min(8, len(bs))):
        line = f"flip {i}: standalone device count {int(cnt[i])}, oracle {odbg[i][0]}, fuel {bf[i]!r}\n dev {dbg[i].cpu().numpy()}\n orc {odbg[i][1]}"
        print(line, flush=True); out.append(line)
open(os.path.join(ROOT, "gpurun_out", "soak_parity.txt"), "w").write("\n".join(out) + "\n")
