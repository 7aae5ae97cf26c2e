"""GPU diagnostic: where do device danger-zone counts differ from the oracle? Writes gpurun_out/dz_diag.npz"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng
from oracle import oracle as O

g = np.load(os.path.join(ROOT, "tests/golden/danger_golden.npz"))
S, fuel, cnt = g["dz_states"], g["dz_fuel"], g["dz_count"]
out, dbg = eng.danger_zone_count(torch.from_numpy(S).cuda(), torch.from_numpy(fuel).cuda(), debug=True)
out = out.cpu().numpy(); dbg = dbg.cpu().numpy()
print("golden danger set: mismatches", int((out != cnt).sum()), "of", len(cnt))

# batched rollout: collect post-step states where device and oracle disagree
eg = np.load(os.path.join(ROOT, "tests/golden/env_golden.npz"))
n, T = 4096, 200
kw = dict(d_capture=181200.0, max_episode_steps=40, flag=0)
env = eng.EnvBatch(n, mode="cw", auto_reset=False, stm=eg["stm100_columns"], **kw)
rng = np.random.default_rng(100)
Rcw = np.array([27098000.0, 32306000.0, 0.0]); Vcw = np.array([-2350.0, 1970.0, 0.0])
bad_states, bad_fuel, dev_cnt, orc_cnt, dev_dbg, orc_dbg = [], [], [], [], [], []
obs = torch.empty((n, 18), dtype=torch.float64, device="cuda")
total = 0
for t in range(T):
    pa = torch.from_numpy(rng.uniform(-2, 2, (n, 3)).astype(np.float32)).cuda()
    ea = torch.from_numpy(rng.uniform(-2, 2, (n, 3)).astype(np.float32)).cuda()
    r, d = env.step(pa, ea, obs_f64=obs)
    o = obs.cpu().numpy(); fc = env.fuel_c.cpu().numpy(); dn = d.cpu().numpy().astype(bool)
    st = np.concatenate([Rcw + o[:, 6:9], Vcw + o[:, 9:12], Rcw + o[:, 12:15], Vcw + o[:, 15:18]], axis=1)
    dz = env.dangerous_zone.cpu().numpy()
    oc = np.array([O.danger_zone(st[i, 0:3], st[i, 3:6], st[i, 6:9], st[i, 9:12], fc[i]) for i in range(n)])
    live = ~dn
    total += int(live.sum())
    bad = live & (dz != oc)
    for i in np.nonzero(bad)[0]:
        bad_states.append(st[i]); bad_fuel.append(fc[i]); dev_cnt.append(dz[i]); orc_cnt.append(oc[i])
    if dn.any():
        env.reset(mask=d)
        env.istate[1][d.bool()] = 0
print("rollout: mismatching danger-zone evaluations", len(bad_states), "of", total)
if bad_states:
    bs = np.array(bad_states); bf = np.array(bad_fuel)
    o2, d2 = eng.danger_zone_count(torch.from_numpy(bs).cuda(), torch.from_numpy(bf).cuda(), debug=True)
    for i in range(len(bs)):
        c, db = O.danger_zone_debug(bs[i, 0:3], bs[i, 3:6], bs[i, 6:9], bs[i, 9:12], bf[i])
        orc_dbg.append(db)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez(os.path.join(ROOT, "gpurun_out/dz_diag.npz"), states=bs, fuel=bf, dev_cnt=np.array(dev_cnt), orc_cnt=np.array(orc_cnt),
             dev_cnt2=o2.cpu().numpy(), dev_dbg=d2.cpu().numpy(), orc_dbg=np.array(orc_dbg))
    np.set_printoptions(precision=17, linewidth=200)
    for i in range(min(6, len(bs))):
        print("case", i, "dev", dev_cnt[i], "orc", orc_cnt[i], "fuel", bf[i])
        print(" dev", d2[i].cpu().numpy()); print(" orc", orc_dbg[i])
