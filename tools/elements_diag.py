"""GPU diagnostic: device orbital elements / pow2 vs the CPU oracle, bit for bit, on the states where the danger-zone count
flipped in tools/soak_parity.py (gpurun_out/soak_flips.npz) and on random near-GEO states."""
import os, sys, math
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng, _lib as L
from oracle import oracle as O

u = 3.986e14
def dev_elements(rv):
    rv = torch.from_numpy(np.ascontiguousarray(rv)).cuda()
    out = torch.empty_like(rv); kind = torch.empty(rv.shape[0], dtype=torch.int32, device="cuda")
    L.check(L.load().sat_orbital_elements(rv.data_ptr(), rv.shape[0], u, out.data_ptr(), kind.data_ptr(), L.stream_ptr()), "el")
    return out.cpu().numpy()

np.set_printoptions(precision=17, linewidth=200)
p = os.path.join(ROOT, "gpurun_out", "soak_flips.npz")
sets = []
if os.path.exists(p):
    d = np.load(p)
    sets.append(("flipped", np.concatenate([d["states"][:, 0:6], d["states"][:, 6:12]])))
rng = np.random.default_rng(0)
n = 200000
Rcw = np.array([27098000.0, 32306000.0, 0.0]); Vcw = np.array([-2350.0, 1970.0, 0.0])
sets.append(("random", np.concatenate([Rcw + rng.normal(0, 1e5, (n, 3)), Vcw + rng.normal(0, 20, (n, 3))], axis=1)))
for name, rv in sets:
    dev = dev_elements(rv)
    orc = np.array([O.orbital_elements(u, rv[i, 0:3], rv[i, 3:6])[:6] for i in range(len(rv))])
    diff = dev.view(np.int64) != orc.view(np.int64)
    print(name, "states", len(rv), "element mismatches per column (a,e,i,omega,Omega,f):", diff.sum(axis=0))
    vn = np.array([math.sqrt(math.fma(v[2], v[2], math.fma(v[1], v[1], v[0] * v[0]))) for v in rv[:, 3:6]]) if hasattr(math, "fma") else np.linalg.norm(rv[:, 3:6], axis=1)
    p2 = eng.libm_eval("pow2", torch.from_numpy(vn).cuda()).cpu().numpy()
    ref = np.array([math.pow(float(v), 2.0) for v in vn])
    print("  pow2(v_norm) device vs host libm mismatches:", int((p2.view(np.int64) != ref.view(np.int64)).sum()))
    for i in np.nonzero(diff.any(axis=1))[0][:4]:
        print("  row", i, "\n   dev", dev[i], "\n   orc", orc[i])
