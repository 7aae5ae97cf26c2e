"""tuning aid: the tensor-core forward/backward kernel of the fused PPO step (csrc/ppo_fb_tc.cu) against the fp32 FFMA2 kernel on the
same minibatch: workspace buffers (h1, dz2, dz1, xs) and per-group gradients"""
import os, sys, types, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ppo_rl_satellite_b200 import engine as E, _lib as L
import test_gpu_ppo_fused as T
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
use_tanh = int(sys.argv[2]) if len(sys.argv) > 2 else 1
args = T._args(use_tanh=use_tanh)
fused, eager = T._pair(args)
B = 1500
s, act, logp, adv, vt = T._data(eager, B)
index = torch.randperm(B, device="cuda")[:mb].contiguous()
lib = L.load()
f = fused._fused_for(mb)
na, nc = f["nets"]
mp = (mb + 127) // 128 * 128
out = {}
for tc in (0, 1):
    lib.sat_ppo_use_tensor_cores(tc)
    res = {}
    for name, net in (("actor", na), ("critic", nc)):
        net.workspace.zero_()
        if name == "actor": net.actor_grad(s, act, logp, adv.reshape(-1), index.data_ptr(), mb, 0.1, 0.01)
        else: net.critic_grad(s, vt.reshape(-1), index.data_ptr(), mb)
        torch.cuda.synchronize()
        ws = net.workspace
        res[name] = dict(h1=ws[:mp * 256].clone(), dz2=ws[mp * 256:2 * mp * 256].clone(), dz1=ws[2 * mp * 256:3 * mp * 256].clone(),
                         xs=ws[3 * mp * 256:3 * mp * 256 + mp * 32].clone(), grads=net.grads.clone(),
                         scal=ws[3 * mp * 256 + mp * 32 + (mp // 64) * 1024:3 * mp * 256 + mp * 32 + (mp // 64) * 1024 + (mp // 64) * 8].clone())
    out[tc] = res
for name, net in (("actor", na), ("critic", nc)):
    print("==", name)
    for k in ("xs", "h1", "dz2", "dz1"):
        a, b = out[0][name][k], out[1][name][k]
        d = (a - b).abs()
        print(f"  {k:4s}: max |ffma| {a.abs().max().item():.3e}  max diff {d.max().item():.3e} at {int(d.argmax())}")
    heads = 3 if name == "actor" else 1
    offs = [("W1", 0, 256 * 18), ("b1", 256 * 18, 256), ("W2", 256 * 18 + 256, 65536), ("b2", 256 * 18 + 256 + 65536, 256),
            ("W3", 256 * 18 + 512 + 65536, heads * 256), ("tail", 256 * 18 + 512 + 65536 + heads * 256, 8)]
    for nm, o, c in offs:
        a, b = out[0][name]["grads"][o:o + c], out[1][name]["grads"][o:o + c]
        print(f"  grad {nm:4s}: max |ffma| {a.abs().max().item():.3e}  max diff {(a - b).abs().max().item():.3e}")
for name, net in (("actor", na), ("critic", nc)):
    heads = 3 if name == "actor" else 1
    o = 256 * 18 + 512 + 65536 + heads * 256
    print(name, "tail ffma", out[0][name]["grads"][o:o + 8].tolist())
    print(name, "tail tc  ", out[1][name]["grads"][o:o + 8].tolist())
    o3 = 256 * 18 + 512 + 65536
    print(name, "W3 ffma", out[0][name]["grads"][o3:o3 + 6].tolist(), "tc", out[1][name]["grads"][o3:o3 + 6].tolist())
    ob2 = 256 * 18 + 256 + 65536
    print(name, "b2 ffma", out[0][name]["grads"][ob2:ob2 + 4].tolist(), "tc", out[1][name]["grads"][ob2:ob2 + 4].tolist())

for name in ("actor", "critic"):
    a = out[0][name]["scal"].view(-1, 8); b = out[1][name]["scal"].view(-1, 8)[:mp // 128]
    a2 = a.view(-1, 2, 8).sum(1)
    print(name, "per-128-row-tile scalars ffma:\n", a2[:, :7].cpu().numpy().round(5), "\n tc:\n", b[:, :7].cpu().numpy().round(5))
