"""PPO optimiser step at minibatch mb: fused CUDA kernels vs the CUDA-graphed PyTorch step"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ppo_rl_satellite_b200.dropin import ppo_continuous as P

mbs = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["65536", "75776"])]
B = 1 << 21
a = bench._PpoArgs(policy_dist="Gaussian", max_action=1.6, batch_size=B, mini_batch_size=65536, max_train_steps=int(3e6),
                   lr_a=2e-4, lr_c=2e-4, gamma=0.99, lamda=0.95, epsilon=0.1, K_epochs=1, entropy_coef=0.01, set_adam_eps=True,
                   use_grad_clip=True, use_lr_decay=True, use_adv_norm=True, state_dim=18, action_dim=3, hidden_width=256,
                   use_tanh=True, use_orthogonal_init=True, chkpt_dir="/tmp")
s = torch.randn(B, 18, device="cuda"); act = torch.randn(B, 3, device="cuda").clamp(-1.6, 1.6); lp = torch.randn(B, 3, device="cuda") * 0.1 - 1.0
adv = torch.randn(B, 1, device="cuda"); vt = torch.randn(B, 1, device="cuda")
for mb in mbs:
    for mode in ("fused", "graph"):
        agent = P.PPO_continuous(a, "pursuer")
        kw = dict(fused=True) if mode == "fused" else dict(use_graph=True)
        agent.optimize(s, act, lp, adv, vt, mini_batch_size=mb, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); agent.optimize(s, act, lp, adv, vt, mini_batch_size=mb, **kw); e1.record(); e1.synchronize()
        steps = -(-B // mb)
        print(f"mb {mb} {mode}: {e0.elapsed_time(e1) / steps:.3f} ms per optimiser step ({steps} steps, host {1e3 * (time.perf_counter() - t0) / steps:.3f} ms), "
              f"{B / (e0.elapsed_time(e1) * 1e-3):.3e} samples/s/epoch")
