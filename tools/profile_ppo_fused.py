"""four fused PPO optimiser steps at minibatch 65536 (for an ncu launch list / --set full capture)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ppo_rl_satellite_b200.dropin import ppo_continuous as P
B, mb = 1 << 18, 65536
a = bench._PpoArgs(policy_dist="Gaussian", max_action=1.6, batch_size=B, mini_batch_size=mb, max_train_steps=int(3e6),
                   lr_a=2e-4, lr_c=2e-4, gamma=0.99, lamda=0.95, epsilon=0.1, K_epochs=1, entropy_coef=0.01, set_adam_eps=True,
                   use_grad_clip=True, use_lr_decay=True, use_adv_norm=True, state_dim=18, action_dim=3, hidden_width=256,
                   use_tanh=True, use_orthogonal_init=True, chkpt_dir="/tmp")
agent = P.PPO_continuous(a, "pursuer")
s = torch.randn(B, 18, device="cuda"); act = torch.randn(B, 3, device="cuda").clamp(-1.6, 1.6); lp = torch.randn(B, 3, device="cuda") * 0.1 - 1.0
adv = torch.randn(B, 1, device="cuda"); vt = torch.randn(B, 1, device="cuda")
agent.optimize(s, act, lp, adv, vt, mini_batch_size=mb, fused=True)
torch.cuda.synchronize()
print("ok")
