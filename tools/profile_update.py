"""few eager PPO optimiser steps at minibatch 65536 for an ncu launch list (where does the update time go?)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ppo_rl_satellite_b200.dropin import ppo_continuous as P
a = bench._PpoArgs(policy_dist="Gaussian", max_action=1.6, batch_size=1 << 18, mini_batch_size=65536, max_train_steps=int(3e6),
                   lr_a=2e-4, lr_c=2e-4, gamma=0.99, lamda=0.95, epsilon=0.1, K_epochs=1, entropy_coef=0.01, set_adam_eps=True,
                   use_grad_clip=True, use_lr_decay=True, use_adv_norm=True, state_dim=18, action_dim=3, hidden_width=256,
                   use_tanh=True, use_orthogonal_init=True, chkpt_dir="/tmp")
agent = P.PPO_continuous(a, "pursuer")
B = 1 << 18
s = torch.randn(B, 18, device="cuda"); act = torch.randn(B, 3, device="cuda").clamp(-1.6, 1.6); lp = torch.randn(B, 3, device="cuda")
adv = torch.randn(B, 1, device="cuda"); vt = torch.randn(B, 1, device="cuda")
agent.optimize(s, act, lp, adv, vt, mini_batch_size=65536)      # 4 steps warm-up
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); agent.optimize(s, act, lp, adv, vt, mini_batch_size=65536); e1.record(); e1.synchronize()
print("eager: %.2f ms per optimiser step" % (e0.elapsed_time(e1) / 4))
