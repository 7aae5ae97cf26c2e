"""is VectorTrainer.collect() GPU- or host-bound at config-5 shard size? wall time per step for several env counts"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ppo_rl_satellite_b200 import engine as eng, rollout
from ppo_rl_satellite_b200.dropin import ppo_continuous as P
a = bench._PpoArgs(policy_dist="Gaussian", max_action=1.6, batch_size=1, mini_batch_size=65536, max_train_steps=int(3e6),
                   lr_a=2e-4, lr_c=2e-4, gamma=0.99, lamda=0.95, epsilon=0.1, K_epochs=10, entropy_coef=0.01, set_adam_eps=True,
                   use_grad_clip=True, use_lr_decay=True, use_adv_norm=True, state_dim=18, action_dim=3, hidden_width=256,
                   use_tanh=True, use_orthogonal_init=True, chkpt_dir="/tmp")
T = 256
for n in (512, 2048, 8192, 32768):
    env = eng.EnvBatch(n, mode="rk4", substeps=100, h=1.0, d_capture=20000.0, max_episode_steps=1000)
    agent, opp = P.PPO_continuous(a, "pursuer"), P.PPO_continuous(a, "evader")
    tr = rollout.VectorTrainer(env, agent, opp, T)
    tr.collect(); torch.cuda.synchronize()
    t0 = time.perf_counter(); tr.collect(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"envs {n:6d}: host issue {1e6 * (t1 - t0) / T:6.1f} us/step, wall {1e6 * (t2 - t0) / T:6.1f} us/step")
