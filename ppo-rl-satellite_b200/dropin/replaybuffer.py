"""replaybuffer.ReplayBuffer -- same interface as the reference (replaybuffer.py:3-38): host numpy storage for the
single-env driver loop of CPPO_main.py. The batched, device-resident rollout storage is
ppo_rl_satellite_b200.rollout.RolloutBuffer."""
import numpy as np
import torch


class ReplayBuffer:
    def __init__(self, args):
        self.state_dim, self.action_dim, self.batch_size = args.state_dim, args.action_dim, args.batch_size
        self.s = np.zeros((args.batch_size, args.state_dim))
        self.a = np.zeros((args.batch_size, args.action_dim))
        self.a_logprob = np.zeros((args.batch_size, args.action_dim))
        self.r = np.zeros((args.batch_size, 1))
        self.s_ = np.zeros((args.batch_size, args.state_dim))
        self.dw = np.zeros((args.batch_size, 1))
        self.done = np.zeros((args.batch_size, 1))
        self.count = 0

    def store(self, s, a, a_logprob, r, s_, dw, done):
        index = self.count % self.batch_size
        self.s[index], self.a[index], self.a_logprob[index] = s, a, a_logprob
        self.r[index], self.s_[index], self.dw[index], self.done[index] = r, s_, dw, done
        self.count += 1

    def numpy_to_tensor(self, device=None):
        """7 float32 tensors (s, a, a_logprob, r, s_, dw, done); on `device` when given (the reference returns CPU)."""
        out = tuple(torch.tensor(x, dtype=torch.float) for x in
                    (self.s, self.a, self.a_logprob, self.r, self.s_, self.dw, self.done))
        return tuple(t.to(device) for t in out) if device is not None else out
