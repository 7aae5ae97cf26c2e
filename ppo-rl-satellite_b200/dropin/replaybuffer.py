"""Transition store for the single-env driver loop of CPPO_main.py (interface of the reference's replaybuffer.py:3-38:
`ReplayBuffer(args)`, `.store(...)`, `.count`, `.numpy_to_tensor()`).

Layout: ONE float64 record block [batch_size, 2*state_dim + 2*action_dim + 3]; the seven fields the reference keeps as
separate arrays are column views of it (`.s`, `.a`, `.a_logprob`, `.r`, `.s_`, `.dw`, `.done`), so a transition is one row
write and the whole batch moves to the GPU as one contiguous transfer. The batched, device-resident rollout storage is
ppo_rl_satellite_b200.rollout.RolloutBuffer."""
import numpy as np
import torch

_FIELDS = ("s", "a", "a_logprob", "r", "s_", "dw", "done")


class ReplayBuffer:
    def __init__(self, args):
        sd, ad = int(args.state_dim), int(args.action_dim)
        self.batch_size, self.state_dim, self.action_dim = int(args.batch_size), sd, ad
        widths = dict(s=sd, a=ad, a_logprob=ad, r=1, s_=sd, dw=1, done=1)
        self._cols, lo = {}, 0
        for name in _FIELDS:
            self._cols[name] = slice(lo, lo + widths[name])
            lo += widths[name]
        self._rec = np.zeros((self.batch_size, lo), dtype=np.float64)
        self.count = 0                                   # the driver resets it after every update (CPPO_main.py:147)

    def __getattr__(self, name):                         # field views: buf.s, buf.a, ... like the reference's arrays
        cols = self.__dict__.get("_cols")
        if cols is not None and name in cols:
            return self._rec[:, cols[name]]
        raise AttributeError(name)

    def store(self, s, a, a_logprob, r, s_, dw, done):
        row = self._rec[self.count % self.batch_size]
        for name, value in zip(_FIELDS, (s, a, a_logprob, r, s_, dw, done)):
            row[self._cols[name]] = value
        self.count += 1

    def numpy_to_tensor(self, device=None):
        """(s, a, a_logprob, r, s_, dw, done) as float32 tensors of shape [batch_size, width]; CPU like the reference,
        or on `device` (one transfer of the record block, then one contiguous tensor per field)."""
        block = torch.from_numpy(self._rec).to(dtype=torch.float32)
        if device is not None:
            block = block.to(device)
        return tuple(block[:, self._cols[name]].contiguous() for name in _FIELDS)
