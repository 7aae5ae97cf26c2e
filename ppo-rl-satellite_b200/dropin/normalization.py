"""normalization -- drop-in for the reference's normalization.py (:7-63) on the CUDA running-statistics kernels.

x may be one sample (shape,) as in the reference -- the update is then the reference's Welford step incl. its
first-sample rule (mean = std = x) -- or a batch [N, shape], merged with Chan's parallel update.
"""
import numpy as np

try:
    from ._boot import engine as _eng
except ImportError:  # imported as a top-level module (dropin/ on sys.path, the CPPO_main.py case)
    from _boot import engine as _eng


class RunningMeanStd:
    def __init__(self, shape):
        self.shape = int(np.prod(shape)) if not np.isscalar(shape) else int(shape)
        self._rs = _eng.RunningStats(self.shape)

    n = property(lambda self: self._rs.n)
    mean = property(lambda self: self._rs.mean.cpu().numpy())
    S = property(lambda self: self._rs.S.cpu().numpy())
    std = property(lambda self: self._rs.std.cpu().numpy())

    def _as_batch(self, x):
        import torch
        a = np.asarray(x, dtype=np.float64)
        one = a.ndim <= 1 and a.size == self.shape
        return torch.as_tensor(np.ascontiguousarray(a.reshape(-1, self.shape)), device="cuda"), one

    def update(self, x):
        xb, _ = self._as_batch(x)
        self._rs.update_normalize(xb, update=True)


class Normalization:
    def __init__(self, shape):
        self.running_ms = RunningMeanStd(shape=shape)

    def __call__(self, x, update=True):
        xb, one = self.running_ms._as_batch(x)
        out = self.running_ms._rs.update_normalize(xb, update=update).cpu().numpy()
        return out[0] if one else out


class RewardScaling:
    def __init__(self, shape, gamma):
        self.shape, self.gamma = shape, gamma
        self.running_ms = RunningMeanStd(shape=self.shape)
        self.R = np.zeros(self.shape)

    def __call__(self, x):
        self.R = self.gamma * self.R + x
        self.running_ms.update(self.R)
        return x / (self.running_ms.std + 1e-8)

    def reset(self):
        self.R = np.zeros(self.shape)
