"""CPU tests of host-side pieces that need neither the GPU nor the CUDA library."""
import importlib.util
import os
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "ppo-rl-satellite_b200", "dropin", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_replay_buffer_contract():
    """replaybuffer.py:3-38: store() fills row count % batch_size, the driver resets count, numpy_to_tensor() returns the
    seven float32 tensors (s, a, a_logprob, r, s_, dw, done) with shapes [B, width]"""
    rb = _load("replaybuffer")
    args = types.SimpleNamespace(state_dim=18, action_dim=3, batch_size=4)
    buf = rb.ReplayBuffer(args)
    rng = np.random.default_rng(0)
    rows = []
    for i in range(6):                                       # two more than the batch: rows 0 and 1 are overwritten
        row = (rng.normal(size=18), rng.normal(size=3), rng.normal(size=3), float(rng.normal()), rng.normal(size=18),
               bool(i % 2), bool(i % 3 == 0))
        buf.store(*row)
        rows.append(row)
    assert buf.count == 6
    t = buf.numpy_to_tensor()
    assert [tuple(x.shape) for x in t] == [(4, 18), (4, 3), (4, 3), (4, 1), (4, 18), (4, 1), (4, 1)]
    assert all(x.dtype == torch.float32 and x.is_contiguous() for x in t)
    for slot, src in ((0, 4), (1, 5), (2, 2), (3, 3)):
        for field, value in zip(t, rows[src]):
            np.testing.assert_allclose(field[slot].numpy().ravel(), np.asarray(value, dtype=np.float64).ravel().astype(np.float32))
    assert buf.s.shape == (4, 18) and buf.done.shape == (4, 1) and buf.a_logprob.shape == (4, 3)     # field views
    buf.count = 0                                            # what CPPO_main.py:147 does after an update
    buf.store(*rows[0])
    np.testing.assert_array_equal(buf.s[0], rows[0][0])
