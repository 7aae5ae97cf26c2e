"""Loads the upstream reference (read-only, /root/reference) for golden generation and live
cross-checks. Test-side only. The GPU box has no /root/reference: everything that must run there
uses the committed fixtures in tests/golden/ instead.

Recipe from SURVEY.md Appendix A: stub the three absent packages (gym, matplotlib, plotly) before
import; exec only lines 9-40 of the RK4 script (its module body runs 86 400 steps and reads CSVs
that are not shipped).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

REF = os.environ.get("SAT_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "environment.py"))


def _install_stubs():
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        spaces = types.ModuleType("gym.spaces")

        class Box:
            def __init__(self, low=None, high=None, shape=None, dtype=None):
                self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        class Discrete:
            def __init__(self, n):
                self.n = n
                self.shape = ()

        spaces.Box, spaces.Discrete = Box, Discrete
        gym.spaces = spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    for name in ("matplotlib", "matplotlib.pyplot", "plotly", "plotly.graph_objects", "mpl_toolkits",
                 "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            def _ga(attr, _n=name):  # any public attribute is a no-op callable
                if attr.startswith("__"):
                    raise AttributeError(attr)
                return lambda *a, **k: None
            m.__getattr__ = _ga
            m.__file__ = "<stub %s>" % name
            m.rcParams = {}
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["plotly"].graph_objects = sys.modules["plotly.graph_objects"]


_mods = {}


def load():
    """returns dict(environment=..., satellite_function=..., ppo_continuous=..., replaybuffer=...,
    normalization=..., CPPO_main=...) of the *reference's own* modules."""
    if _mods:
        return _mods
    assert available(), "reference tree not present"
    _install_stubs()
    saved = {k: sys.modules.get(k) for k in ("environment", "satellite_function", "ppo_continuous",
                                             "replaybuffer", "normalization", "CPPO_main", "plot_function",
                                             "single_pluse_model")}
    for k in saved:
        sys.modules.pop(k, None)
    sys.path.insert(0, REF)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import satellite_function, environment, ppo_continuous, replaybuffer, normalization, CPPO_main  # noqa
        _mods.update(environment=environment, satellite_function=satellite_function,
                     ppo_continuous=ppo_continuous, replaybuffer=replaybuffer,
                     normalization=normalization, CPPO_main=CPPO_main)
    finally:
        sys.path.remove(REF)
        # leave the reference modules registered under private names only, so the product's
        # same-named drop-in modules can be imported in the same process
        for k in saved:
            m = sys.modules.pop(k, None)
            if m is not None:
                sys.modules["_ref_" + k] = m
            if saved[k] is not None:
                sys.modules[k] = saved[k]
    return _mods


def load_rk4_script(j2=None):
    """exec lines 9-40 of the RK4 script; returns its namespace (StateEq, RungeKutta, mu, Re, J2)."""
    path = os.path.join(REF, "轨道外推-龙格库塔算法.py")
    src = open(path, encoding="utf-8").read().splitlines()
    ns = {"np": np}
    exec("\n".join(src[8:40]), ns)
    if j2 is not None:
        ns["J2"] = j2
    return ns


class Args:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def make_env(d_capture=20000, max_episode_steps=1000):
    env_mod = load()["environment"]
    with contextlib.redirect_stdout(io.StringIO()):
        env = env_mod.satellites(d_capture=d_capture, args=Args(max_episode_steps=max_episode_steps))
    env.d_capture = d_capture
    return env


def quiet_step(env, pa, ea, count):
    with contextlib.redirect_stdout(io.StringIO()):
        return env.step(pa, ea, count)
