"""Importable alias of the package directory `ppo-rl-satellite_b200/` (a hyphen is not a valid
Python identifier, so `import ppo_rl_satellite_b200` resolves its submodules from that directory).

    from ppo_rl_satellite_b200 import engine            # batched CUDA engine
    from ppo_rl_satellite_b200.dropin import environment # reference-named drop-in modules
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ppo-rl-satellite_b200")]
__version__ = "0.1.0"
