"""Makes `ppo_rl_satellite_b200` importable when only the dropin/ directory was put on sys.path."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from ppo_rl_satellite_b200 import engine, _lib  # noqa: E402,F401
