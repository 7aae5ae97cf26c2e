# Package body lives here; import it as `ppo_rl_satellite_b200` (see ../ppo_rl_satellite_b200/__init__.py).
