"""Drop-in modules with the reference's own module names (environment, satellite_function, ppo_continuous,
replaybuffer, normalization, orbit_rk4). Put this directory in front of the reference directory on sys.path
and CPPO_main.py drives the CUDA path unchanged."""
