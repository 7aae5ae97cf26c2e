"""ncu driver: the FP64 / FP32 FMA-chain peak microbenchmarks (roofline denominators reported beside the nominal rates)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng
print("dfma chain", eng.measure_vector_peak("fp64", repeats=2))
print("ffma chain", eng.measure_vector_peak("fp32", repeats=2))
print("ffma2 chain", eng.measure_vector_peak("fp32x2", repeats=2))
