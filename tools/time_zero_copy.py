"""How fast do SM stores to pinned host memory go, alone and next to a compute kernel? (design input for the host path)"""
import os, sys, time, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng, _lib as L
n = 65536
env = eng.EnvBatch(n, mode="rk4", substeps=100, h=1.0, d_capture=20000.0, max_episode_steps=1000)
lib = L.load()
host = torch.empty((n, 18), dtype=torch.float32).pin_memory()
dev = torch.empty((n, 18), dtype=torch.float32, device="cuda")
pa = torch.rand((n, 3), device="cuda") * 4 - 2; ea = torch.rand((n, 3), device="cuda") * 4 - 2
s2 = torch.cuda.Stream()
def observe(ptr, stream):
    L.check(lib.sat_env_observe(C.byref(env.st), ptr, None, stream), "observe")
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e6
cur = lambda: torch.cuda.current_stream().cuda_stream
print(f"observe -> device memory : {timeit(lambda: observe(dev.data_ptr(), cur())):.0f} us")
print(f"observe -> pinned host    : {timeit(lambda: observe(host.data_ptr(), cur())):.0f} us  (4.7 MB)")
print(f"env step alone            : {timeit(lambda: env.step(pa, ea)):.0f} us")
def both():
    ev = torch.cuda.Event(); ev.record()
    s2.wait_event(ev)
    observe(host.data_ptr(), s2.cuda_stream)
    env.step(pa, ea)
    torch.cuda.current_stream().wait_stream(s2)
print(f"env step || observe->host : {timeit(both):.0f} us")
def serial():
    env.step(pa, ea); observe(host.data_ptr(), cur())
print(f"env step ; observe->host  : {timeit(serial):.0f} us")
def dma():
    env.step(pa, ea); observe(dev.data_ptr(), cur()); host.copy_(dev, non_blocking=True)
print(f"env step ; observe ; DMA  : {timeit(dma):.0f} us")
