"""Builds libsatb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python ppo-rl-satellite_b200/build.py          # or: __graft_entry__.build()

nvcc cross-compiles without a GPU. The .so is git-ignored but travels to the GPU box with the
repository snapshot. Per-file flags: the env / RK4 / GAE units are compiled with -fmad=false so the
only fused multiply-adds in them are the explicit fma() calls (bit-exact numpy arithmetic, see
DESIGN.md); the actor unit uses the default contraction.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsatb200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
UNITS = {
    "rk4.cu": ["-fmad=false"],
    "env_step.cu": ["-fmad=false"],
    "gae.cu": ["-fmad=false"],
    "cw_ode.cu": ["-fmad=false"],
    "reach.cu": ["-fmad=false"],
    "actor.cu": [],
    "actor_tc.cu": [],
    "ppo_update.cu": [],
    "ppo_fb_tc.cu": [],
    "ppo_wgrad2_tc.cu": [],
    "capi.cu": [],
}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "satb200.h"))
    nvcc = _nvcc()
    jobs = []
    objs = []
    for src, extra in UNITS.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            defs = [f"-D{d}" for d in os.environ.get("SAT_NVCC_DEFINES", "").split() if d]     # tuning experiments only
            cmd = [nvcc] + ARCH + COMMON + extra + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)

    with ThreadPoolExecutor(max_workers=min(5, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
