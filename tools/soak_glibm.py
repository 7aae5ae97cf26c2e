"""Long bit-identity soak of csrc/glibm.cuh against the system libm (host build of the device header).
    python tools/soak_glibm.py [samples_per_distribution=30000000] [seeds=4]
Writes nothing; prints the per-function totals. ~1e9 samples per seed at the default size."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = sys.argv[1] if len(sys.argv) > 1 else "30000000"
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
exe = os.path.join(tempfile.mkdtemp(), "glibm_host")
subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-I", os.path.join(ROOT, "ppo-rl-satellite_b200", "csrc"),
                os.path.join(ROOT, "tests", "glibm_host.cpp"), "-o", exe, "-lm"], check=True)
procs = [subprocess.Popen([exe, n, str(100 + s)], stdout=subprocess.PIPE, text=True) for s in range(seeds)]
tot, bad = {}, {}
for p in procs:
    for ln in p.communicate()[0].splitlines():
        f = ln.split()
        if len(f) == 4 and f[2].isdigit():
            tot[f[0]] = tot.get(f[0], 0) + int(f[2]); bad[f[0]] = bad.get(f[0], 0) + int(f[3])
        elif "MISMATCH" in ln:
            print(ln)
for k in tot:
    print(f"{k:10s} samples {tot[k]:>14,d}  mismatches vs libm {bad[k]}")
print("total", sum(tot.values()), "mismatches", sum(bad.values()))
