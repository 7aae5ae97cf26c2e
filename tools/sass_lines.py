"""Per-source-line instruction counts of one kernel: joins `ncu --page source --csv` (SASS view, per-instruction executed
counts) with `nvdisasm -g -c` line info of the same cubin.
usage: sass_lines.py <ncu_source.csv> <nvdisasm.sass> <mangled-name-substring> [warps]"""
import re, csv, sys, collections
csvp, sassp, key = sys.argv[1:4]
warps = float(sys.argv[4]) if len(sys.argv) > 4 else 4096.0
lines = open(sassp).read().splitlines()
heads = [i for i, l in enumerate(lines) if l.startswith('.text.')]
start = [i for i in heads if key in lines[i]][0]
end = min([i for i in heads if i > start] + [len(lines)])
cur, instr = None, []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m2 = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m2: instr.append((cur, m2.group(2)))
rows = list(csv.reader(open(csvp)))
h = rows[1]; si = h.index('# Samples'); ii = h.index('Instructions Executed'); ti = h.index('Thread Instructions Executed')
data = []
for r in rows[2:]:
    try: data.append((int(r[si]), int(r[ii]), r[1], int(r[ti])))
    except Exception: pass
if len(data) > len(instr) and len(data) % len(instr) == 0:
    data = data[:len(instr)]          # the csv lists every matching launch back to back: keep the first
assert len(instr) == len(data), (len(instr), len(data))
agg = collections.defaultdict(lambda: [0, 0, 0])
for (c, _), d in zip(instr, data):
    k = c if c else ('?', 0)
    agg[k][0] += d[0]; agg[k][1] += d[1]; agg[k][2] += 1
tot_s = sum(v[0] for v in agg.values()); tot_i = sum(v[1] for v in agg.values())
print(f"samples {tot_s}, warp-instructions {tot_i} = {tot_i / warps:.0f} per warp")
import os
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ppo-rl-satellite_b200", "csrc")
src = {f: open(os.path.join(root, f)).read().splitlines() for f in os.listdir(root) if f.endswith((".cu", ".cuh"))}
print("CALL sites:")
for (c, txt), d in zip(instr, data):
    if 'CALL' in txt and d[1] / warps > 0.05:
        t = src.get(c[0], [""] * 100000)[c[1] - 1].strip()[:70] if c else ""
        print(f"  {d[1] / warps:6.2f}/warp  lanes {d[3] / max(1, d[1]):5.1f}  {c[0]}:{c[1]}  | {t}")
print("top lines:")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("TOP", "40"))]:
    t = src[f][l - 1].strip()[:86] if f in src and l - 1 < len(src[f]) else ''
    print(f"{f[:13]:13s}:{l:4d} {100 * v[1] / tot_i:5.2f}% ({v[1] / warps:7.1f}/warp {v[2]:4d} sass) smp {100 * v[0] / tot_s:5.2f}%  {t}")
