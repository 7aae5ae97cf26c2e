"""Per-PHASE instruction counts of one kernel: like sass_lines.py, but every SASS instruction is attributed to the
outermost source line of its inline chain (nvdisasm -gi), i.e. to the statement of the kernel body it belongs to.
usage: sass_phases.py <ncu_source.csv> <nvdisasm -gi -c output> <mangled-name-substring> [warps]"""
import re, csv, sys, collections
csvp, sassp, key = sys.argv[1:4]
warps = float(sys.argv[4]) if len(sys.argv) > 4 else 4096.0
lines = open(sassp).read().splitlines()
heads = [i for i, l in enumerate(lines) if l.startswith('.text.')]
start = [i for i in heads if key in lines[i]][0]
end = min([i for i in heads if i > start] + [len(lines)])
chain, instr, pending = [], [], []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        pending.append((m.group(1).split('/')[-1], int(m.group(2)), m.group(3).split('/')[-1] if m.group(3) else None, int(m.group(4)) if m.group(4) else None))
        continue
    m2 = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m2:
        if pending:
            chain = pending; pending = []
        inner = (chain[0][0], chain[0][1]) if chain else ('?', 0)
        outer = inner
        for c in chain:
            outer = (c[2], c[3]) if c[2] else (c[0], c[1])
        instr.append((inner, outer, m2.group(2)))
rows = list(csv.reader(open(csvp)))
h = rows[1]; si = h.index('# Samples'); ii = h.index('Instructions Executed'); ti = h.index('Thread Instructions Executed')
data = []
for r in rows[2:]:
    try: data.append((int(r[si]), int(r[ii]), int(r[ti])))
    except Exception: pass
data = data[:len(instr)]
assert len(instr) == len(data), (len(instr), len(data))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for (inner, outer, txt), d in zip(instr, data):
    a = agg[outer]; a[0] += d[0]; a[1] += d[1]; a[2] += 1; a[3] += d[2]
tot_s = sum(v[0] for v in agg.values()); tot_i = sum(v[1] for v in agg.values())
print(f"samples {tot_s}, warp-instructions {tot_i} = {tot_i / warps:.0f} per warp, static SASS {len(instr)}")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{str(f)[:14]:14s}:{l:4d} {100 * v[1] / tot_i:5.2f}% ({v[1] / warps:7.1f}/warp, {v[2]:5d} sass, lanes {v[3] / max(1, v[1]):4.1f}) samples {100 * v[0] / tot_s:5.2f}%")
