"""Print one replay (the middle one) of a tools/timeline.py dump with short kernel names: start, duration, gap on its stream."""
import re, sys
path = sys.argv[1]
L = [l for l in open(path).read().split('\n') if l and not l.startswith('#')]
rows = []
for l in L:
    m = re.match(r'\s*([\d.]+)\s+([\d.]+)\s+(-?[\d.]+)\s+(\*?)\s*(\d+)\s+(.*)', l)
    rows.append((float(m.group(1)), float(m.group(2)), int(m.group(5)), m.group(6)))
idx = [i for i, r in enumerate(rows) if 'seed_advance' in r[3]]
one = rows[idx[1]:idx[2]] if len(idx) >= 3 else rows
t0 = one[0][0]
print("replay span %.1f us, %d activities" % (one[-1][0] + one[-1][1] - t0, len(one)))
last = {}
for s, d, st, n in one:
    gap = s - last.get(st, t0); last[st] = s + d
    n = re.sub(r'^void ', '', n)
    n = re.sub(r'at::native::', '', n)
    short = re.match(r'[\w:]+(<[^(]{0,40})?', n)
    print(f"{s - t0:8.1f} {d:7.1f} {gap:6.1f} s{st:<4}{(short.group(0) if short else n)[:70]}")
