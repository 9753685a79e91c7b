"""Aggregate an ncu --metrics gpu__time_duration.sum CSV launch list by kernel (second half = last step)."""
import csv, collections, re, sys
path = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
rows = rows[int(len(rows) * frac):]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0
for row in rows:
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:100]
    t = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    t = t / 1e3 if u.startswith('n') else (t * 1e3 if u.startswith('m') else t)
    g = row['Grid Size']
    agg[name][0] += 1; agg[name][1] += t; tot += t
print(f"{len(rows)} launches, total {tot:.1f} us (serialised, cold-cache)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1]:10.1f} us {v[0]:5d}x  {100*v[1]/tot:5.1f}%  {k}")
