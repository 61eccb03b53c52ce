"""Aggregate an ncu launch-list CSV (gpu__time_duration.sum) over the LAST full training step."""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
names = [r['Kernel Name'] for r in rows]
packs = [i for i, n in enumerate(names) if 'pack_weights' in n]
s, e = packs[-4], packs[-2]   # two pack launches (forward / data-gradient forms) per step
agg, tot = collections.OrderedDict(), 0.0
for r in rows[s:e]:
    n = re.sub(r'\(.*', '', r['Kernel Name']).replace('cvae::', '').replace('void ', '')
    g = r['Grid Size']
    key = n if 'conv' not in n else f"{n} {g}"
    d = float(r['Metric Value']) / 1000.0
    agg.setdefault(key, [0, 0.0]); agg[key][0] += 1; agg[key][1] += d; tot += d
for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{d:9.1f} us  x{c:2d}  {n[:100]}")
print(f"total {tot:.1f} us over {e - s} launches")
