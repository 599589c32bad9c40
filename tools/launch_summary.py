"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
per = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(row['Metric Unit'], 1.0)
    per[row['Kernel Name'].split('(')[0].replace('an3d::', '').replace('<unnamed>::', '')].append(v)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tot = sum(sum(v) for v in per.values())
print(f"{len(per)} kernels, {sum(len(v) for v in per.values())} launches, {tot / steps:.1f} us per step (serialised, cold cache)")
for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:52]:52s} n={len(v)//steps:3d} sum={sum(v)/steps:8.1f} us {100*sum(v)/tot:5.1f}%  first: {' '.join(f'{x:.0f}' for x in v[:6])}")
