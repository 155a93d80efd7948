"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name, share of the total."""
import collections, csv, io, sys
path, div = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = [l for l in open(path) if not l.startswith('==')]
r = list(csv.DictReader(io.StringIO(''.join(rows))))
d = collections.defaultdict(lambda: [0, 0.0])
for x in r:
    try:
        v = float(x['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    k = x['Kernel Name'][:90]
    d[k][0] += 1
    d[k][1] += v * {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(x['Metric Unit'], 1)
tot = sum(v[1] for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]:
    print(f"{v[1] / div:10.1f} us {v[0]:5d} launches {v[1] / tot * 100:5.1f}%  {k}")
print(f"total {tot / div:.1f} us over {len(r)} launches (divided by {div})")
