#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: hottest instructions by samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
def num(x):
    try: return int(float(x))
    except Exception: return 0
tot = sum(num(r[ix['# Samples']]) for r in data)
print('instructions', len(data), 'total samples', tot)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(num(r[ix[h]]) for r in data) for h in stalls}
print('stall totals', sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -num(r[ix['# Samples']]))[:n]:
    s = {h: num(r[ix[h]]) for h in stalls}
    best = sorted(s.items(), key=lambda kv: -kv[1])[:2]
    print(r[ix['Address']][-5:], r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']][:64], best)
