#!/usr/bin/env python
"""Compact role-level view of an ncu source page of the fused MLP kernel: mbarrier waits (spin counts) and MMA issue."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
num = lambda x: int(float(x)) if x not in ('', None) else 0
tot = sum(num(r[ix['# Samples']]) for r in data)
print('total samples', tot)
mma = [i for i, r in enumerate(data) if 'UTCHMMA' in r[ix['Source']]]
waits = [i for i, r in enumerate(data) if 'TRYWAIT' in r[ix['Source']]]
print('UTCHMMA count', len(mma), 'first/last idx', mma[0], mma[-1])
lo, hi2 = mma[0] - 120, mma[-1] + 40
s_mma = sum(num(r[ix['# Samples']]) for r in data[lo:hi2])
print('samples in MMA-issuer region', s_mma, '(%.1f%% of one warp share %.0f)' % (100.0 * s_mma / (tot / 20), tot / 20))
agg = {}
for i in waits:
    r = data[i]
    key = r[ix['Source']].split('TRYWAIT')[1].strip()[:40]
    spin = sum(num(q[ix['# Samples']]) for q in data[i - 1:i + 6])
    region = 'mma' if lo <= i <= hi2 else 'other'
    a = agg.setdefault((region, key), [0, 0, 0]); a[0] += 1; a[1] += num(r[ix['Instructions Executed']]); a[2] += spin
for (region, key), (n, ex, sp) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:14]:
    print(f'{region:6s} {key:42s} sites {n:4d} executed {ex:12d} samples {sp:8d}')
sm = sum(num(data[i][ix['# Samples']]) for i in mma)
print('samples on UTCHMMA instrs', sm, ' on UTCBAR', sum(num(r[ix['# Samples']]) for r in data if 'UTCBAR' in r[ix['Source']]))
