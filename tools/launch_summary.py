#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): last 1/Nth of the launches, per kernel."""
import collections, csv, re, sys
path, nsteps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = list(csv.reader(open(path)))
for i, r in enumerate(rows):
    if r and r[0] == 'ID':
        hdr, start = r, i + 1
        break
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
L = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[start:] if len(r) > vi]
n = len(L) // nsteps
agg = collections.OrderedDict()
for k, v in L[-n:]:
    k = re.sub(r'\(.*', '', k)[:80]
    agg.setdefault(k, [0, 0]); agg[k][0] += v; agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0]/1e3:9.1f} us  {100*v[0]/tot:5.1f}%  x{v[1]:3d}  {k}")
print(f"{tot/1e3:9.1f} us total, {n} launches per step")
