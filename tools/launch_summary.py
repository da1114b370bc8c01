#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('<unnamed>::', '')
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:60]:60s} {c:5d} {t:10.1f} {t / c:9.1f} {t / tot * 100:5.1f}%")
