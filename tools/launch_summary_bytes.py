#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list by kernel:
launches, average duration, average DRAM bytes per launch, achieved DRAM GB/s."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
ii, ki, mi, vi, ui = hdr.index('ID'), hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('<unnamed>::', '')
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    if r[mi].startswith('gpu__time'):
        v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else (v * 1e6 if u == 's' else v))  # -> us
        key = 't'
    else:
        v = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        key = 'b'
    d = per.setdefault((r[ii], name), {'t': 0.0, 'b': 0.0})
    d[key] += v
agg = collections.OrderedDict()
for (_, name), d in per.items():
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += d['t']; a[2] += d['b']
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':64s} {'n':>5s} {'avg_us':>9s} {'avg_dram_MB':>12s} {'dram_GB/s':>10s} {'share':>6s}")
for k, (c, t, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:64]:64s} {c:5d} {t / c:9.2f} {b / c / 1e6:12.3f} {(b / t / 1e3 if t > 0 else 0):10.1f} {t / tot * 100:5.1f}%")
