#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python profiles/launch_summary.py <csv>"""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = OrderedDict(); seq = []
for r in rows[1:]:
    name = r[ki].split('(')[0][:90]; v = float(r[vi].replace(',', '')); u = r[ui]
    v = v * 1000 if u == 'ms' else (v / 1000 if u == 'ns' else v)
    seq.append((name, v)); a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in seq)
print('launches', len(seq), 'total_us', round(tot))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%6d x %10.1f us avg  %5.1f%%  %s' % (c, t / c, 100 * t / tot, k))
