#!/usr/bin/env python
"""Per-kernel mean duration and share from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0][:60]
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print('%-62s n=%3d  mean %9.1f us  share %5.1f%%' % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
