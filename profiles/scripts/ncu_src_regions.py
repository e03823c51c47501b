#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export by address ranges.
usage: ncu_src_regions.py file.csv name:lo:hi [name:lo:hi ...]   (hex offsets from the kernel start)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = rows[2:]
ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
ins = []
base = None
for r in data:
    if len(r) <= iex or not r[ia].startswith('0x'):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    ins.append((a - base, r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0), r))
tot_s = sum(x[2] for x in ins)
tot_e = sum(x[3] for x in ins)
print('instructions', len(ins), 'samples', tot_s, 'warp-instructions executed', tot_e)
for spec in sys.argv[2:]:
    name, lo, hi = spec.split(':')
    lo, hi = int(lo, 16), int(hi, 16)
    s = sum(x[2] for x in ins if lo <= x[0] < hi)
    e = sum(x[3] for x in ins if lo <= x[0] < hi)
    print('%-24s samples %7d (%4.1f%%)  executed %10d (%4.1f%%)  samples/exec %.3f' % (name, s, 100 * s / tot_s, e, 100 * e / tot_e, s / max(e, 1)))
    agg = {}
    for x in ins:
        if lo <= x[0] < hi:
            for i in stall_cols:
                v = x[4][i]
                if v and v != '0':
                    agg[hdr[i]] = agg.get(hdr[i], 0) + int(v)
    print('     ', ', '.join('%s %d' % kv for kv in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
