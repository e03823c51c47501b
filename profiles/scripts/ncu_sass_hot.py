#!/usr/bin/env python
"""Hot spots of an `ncu --page source --csv --print-source sass` export.
usage: ncu_sass_hot.py sass.csv [top_n]   -> totals per stall reason, per opcode, and the top instructions by not-issued samples"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and not h.endswith('(Not Issued)')]
tot = Counter()
by_op_exec, by_op_samp = Counter(), Counter()
wf, wf_ideal = 0, 0
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix['Source']].strip()
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    ex = int(r[ix['Instructions Executed']] or 0)
    sm = int(r[ix['# Samples']] or 0)
    by_op_exec[op] += ex
    by_op_samp[op] += sm
    for c in stall_cols:
        tot[c] += int(r[ix[c]] or 0)
    wf += int(r[ix['L1 Wavefronts Shared']] or 0)
    wf_ideal += int(r[ix['L1 Wavefronts Shared Ideal']] or 0)
    data.append((sm, ex, src, {c: int(r[ix[c]] or 0) for c in stall_cols}))
n_ex = sum(by_op_exec.values())
n_sm = sum(by_op_samp.values())
print('warp-instructions executed %d, samples %d, smem wavefronts %d (ideal %d)' % (n_ex, n_sm, wf, wf_ideal))
print('stall reasons (all samples):', ', '.join('%s %.1f%%' % (k.replace('stall_', ''), 100.0 * v / n_sm) for k, v in tot.most_common(10)))
print('opcode           exec%   samp%')
for op, ex in by_op_exec.most_common(28):
    print('%-16s %5.2f%%  %5.2f%%' % (op, 100.0 * ex / n_ex, 100.0 * by_op_samp[op] / n_sm))
print('top instructions by samples:')
for sm, ex, src, st in sorted(data, key=lambda x: -x[0])[:top_n]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print('%6d %5.2f%%  ex %9d  %-58s %s' % (sm, 100.0 * sm / n_sm, ex, src[:58], ' '.join('%s:%d' % (k.replace('stall_', ''), v) for k, v in top if v)))
