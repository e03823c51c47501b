#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass` export by CUDA source line.

The sass CSV carries no line numbers, so they are taken from `nvdisasm -g -c` of the cubin the
capture ran (same build!) and matched to the CSV rows by instruction order.

usage: ncu_by_line.py sass.csv disasm.txt mangled_name_substring [top_n]
  disasm.txt = nvdisasm -g -c <cubin>   (cubin from: cuobjdump -xelf all libb2j.so)
"""
import csv
import re
import sys

csv_path, dis_path, func = sys.argv[1], sys.argv[2], sys.argv[3]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# ---- nvdisasm: instruction order -> (file, line) for the chosen function
lines = open(dis_path, errors='replace').read().split('\n')
start = None
for i, l in enumerate(lines):
    if l.startswith('.text.') and func in l and l.rstrip().endswith(':'):
        start = i
        break
if start is None:
    sys.exit('function not found in disassembly: ' + func)
marker = re.compile(r'//## File "([^"]+)", line (\d+)')
instr = re.compile(r'^\s+/\*[0-9a-f]{4,}\*/\s+(.*?);')
where = []
cur = ('?', 0)
for l in lines[start + 1:]:
    if l.startswith('//-----') or (l.startswith('.text.') and l.rstrip().endswith(':')):
        break
    m = marker.search(l)
    if m:
        if 'inlined at' not in l:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = instr.match(l)
    if m:
        where.append((cur, m.group(1)))

# ---- ncu sass csv
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ia, isamp, iex, ith = (hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed'),
                       hdr.index('Thread Instructions Executed'))
data = [r for r in rows[2:] if len(r) > ith and r[ia].startswith('0x')]
if len(data) != len(where):
    sys.exit('instruction count differs: csv %d vs disassembly %d (capture and cubin are different builds?)'
             % (len(data), len(where)))

agg = {}
tot_s = tot_e = tot_t = 0
for r, (loc, _) in zip(data, where):
    s, e, t = int(r[isamp] or 0), int(r[iex] or 0), int(r[ith] or 0)
    a = agg.setdefault(loc, [0, 0, 0, 0])
    a[0] += s
    a[1] += e
    a[2] += t
    a[3] += 1
    tot_s += s
    tot_e += e
    tot_t += t
print('instructions %d  samples %d  warp-instr executed %d  avg active threads %.1f'
      % (len(data), tot_s, tot_e, tot_t / max(tot_e, 1)))
src_cache = {}


def text(loc):
    f, n = loc
    if f not in src_cache:
        try:
            import glob
            p = glob.glob('/root/repo/**/' + f, recursive=True)[0]
            src_cache[f] = open(p).read().split('\n')
        except Exception:
            src_cache[f] = []
    s = src_cache[f]
    return s[n - 1].strip()[:90] if 0 < n <= len(s) else ''


print('%-28s %6s %7s %7s %6s %5s  %s' % ('file:line', 'sass', 'exec%', 'samp%', 'thr', 's/e', 'source'))
for loc, (s, e, t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print('%-28s %6d %6.2f%% %6.2f%% %6.1f %5.2f  %s' % ('%s:%d' % loc, n, 100 * e / max(tot_e, 1), 100 * s / max(tot_s, 1),
                                                          t / max(e, 1), s / max(e, 1) * tot_e / max(tot_s, 1), text(loc)))
