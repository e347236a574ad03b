#!/usr/bin/env python
"""Per-device-function totals (non-inlined callees) of one kernel from ncu's SASS source page.
usage: ncu_funcs.py <ncu --page source --csv> <nvdisasm -g -c .sass> <kernel substring> <records>"""
import csv
import re
import sys

src_csv, sass, kname, nrec = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
rows = list(csv.reader(open(src_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
cols = {c: i for i, c in enumerate(rows[h])}
insts = rows[h + 1:]
lines = open(sass).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and kname in l)
func, fn_of = kname + '(body)', []
for l in lines[start + 1:]:
    if (l.startswith('.text.') or l.startswith('.section')) and fn_of:
        break
    m = re.match(r'^(\$?[_A-Za-z0-9$]+):\s*$', l)
    if m and '_Z' in m.group(1) and '$' in m.group(1):
        func = m.group(1).split('$')[-1]
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        fn_of.append(func)
agg = {}
for i in range(min(len(fn_of), len(insts))):
    r = insts[i]
    a = agg.setdefault(fn_of[i], [0, 0, 0, 0])
    a[0] += int(float(r[cols['Instructions Executed']] or 0))
    a[1] += int(float(r[cols['Thread Instructions Executed']] or 0))
    a[2] += int(float(r[cols['# Samples']] or 0))
    a[3] += 1
tot = [sum(v[k] for v in agg.values()) for k in range(3)]
print(f"# thread-instructions per record: {tot[1] / nrec:.0f}; warp-instructions per record: {tot[0] / nrec:.0f}")
for f, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{f[:44]:44s} sass {v[3]:5d}  thread-inst/rec {v[1] / nrec:7.0f} ({100 * v[1] / tot[1]:4.1f}%)  "
          f"samples {100 * v[2] / tot[2]:4.1f}%  simt {v[1] / max(1, v[0]):4.1f}")
