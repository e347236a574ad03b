#!/usr/bin/env python
"""Per-source-line totals of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: ncu_src_lines.py <csv> [top N]   (rows carry the CUDA line each SASS instruction belongs to)"""
import csv
import collections
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
cols = {}
for i, c in enumerate(rows[h]):
    cols.setdefault(c, i)
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) < len(rows[h]) - 5 or not r[0].strip().isdigit():
        continue
    ln = int(r[0])
    a = agg.setdefault(ln, [0, 0, 0, r[1]])
    def num(name):
        try:
            return float(r[cols[name]] or 0)
        except ValueError:
            return 0.0
    a[0] += num("Instructions Executed")
    a[1] += num("Thread Instructions Executed")
    a[2] += num("# Samples")
tot = [sum(v[k] for v in agg.values()) for k in range(3)]
print(f"# totals: warp-inst {tot[0]:.0f}, thread-inst {tot[1]:.0f} (avg lanes {tot[1] / max(1, tot[0]):.1f}), samples {tot[2]:.0f}")
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:5d} warp-inst {100 * v[0] / tot[0]:5.1f}%  lanes {v[1] / max(1, v[0]):5.1f}  samples {100 * v[2] / max(1, tot[2]):5.1f}% | {v[3].strip()[:100]}")
