#!/usr/bin/env python
"""Per-function totals (instructions, active lanes, stall samples) of one kernel from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`; inlined helpers are attributed to themselves.
usage: ncu_src_funcs.py <csv> <source file>"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
cols = {}
for i, c in enumerate(rows[h]):
    cols.setdefault(c, i)
src = open(sys.argv[2]).read().split("\n")
starts = [(i + 1, l) for i, l in enumerate(src) if re.match(r"^(__device__|__global__|template)", l)]


def fn_of(ln):
    name = "?"
    for s, l in starts:
        if s <= ln:
            m = re.search(r"(\w+)\(", l)
            name = m.group(1) if m else l[:30]
    return name


agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[h + 1:]:
    if not r[0].strip().isdigit():
        continue
    try:
        a = agg[fn_of(int(r[0]))]
        a[0] += float(r[cols["Instructions Executed"]] or 0)
        a[1] += float(r[cols["Thread Instructions Executed"]] or 0)
        a[2] += float(r[cols["# Samples"]] or 0)
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values())
ts = sum(v[2] for v in agg.values())
print(f"# warp-inst {tot:.0f}, samples {ts:.0f}")
for f, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] / tot > 0.003:
        print(f"{f:22s} warp-inst {100 * v[0] / tot:5.1f}%  lanes {v[1] / max(1, v[0]):5.1f}  samples {100 * v[2] / ts:5.1f}%")
