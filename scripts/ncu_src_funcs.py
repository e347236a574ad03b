#!/usr/bin/env python
"""Per-function totals (instructions, active lanes, stall samples) of one kernel from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`; inlined helpers are attributed to themselves.
The CSV has one section per source file: every line is attributed to the function that spans it IN ITS OWN FILE.
usage: ncu_src_funcs.py <csv> [top N lines]"""
import collections
import csv
import os
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
starts_cache = {}


def starts_of(path):
    if path not in starts_cache:
        try:
            src = open(path).read().split("\n")
        except OSError:
            src = []
        starts_cache[path] = [(i + 1, l) for i, l in enumerate(src) if re.match(r"^(__device__|__global__|template|static)", l)]
    return starts_cache[path]


def fn_of(path, ln):
    name = os.path.basename(path)
    for s, l in starts_of(path):
        if s <= ln:
            m = re.search(r"(\w+)\(", l)
            name = m.group(1) if m else l[:30]
    return name


agg = collections.defaultdict(lambda: [0, 0, 0])
lines = collections.defaultdict(lambda: [0, 0, 0, ""])
path, cols = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1]
        continue
    if r[0] == "Line No":
        cols = {}
        for i, c in enumerate(r):
            cols.setdefault(c, i)
        continue
    if cols is None or not r[0].strip().isdigit() or len(r) < len(cols) - 5:
        continue
    try:
        vals = [float(r[cols[c]] or 0) for c in ("Instructions Executed", "Thread Instructions Executed", "# Samples")]
    except ValueError:
        continue
    a = agg[fn_of(path, int(r[0]))]
    b = lines[(os.path.basename(path), int(r[0]))]
    for i in range(3):
        a[i] += vals[i]
        b[i] += vals[i]
    b[3] = r[1]
tot = sum(v[0] for v in agg.values())
ts = sum(v[2] for v in agg.values())
print(f"# warp-inst {tot:.0f}, samples {ts:.0f}")
for f, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] / tot > 0.003:
        print(f"{f:22s} warp-inst {100 * v[0] / tot:5.1f}%  lanes {v[1] / max(1, v[0]):5.1f}  samples {100 * v[2] / ts:5.1f}%")
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:<5d} warp-inst {100 * v[0] / tot:5.1f}%  lanes {v[1] / max(1, v[0]):5.1f}  samples {100 * v[2] / ts:5.1f}% | {v[3].strip()[:90]}")
