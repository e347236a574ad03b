#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counters to CUDA source lines.

usage: ncu_lines.py <ncu --page source --csv file> <nvdisasm -g -c .sass> <kernel name substring> [top N]
The ncu CSV gives counters per SASS address of one kernel; nvdisasm -g gives the
"//## File ..., line N" markers.  Instructions are matched by order within the kernel.
"""
import csv
import re
import sys

src_csv, sass_file, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = {h: i for i, h in enumerate(rows[hdr])}
insts = rows[hdr + 1:]
# nvdisasm listing of the same kernel
lines = open(sass_file).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
cur = None
line_of = []
inl = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        if line_of:
            break
    m = inl.search(l)
    if m:
        cur = int(m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        line_of.append(cur)
n = min(len(line_of), len(insts))
agg = {}
samples_total = 0
for i in range(n):
    r = insts[i]
    ln = line_of[i]
    a = agg.setdefault(ln, [0, 0, 0])
    a[0] += int(float(r[cols["Instructions Executed"]] or 0))
    a[1] += int(float(r[cols["Thread Instructions Executed"]] or 0))
    a[2] += int(float(r[cols["# Samples"]] or 0))
tot = [sum(v[k] for v in agg.values()) for k in range(3)]
print(f"# {len(insts)} SASS instructions in ncu, {len(line_of)} in nvdisasm; totals: warp-inst {tot[0]}, thread-inst {tot[1]}, samples {tot[2]}")
src = open("/root/repo/variantcalling_b200/csrc/kernels.cu").read().split("\n")
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    text = src[ln - 1].strip()[:90] if ln and ln <= len(src) else "?"
    print(f"{ln!s:>5} inst {100*v[0]/max(1,tot[0]):5.1f}%  samples {100*v[2]/max(1,tot[2]):5.1f}%  | {text}")
