#!/usr/bin/env python
"""Two bgzip'ed VCFs hold the same text and their .tbi files index the same contigs with ranges that select the same
records: the multi-rank output of the tool against the single-process one.  Prints one JSON line.
usage: check_same_vcf.py a.vcf.gz b.vcf.gz"""
import gzip
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from variantcalling_b200 import bgzf_io  # noqa: E402

a, b = sys.argv[1:3]
same = True
n = 0
with gzip.open(a) as fa, gzip.open(b) as fb:
    while True:
        x, y = fa.read(1 << 24), fb.read(1 << 24)
        if x != y:
            same = False
            break
        if not x:
            break
        n += x.count(b"\n")
ia, ib = bgzf_io.read_tbi(a + ".tbi"), bgzf_io.read_tbi(b + ".tbi")
ranges_ok = list(ia) == list(ib)
if ranges_ok:
    for c in list(ia)[:: max(1, len(ia) // 6)]:
        ranges_ok &= bgzf_io.inflate(a, *ia[c]).tobytes() == bgzf_io.inflate(b, *ib[c]).tobytes()
print(json.dumps({"same_text": same, "lines": n, "same_contigs_and_ranges": bool(ranges_ok), "contigs": len(ia),
                  "bytes": [os.path.getsize(a), os.path.getsize(b)]}))
