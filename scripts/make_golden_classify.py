#!/usr/bin/env python
"""Generate tests/golden/classify_rules.json from the REFERENCE's own per-record classification rules: the nested
functions ``classify`` and ``classify_gt`` of ``vcf2concordance`` (ugbio_comparison/comparison_utils.py:153-213) and
the frame-level fix-ups that follow them (:214-229).  The module itself cannot be imported here (pysam, rtg ...), and
the two functions are closures, so their source is cut out of the reference file with ``ast`` and compiled as it
stands; the fix-ups are replayed with the reference's own pandas statements, copied as data (``FIXUPS``) and
``exec``-ed on the frame.  Input: every combination of called / truth genotype over the alleles {None, 0, 1, 2} in
ploidy 1 and 2, times the vcfeval BASE values."""
import ast
import itertools
import json
import os
import textwrap

import pandas as pd

REF = "/root/reference/ugbio_utils/src/comparison/ugbio_comparison/comparison_utils.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

src = open(REF).read()
tree = ast.parse(src)
outer = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "vcf2concordance")
funcs = {n.name: ast.get_source_segment(src, n) for n in outer.body if isinstance(n, ast.FunctionDef) and n.name in ("classify", "classify_gt")}
assert set(funcs) == {"classify", "classify_gt"}
ns = {"defaultdict": dict, "pd": pd}
for code in funcs.values():
    exec(textwrap.dedent(code), ns)  # noqa: S102  (the reference's own function bodies)

# the statements between the two apply() calls and the index assignment, comparison_utils.py:214-229
lines = src.split("\n")
start = next(i for i, ln in enumerate(lines) if 'concordance_df["classify_gt"] = concordance_df.apply(classify_gt' in ln)
stop = next(i for i, ln in enumerate(lines) if "concordance_df.index = pd.Index" in ln)
FIXUPS = textwrap.dedent("\n".join(lines[start + 1:stop]))

alleles = [None, 0, 1, 2]
gts = [(a,) for a in alleles] + list(itertools.product(alleles, alleles))
rows = [{"gt_ultima": gu, "gt_ground_truth": gt, "base": base} for gu in gts for gt in gts for base in ("TP", "FN", "FN_CA", "IGN", None)]
concordance_df = pd.DataFrame(rows)
concordance_df["classify"] = concordance_df.apply(ns["classify"], axis=1, result_type="reduce")
concordance_df["classify_gt"] = concordance_df.apply(ns["classify_gt"], axis=1, result_type="reduce")
exec(FIXUPS, {"concordance_df": concordance_df, "pd": pd})  # noqa: S102
out = {"generator": "scripts/make_golden_classify.py", "reference": "ugbio_comparison/comparison_utils.py:153-229",
       "rows": [{"gt_ultima": list(r.gt_ultima), "gt_ground_truth": list(r.gt_ground_truth), "base": r.base,
                 "classify": r.classify, "classify_gt": r.classify_gt} for r in concordance_df.itertuples()]}
path = os.path.join(ROOT, "tests", "golden", "classify_rules.json")
json.dump(out, open(path, "w"), separators=(",", ":"))
print(f"{len(rows)} rows -> {path}")
