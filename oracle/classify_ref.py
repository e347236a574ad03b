"""ORACLE (test infrastructure, not product code).

The per-record classification rules of ``vcf2concordance`` restated (ugbio_comparison/comparison_utils.py:153-229):
``classify`` (:153-182, allele match), ``classify_gt`` (:186-213, allele + genotype match) and the frame-level
fix-ups (:214-229): a genotype-level tp whose allele-level call is fp becomes fp; records vcfeval marked FN / FN_CA
in BASE turn fp into fn in both columns.  Pinned cell by cell by tests/golden/classify_rules.json, which
scripts/make_golden_classify.py produces from the reference's own function source."""
from __future__ import annotations


def _is_none(gt) -> bool:
    return tuple(gt) in ((None, None), (None,))


def classify(gt_ultima, gt_ground_truth) -> str:
    if _is_none(gt_ultima):
        return "fn"
    if _is_none(gt_ground_truth):
        return "fp"
    set_gtr = set(gt_ground_truth) - {0}
    set_ultima = set(gt_ultima) - {0}
    if set_gtr & set_ultima:
        return "tp"
    if set_ultima - set_gtr:
        return "fp"
    return "fn"


def classify_gt(gt_ultima, gt_ground_truth) -> str:
    n_ref_gtr = sum(1 for y in gt_ground_truth if y == 0)
    n_ref_ultima = sum(1 for y in gt_ultima if y == 0)
    if _is_none(gt_ultima):
        return "fn"
    if _is_none(gt_ground_truth):
        return "fp"
    if n_ref_gtr < n_ref_ultima:
        return "fn"
    if n_ref_gtr > n_ref_ultima:
        return "fp"
    if tuple(gt_ultima) != tuple(gt_ground_truth):
        return "fp"
    return "tp"


def classify_records(gt_ultima: list, gt_ground_truth: list, base: list) -> tuple[list[str], list[str]]:
    cls, cls_gt = [], []
    for gu, gt, b in zip(gt_ultima, gt_ground_truth, base):
        c, g = classify(gu, gt), classify_gt(gu, gt)
        if g == "tp" and c == "fp":
            g = "fp"
        if b in ("FN", "FN_CA"):
            if c == "fp":
                c = "fn"
            if g == "fp":
                g = "fn"
        cls.append(c)
        cls_gt.append(g)
    return cls, cls_gt
