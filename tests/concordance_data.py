"""Synthetic concordance frames (the columns evaluate_concordance reads from the comparison HDF5:
``filter``, ``tree_score``, ``classify`` / ``classify_gt``, ``indel``, ``hmer_indel_length``,
``hmer_indel_nuc``), deterministic, shared by the golden generator and the tests."""
from __future__ import annotations

import numpy as np
import pandas as pd


def make_frame(n: int, seed: int, *, p_fn=0.03, score_decimals=None, with_tn=False, nan_scores=0) -> pd.DataFrame:
    rng = np.random.default_rng(seed)
    indel = rng.random(n) < 0.25
    hmer = np.where(indel & (rng.random(n) < 0.7), rng.integers(1, 16, size=n), 0)
    hmer = np.where(~indel & (rng.random(n) < 0.01), rng.integers(1, 5, size=n), hmer)  # hmer length on a non-indel row
    truth = rng.random(n) < 0.8
    # scores: float32 values (TREE_SCORE read back from a VCF), higher for true calls
    score = np.clip(rng.normal(np.where(truth, 60, 25), 18), 0, None).astype(np.float32).astype(np.float64)
    if score_decimals is not None:
        score = np.round(score, score_decimals)  # many ties
    classify = np.where(truth, "tp", "fp").astype(object)
    fn = rng.random(n) < p_fn
    classify[fn] = "fn"
    if with_tn:
        classify[(rng.random(n) < 0.02) & ~fn] = "tn"
    low = score <= 30.0  # noqa: PLR2004
    other = rng.random(n) < 0.04
    filt = np.where(low & other, "LOW_SCORE;HPOL_RUN", np.where(low, "LOW_SCORE", np.where(other, "HPOL_RUN", "PASS"))).astype(object)
    filt[rng.random(n) < 0.01] = "PASS;PASS"
    score[fn] = np.nan if nan_scores else score[fn]
    if nan_scores:
        score[rng.integers(0, n, size=nan_scores)] = np.nan
    # genotype-aware classification differs on a few rows
    classify_gt = classify.copy()
    flip = (rng.random(n) < 0.03) & (classify == "tp")
    classify_gt[flip] = "fp"
    return pd.DataFrame({"filter": filt, "tree_score": score, "classify": classify, "classify_gt": classify_gt,
                         "indel": indel, "hmer_indel_length": hmer.astype(np.int64),
                         "hmer_indel_nuc": np.where(hmer > 0, "A", None)})


def make_cases():
    """(name, frame, classify column, group column or None)"""
    yield "mixed_6k", make_frame(6000, 1), "classify", None
    yield "mixed_gt_ties", make_frame(9000, 2, score_decimals=0), "classify_gt", None
    yield "tiny_groups", make_frame(120, 3, p_fn=0.1), "classify", None           # groups below the 20-call cutoff
    yield "with_tn_and_nan", make_frame(2500, 4, with_tn=True, nan_scores=25), "classify", None
    f = make_frame(2000, 5)
    f.loc[f["indel"], "classify"] = "fn"                                          # every indel group: only false negatives
    yield "all_fn_indels", f, "classify", None
    f = make_frame(2000, 6)
    f = f[~f["indel"]].reset_index(drop=True)                                     # empty indel groups
    yield "snps_only", f, "classify", None
    f = make_frame(3000, 7)
    f["my_group"] = np.where(f["indel"], "Non-hmer INDEL", "SNP")
    f.loc[::17, "my_group"] = None                                                # rows without a type are dropped
    yield "custom_group_column", f, "classify", "my_group"
