"""CPU: the oracle's restatement of evaluate_concordance's metrics against the reference's own
functions (tests/golden/concordance_metrics.json.gz, scripts/make_golden_concordance.py)."""
import gzip
import json
import os

import numpy as np
import pytest

from oracle import concordance_ref as CR
from tests.concordance_data import make_cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "concordance_metrics.json.gz")


def load_golden():
    with gzip.open(GOLDEN, "rt") as fh:
        return json.load(fh)


def as_float(values):
    return np.array([np.nan if v is None else v for v in values], dtype=np.float64)


def check_accuracy(acc, want):
    assert list(acc["group"]) == want["group"]
    for col in CR.METRIC_COLUMNS:
        # the reference rounds to 5 decimals in its own (pandas < 3) environment; see the generator's note
        np.testing.assert_array_equal(np.round(acc[col].to_numpy(dtype=np.float64), 5), np.round(as_float(want[col]), 5), err_msg=col)


def check_curve(curve, want):
    assert list(curve["group"]) == want["group"]
    for col in ("precision", "recall", "f1"):
        for g, got, exp in zip(want["group"], curve[col], want[col]):
            np.testing.assert_array_equal(np.asarray(got, dtype=np.float64), as_float(exp), err_msg=f"{col} of {g}")
    np.testing.assert_array_equal(np.asarray(list(curve["threshold"]), dtype=np.float64), as_float(want["threshold"]))


@pytest.mark.parametrize("case", list(make_cases()), ids=lambda c: c[0])
def test_oracle_reproduces_the_reference(case):
    name, df, classify_col, group_col = case
    want = load_golden()[name]
    check_accuracy(CR.calc_accuracy_metrics(df, classify_col, group_col), want["accuracy"])
    check_curve(CR.calc_recall_precision_curve(df, classify_col, group_col), want["curve"])


class DeviceModel:
    """NumPy model of what csrc/concordance.cu returns (the ugvc_conc_run / ugvc_conc_curve contract of
    include/ugvc_b200.h), so that the host mirror's arithmetic is covered without a GPU."""

    def run(self, scores, pred, cls, indel, hmer, group=None, want_curves=True):
        scores, pred, cls = np.asarray(scores, np.float64), np.asarray(pred).astype(bool), np.asarray(cls)
        indel, hmer = np.asarray(indel).astype(bool), np.asarray(hmer)
        if group is None:
            gid = np.full(len(cls), -1)
            gid[~indel] = 0
            gid[indel & (hmer == 0)] = 1
            for g, lo, hi in ((2, 1, 4), (3, 5, 7), (4, 8, 10), (5, 11, 12), (6, 13, 1 << 30)):
                gid[indel & (hmer >= lo) & (hmer <= hi)] = g
        else:
            gid = np.asarray(group)
        truth, fnm = (cls == 1) | (cls == 2), cls == 2
        members = [gid == g for g in range(7)] + [indel, hmer > 0]
        counts = np.zeros((9, 6), np.int64)
        self.curves, curve_len, cutoff, selected = {}, np.zeros(9, np.int64), np.zeros(9), np.zeros((9, 2), np.int64)
        for g, m in enumerate(members):
            called = m & ~fnm
            t, p = truth[called], pred[called]
            counts[g] = [(t & p).sum(), (p & ~t).sum(), (~p & t).sum(), t.sum(), called.sum(), (m & fnm).sum()]
            if want_curves and g < 8 and called.any():
                raw_p, raw_r, thr = CR.sklearn_pr_curve(t, scores[called])
                self.curves[g] = (raw_p[:-1], raw_r[:-1], thr)
                curve_len[g] = len(thr)
                cutoff[g] = np.sort(scores[called])[::-1][min(called.sum() - 1, 19)]
                selected[g] = [called.sum(), t.sum()]
        return {"counts": counts, "curve_len": curve_len, "cutoff": cutoff, "selected": selected}

    def curve(self, g, n):
        return self.curves[g]


@pytest.mark.parametrize("case", list(make_cases()), ids=lambda c: c[0])
def test_host_mirror_arithmetic_on_a_device_model(case):
    from variantcalling_b200 import concordance as PC

    name, df, classify_col, group_col = case
    want = load_golden()[name]
    check_accuracy(PC.calc_accuracy_metrics(df.copy(), classify_col, None, group_col, ctx=DeviceModel()), want["accuracy"])
    check_curve(PC.calc_recall_precision_curve(df.copy(), classify_col, None, group_col, ctx=DeviceModel()), want["curve"])
