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
