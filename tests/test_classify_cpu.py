"""CPU: the restated per-record classification rules (oracle/classify_ref.py) reproduce, row by row, the output of
the reference's own nested functions and frame fix-ups (tests/golden/classify_rules.json, made by
scripts/make_golden_classify.py from comparison_utils.py:153-229)."""
import json
import os

from oracle import classify_ref as CR

GOLD = os.path.join(os.path.dirname(__file__), "golden", "classify_rules.json")


def load():
    rows = json.load(open(GOLD))["rows"]
    return ([tuple(r["gt_ultima"]) for r in rows], [tuple(r["gt_ground_truth"]) for r in rows], [r["base"] for r in rows],
            [r["classify"] for r in rows], [r["classify_gt"] for r in rows])


def test_restatement_equals_the_reference_rows():
    gu, gt, base, want_c, want_g = load()
    assert len(gu) == 2000
    got_c, got_g = CR.classify_records(gu, gt, base)
    assert got_c == want_c and got_g == want_g


def test_documented_examples():
    assert CR.classify((0, 1), (0, 1)) == "tp" and CR.classify_gt((0, 1), (1, 1)) == "fn"  # truth has fewer ref alleles
    assert CR.classify((None, None), (0, 1)) == "fn" and CR.classify((0, 1), (None,)) == "fp"
    assert CR.classify((1, 2), (0, 3)) == "fp" and CR.classify((0, 0), (0, 1)) == "fn"
