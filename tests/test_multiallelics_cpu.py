"""CPU: the oracle's restatement of the --treat_multiallelics branch against golden data frames
produced by the reference's own functions (scripts/make_golden_multiallelic.py)."""
import gzip
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import multiallelic_ref as MR
from oracle import ref_pipeline as R
from oracle.vcf_reader import OracleVariantFile

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load_golden(seed):
    with gzip.open(os.path.join(GOLDEN, f"multiallelic_split_{seed}.json.gz"), "rt") as fh:
        return json.load(fh)


def read_fasta(text):
    out, name = {}, None
    for ln in text.splitlines():
        if ln.startswith(">"):
            name = ln[1:].split()[0]
            out[name] = []
        elif name:
            out[name].append(ln)
    return {k: "".join(v) for k, v in out.items()}


def cell(v):
    """Same normalisation as the golden writer."""
    if isinstance(v, (tuple, list, np.ndarray)):
        return [cell(x) for x in v]
    if v is None:
        return None
    if isinstance(v, (bool, np.bool_)):
        return bool(v)
    if isinstance(v, (int, np.integer)):
        return int(v)
    if isinstance(v, (float, np.floating)):
        return "NaN" if np.isnan(v) else float(v)
    return str(v)


@pytest.mark.parametrize("seed", [11])
def test_split_and_merge_match_the_reference(seed):
    g = load_golden(seed)
    vf = OracleVariantFile(g["vcf_text"].encode())
    ref = read_fasta(g["fasta"])
    for contig, want in g["contigs"].items():
        df = R.get_vcf_df(vf, contig, g["customs"])
        split = MR.process_multiallelic_spandel(df, ref[contig], vf.header)
        assert list(split.columns) == want["columns"]
        assert [str(t) for t in split.dtypes] == want["dtypes"]
        assert [cell(i) for i in split.index] == want["index"]
        got_rows = [[cell(v) for v in row] for row in split.itertuples(index=False, name=None)]
        for r, (a, b) in enumerate(zip(got_rows, want["rows"])):
            assert a == b, f"{contig} row {r}: " + str([(c, x, y) for c, x, y in zip(want['columns'], a, b) if x != y])
        assert len(got_rows) == len(want["rows"])
        scores = np.array(want["scores"])
        original = df.copy()
        src = [x in original.index for x in split.index]
        dst = [x in split.index for x in original.index]
        original["ml_lik"] = pd.Series([list(x) for x in scores[src, :]], index=original.loc[dst].index)
        merged = MR.combine_multiallelic_spandel(split, original, scores)
        got = [cell(v) for v in merged["ml_lik"]]
        assert got == want["ml_lik"]


def test_index_helpers_known_answers():
    # ugbio_filtering tests/unit/test_multiallelics.py:31-55,140-160 (in-code tables)
    assert [MR.pl_index(p) for p in ((0, 0), (0, 1), (1, 1), (0, 2), (1, 2), (2, 2), (0, 3))] == [0, 1, 2, 3, 4, 5, 6]
    assert [R.get_gt_from_pl_idx(i) for i in (0, 1, 2, 3, 4, 5, 55)] == [(0, 0), (0, 1), (1, 1), (0, 2), (1, 2), (2, 2), (0, 10)]
    assert MR.gt_subset((1, 2), (0, 1)) == (1, 1) and MR.gt_subset((1, 2), (1, 2)) == (0, 1) and MR.gt_subset((0, 2), (0, 1)) == (0, 0)
    with pytest.raises(AssertionError):
        MR.gt_subset((0, 0), (1, 2))
    assert MR.pl_subset((10, 20, 30, 40, 50, 60), (1, 2)) == (0, 20, 30)
    assert MR.pl_subset((10, 20, 30, 40, 50, 60), (0, 2), normed=False) == (10, 40, 60)
