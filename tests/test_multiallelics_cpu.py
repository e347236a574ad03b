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


# ---------------------------------------------------------------- product host code (no GPU needed)
def cpu_index(text: bytes):
    """What the index pass (K0 + K1 with the no-model plan) returns, computed in Python."""
    from variantcalling_b200.lib import RECINFO_DTYPE

    lines = text.split(b"\n")[:-1]
    ls = np.zeros(len(lines) + 1, dtype=np.int64)
    ri = np.zeros(len(lines), dtype=RECINFO_DTYPE)
    at = 0
    for i, ln in enumerate(lines):
        c = ln.split(b"\t")
        ls[i] = at
        at += len(ln) + 1
        n_all = 1 + (0 if c[4] == b"." else c[4].count(b",") + 1)
        ri["pos"][i] = int(c[1])
        ri["flags"][i] = (n_all << 1) | (len(c[3]) << 8)
    ls[-1] = at
    return ls, ri


def contig_text(ds, contig):
    return ("\n".join(ln for ln in ds["lines"] if ln.split("\t", 1)[0] == contig) + "\n").encode()


@pytest.mark.parametrize("seed", [11, 12])
def test_product_split_rows_and_merge_match_the_oracle(seed):
    """variantcalling_b200.multiallelics rewrites the split rows as VCF lines; read back through the
    oracle loader + transformer they must give the features of the oracle's split frame, and its
    merge must give the oracle's merged likelihoods."""
    from tests import util
    from variantcalling_b200 import multiallelics as PM
    from variantcalling_b200.vcf_header import VcfHeader

    ds, tr, _model, _ = util.make_multiallelic_case(seed, "lr")
    hdr = VcfHeader(ds["header_text"])
    rng = np.random.default_rng(seed)
    for contig in ("chrM1", "chrM2"):
        df = R.get_vcf_df(ds["vf"], contig, ds["customs"])
        want_split = MR.process_multiallelic_spandel(df, ds["ref"][contig], ds["vf"].header)
        text = contig_text(ds, contig)
        ls, ri = cpu_index(text)
        plan = PM.SplitPlan(hdr, hdr.loader_columns(ds["customs"]), ds["ref"][contig])
        new_text = plan.build(np.frombuffer(text, dtype=np.uint8), ls, ri).tobytes()
        # same overlap sets as the reference's row loop
        sets = MR.overlapping_sets(df)
        assert sorted(g.origin for g in plan.groups) == sorted(sum(sets, []))
        got_df = R.get_vcf_df(OracleVariantFile(ds["header_text"].encode() + new_text), contig, ds["customs"])
        assert got_df.shape[0] == want_split.shape[0]
        with pd.option_context("future.infer_string", False):
            want_x = tr.transform(R.harness_float_columns(want_split)).to_numpy(dtype=np.float64)
            got_x = tr.transform(R.harness_float_columns(got_df)).to_numpy(dtype=np.float64)
        bad = np.argwhere(want_x.astype(np.float32) != got_x.astype(np.float32))
        assert bad.size == 0, (bad[:5], list(tr.get_feature_names_out())[bad[0][1]], want_x[tuple(bad[0])], got_x[tuple(bad[0])],
                               new_text.split(b"\n")[bad[0][0]])
        scores = rng.dirichlet(np.ones(3), size=want_split.shape[0])
        original = df.copy()
        src = [x in original.index for x in want_split.index]
        dst = [x in want_split.index for x in original.index]
        original["ml_lik"] = pd.Series([list(x) for x in scores[src, :]], index=original.loc[dst].index)
        merged = MR.combine_multiallelic_spandel(want_split, original, scores)
        lik = plan.merge(scores)
        for i, want in enumerate(merged["ml_lik"]):
            want = np.asarray(want, dtype=np.float64)
            assert np.array_equal(lik[i, :want.size], want) and not lik[i, want.size:].any(), (i, lik[i], want)
        ph, q, low = PM.score_math(lik, 30.0)
        ph_o, q_o, _gq = R.score_math(list(merged["ml_lik"]))
        assert np.array_equal(ph, ph_o) and np.array_equal(q, q_o) and np.array_equal(low, (q_o <= 30.0).astype(np.uint8))  # noqa: PLR2004


def test_find_overlaps_equals_the_row_loop_on_random_layouts():
    from variantcalling_b200 import multiallelics as PM

    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(1, 40))
        pos = np.sort(rng.integers(1, 60, size=n)).astype(np.int64)
        alleles = []
        for _i in range(n):
            kind = rng.random()
            ref = "A" * int(rng.integers(1, 6)) if kind < 0.4 else "A"
            alts = ["C"]
            if kind < 0.4 and rng.random() < 0.8:
                alts = ["A" * int(rng.integers(1, len(ref) + 1))]
            if rng.random() < 0.3:
                alts.append("*")
            if rng.random() < 0.3:
                alts.append("AT")
            alleles.append((ref, *alts))
        df = pd.DataFrame({"alleles": alleles, "pos": pos})
        want = MR.overlapping_sets(df)
        del_len = np.array([max(len(a[0]) - len(y) for y in a) for a in alleles])
        singles, clusters = PM.find_overlaps(pos, np.array([len(a) for a in alleles]), del_len,
                                             np.array(["*" in a for a in alleles]))
        assert sorted([[m] for m in singles] + clusters) == want


def test_device_rule_table_follows_the_header():
    """rules_blob (what ugvc_ma_set_rules takes): per-allele tags of the loaded columns are sub-sampled by their Number,
    the derived tags are replaced, everything else travels as it is (SplitPlan.convert / vcftools.py:687-778)."""
    import struct

    from tests import util
    from variantcalling_b200 import multiallelics as PM
    from variantcalling_b200.vcf_header import VcfHeader

    ds = util.make_dataset(n_records=50, n_custom=2, seed=3)
    extra = ['##INFO=<ID=GLS,Number=G,Type=Float,Description="x">', '##INFO=<ID=PAIR,Number=2,Type=Integer,Description="x">',
             '##FORMAT=<ID=GP,Number=G,Type=Float,Description="x">']
    hdr = VcfHeader("\n".join(ds["header"][:-1] + extra + ds["header"][-1:]) + "\n")
    cols = dict(hdr.loader_columns(ds["customs"]), gls="GLS", pair="PAIR", gp="GP")
    blob = PM.rules_blob(PM.SplitPlan(hdr, cols, ""))
    magic, n_info, n_fmt, flags = struct.unpack_from("<IIII", blob, 0)
    assert magic == 0x4D41524C and len(blob) == 148 + 36 * (n_info + n_fmt)  # noqa: PLR2004
    rules = {}
    for i in range(n_info + n_fmt):
        name, ln, action, number = struct.unpack_from("<32sBBH", blob, 148 + 36 * i)
        rules[(i >= n_info, name[:ln].decode())] = (action, number)
    assert rules[(False, "ac")][0] == PM.MA_SUB_A and rules[(False, "mq0c")][0] == PM.MA_SUB_R
    assert rules[(True, "ad")][0] == PM.MA_SUB_R
    assert (False, "hapcomp") not in rules and (False, "dp") not in rules  # Number=A overridden to 1; scalars are kept
    assert (True, "pl") not in rules and (True, "gt") not in rules          # handled by the split itself
    assert rules[(False, "gls")][0] == PM.MA_ERR_G and rules[(True, "gp")][0] == PM.MA_ERR_G
    assert rules[(False, "pair")] == (PM.MA_ERR_NUM, 2)
    assert [rules[(False, c)][0] for c in ("x_ic", "x_il", "x_hil", "x_hin")] == [PM.MA_SPECIAL + k for k in range(4)]
    assert flags & 1 and flags & 2 and (flags >> 4) & 0xF == 0xF  # noqa: PLR2004  QD / GQ in the header, the four derived tags loaded
    spell = [blob[16 + 32 * k: 16 + 32 * k + blob[144 + k]].decode() for k in range(4)]
    assert spell == ["X_IC", "X_IL", "X_HIL", "X_HIN"]


def test_fasta_reader_with_and_without_index(tmp_path):
    """read_fasta_contig: the .fai route (offset / line-width arithmetic of a samtools index) and the linear scan give
    the same sequence, as str and as bytes."""
    from tests import multiallelic_data as MD
    from variantcalling_b200 import multiallelics as PM

    ref = MD.make_reference(4)
    ref["tiny"] = "ACGTN"
    path = str(tmp_path / "ref.fa")
    with open(path, "w") as fh:
        fh.write(MD.fasta_text(ref, width=70))
    for name, seq in ref.items():
        assert PM.read_fasta_contig(path, name) == seq
        assert PM.read_fasta_contig(path, name, as_bytes=True) == seq.encode()
    with pytest.raises(KeyError):
        PM.read_fasta_contig(path, "chrNOPE")
    # samtools faidx columns: name, length, offset of the first base, bases per line, bytes per line
    fai, at = [], 0
    for name, seq in ref.items():
        at += len(name) + 2
        fai.append(f"{name}\t{len(seq)}\t{at}\t70\t71")
        at += len(seq) + (len(seq) + 69) // 70
    with open(path + ".fai", "w") as fh:
        fh.write("\n".join(fai) + "\n")
    for name, seq in ref.items():
        assert PM.read_fasta_contig(path, name) == seq
        assert PM.read_fasta_contig(path, name, as_bytes=True) == seq.encode()
    with pytest.raises(KeyError):
        PM.read_fasta_contig(path, "chrNOPE")


# ---------------------------------------------------------------- the reference's own in-code known answers
# ugbio_filtering tests/unit/test_multiallelics.py:14-120 (literal tables; shared with the device test)
REF_KAT_OVERLAP = dict(alleles=[["A", "T"], ["A", "T", "C"], ["A", "TA"], ["C", "A"], ["TA", "T"], ["T", "*", "A"], ["TAA", "T"],
                                ["A", "AAAT"]], positions=[10, 20, 30, 31, 40, 41, 60, 62], expected=[[1], [4, 5]])
REF_KAT_INDEL_SUBSET = [(("A", "G", "C"), (0, 1), False), (("A", "AG", "C"), (0, 1), True), (("A", "AG", "AC"), (1, 2), False),
                        (("A", "C", "AC"), (0, 2), True)]
REF_KAT_INDEL_SUBSET_SPANDEL = [(("A", "*", "C"), (0, 1), True), (("A", "*", "C"), (0, 2), False), (("A", "AG", "*"), (1, 2), True)]
REF_KAT_INDEL_CLASS = [(("A", "G", "C"), (0, 1), (("NA",), (None,))), (("A", "AG", "C"), (0, 1), (("ins",), (1,))),
                       (("A", "AG", "AC"), (1, 2), (("NA",), (None,))), (("A", "C", "AC"), (0, 2), (("ins",), (1,))),
                       (("A", "AC", "C"), (1, 2), (("del",), (1,)))]
REF_KAT_INDEL_CLASS_SPANDEL = [(("A", "G", "*"), (0, 1), (("NA",), (None,))), (("A", "AG", "*"), (0, 1), (("ins",), (1,))),
                               (("A", "C", "*"), (0, 2), (("del",), (4,))), (("A", "AC", "*"), (1, 2), (("del",), (4,)))]  # spandel x_il = (4, 5)
REF_KAT_HMER_REF = "A" * 20 + "G" + "ACCGCT" + "A" * 20
REF_KAT_HMER_SPANDEL = dict(alleles=("GA", "G"), pos=21)
REF_KAT_HMER = [(("A", "C"), (0, 1), (".", 0)), (("A", "C", "CA"), (1, 2), (".", 0)), (("A", "C", "CC"), (1, 2), ("C", 4)),
                (("A", "CC", "C"), (1, 2), ("C", 4)), (("A", "C", "*"), (1, 2), ("C", 3))]  # all at pos 22


def test_oracle_reproduces_the_reference_unit_test_tables():
    df = pd.DataFrame({"alleles": REF_KAT_OVERLAP["alleles"], "pos": REF_KAT_OVERLAP["positions"]})
    assert MR.overlapping_sets(df) == REF_KAT_OVERLAP["expected"]
    for gt, pair, want in [((0, 1), (1, 2), (0, 0)), ((1, 1), (1, 2), (0, 0)), ((2, 2), (1, 2), (1, 1)), ((1, 2), (1, 2), (0, 1))]:
        assert MR.gt_subset(gt, pair) == want
    for pl, pair, want in [((0, 10, 20), (0, 1), (0, 10, 20)), ((0, 10, 20, 40, 50, 60), (1, 2), (0, 30, 40)),
                           ((0, 10, 20, 40, 50, 60, 100, 120, 130, 140), (0, 3), (0, 100, 140))]:
        assert MR.pl_subset(pl, pair) == want
    for alleles, pair, want in REF_KAT_INDEL_SUBSET:
        assert MR.indel_subset(alleles, pair) == want
    for alleles, pair, want in REF_KAT_INDEL_SUBSET_SPANDEL:
        assert MR.indel_subset(alleles, pair, spandel=pd.Series([0])) == want
    for alleles, pair, want in REF_KAT_INDEL_CLASS:
        assert MR.indel_class_subset(alleles, pair) == want
    for alleles, pair, want in REF_KAT_INDEL_CLASS_SPANDEL:
        assert MR.indel_class_subset(alleles, pair, spandel=pd.Series({"x_il": (4, 5)})) == want
    for alleles, pair, want in REF_KAT_HMER:
        got = MR.hmer_indel_relative(alleles, pair, REF_KAT_HMER_REF, 22, spandel=pd.Series(REF_KAT_HMER_SPANDEL))
        assert got == want, (alleles, pair, got, want)


def test_oracle_flow_key_known_answers():
    # ugbio_core tests/unit/flow_format/test_flow_based_read.py:62-92 (generate_key_from_sequence, flow order ACGT)
    assert MR.flow_key("AAGGTTCC", "ACGT").tolist() == [2, 0, 2, 2, 0, 2]
    assert MR.flow_key("", "ACGT").tolist() == []
    with pytest.raises(ValueError):
        MR.flow_key("AAGGTTCCNN", "ACGT")
    # the product's Python model of the kernels spells the same keys in the TGCA order the branch uses
    from variantcalling_b200 import multiallelics as PM

    for seq in ("AAGGTTCC", "T", "ACGTTTGCA", "GGGG"):
        assert PM._flow_key(seq) == MR.flow_key(seq, "TGCA").tolist()  # noqa: SLF001
