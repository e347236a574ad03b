"""CPU: the oracle and the host-side transformer mirror against (a) the reference's in-code
known answers and (b) golden vectors generated from the reference's own modules
(scripts/make_golden.py)."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import ref_pipeline as R
from oracle.vcf_reader import OracleVariantFile
from variantcalling_b200 import transformers as T
from variantcalling_b200.tprep_constants import VcfType

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def kats():
    return json.load(open(os.path.join(GOLD, "kats.json")))


def test_reference_unit_test_answers():
    # ugbio_utils/src/filtering/tests/unit/test_transformers.py:43-100
    assert T.tuple_break((1, 2, 3)) == 1
    assert T.tuple_break_second((1, 2, 3)) == 2
    assert T.tuple_break_third((1, 2, 3)) == 3
    assert T.motif_encode_left("ATGC") == 4321
    assert T.motif_encode_right("ATGC") == 1234
    assert T.allele_encode("A") == 1
    assert T.gt_encode((1, 1)) == 1 and T.gt_encode((1, 0)) == 0
    assert (T.ins_del_encode("ins"), T.ins_del_encode("del"), T.ins_del_encode("NA")) == (-1, 1, 0)
    assert T.encode_labels([(0, 1), (0, 0), (1, 0), (1, 1)]) == [1, 0, 1, 2]
    with pytest.raises(ValueError):
        T.encode_labels([(0, 1, 2), (0, 0, 1)])
    with pytest.raises(ValueError):
        T.encode_labels([(0, 2), (1, 2)])
    assert T.region_annotation_encode(("Telomere_Centromere", "Clusters")) > 0
    assert T.region_annotation_encode(()) > 0
    with pytest.raises(KeyError):
        T.region_annotation_encode(("Unknown",))
    # core/tests/unit/test_math_utils.py:5-6
    assert np.all(R.phred((0.1, 0.01, 0.001)) == np.array([10.0, 20.0, 30.0]))


def test_kats_from_reference_functions(kats):
    for m, v in kats["motif_encode_left"].items():
        assert T.motif_encode_left(m) == v
    for m, v in kats["motif_encode_right"].items():
        assert T.motif_encode_right(m) == v
    assert T.motif_encode_left(("ACGTA",)) == kats["motif_encode_left_tuple"]
    assert T.motif_encode_left(("A", "T")) == kats["motif_encode_left_tuple_single"]
    for a, v in kats["allele_encode"].items():
        assert T.allele_encode(a) == v
    for g, v in kats["gt_encode"]:
        assert T.gt_encode(tuple(g)) == v
    for k, v in kats["ins_del_encode"].items():
        assert T.ins_del_encode(k) == v
    for k, v in kats["region_annotation_encode"].items():
        assert T.region_annotation_encode(tuple(x for x in k.split(",") if x)) == v
    for p, v in kats["phred"]:
        assert float(R.phred([p])[0]) == v


def test_cnv_transformer_copynumber_column():
    # test_transformers.py:102-137 (copynumber = max(cn, copynumber))
    tr = T.get_transformer(VcfType.CNV)
    df = pd.DataFrame({
        "svtype": ["DEL"], "pytorq0": [0.1], "pytorp2": [0.2], "pytorrd": [0.3], "pytorp1": [0.4], "pytorp3": [0.5],
        "gap_percentage": [0.01], "cnv_dup_reads": [10], "cnv_del_reads": [5], "cnv_dup_frac": [0.6],
        "cnv_del_frac": [0.3], "jalign_dup_support": [8], "jalign_del_support": [4],
        "jalign_dup_support_strong": [6], "jalign_del_support_strong": [3], "svlen": [(1000,)], "cn": [2],
        "copynumber": [3], "cnv_source": [("cn.mops",)]})
    res = tr.fit_transform(df)
    col = [c for c in res.columns if c.startswith("copynumber__") or c.startswith("tmp_col_name_16_")]
    assert res.to_numpy()[0, 16] == 3 or (col and res[col[0]].iloc[0] == 3)


def test_blacklist_rules():
    # filtering/tests/unit/test_variant_filtering_utils.py:17-40
    rows = pd.DataFrame({"alleles": [("C", "T"), ("CCG", "C"), ("G", "GGC")], "filter": ["PASS"] * 3})
    assert list(R.blacklist_cg_insertions(rows)) == ["PASS", "CG_NON_HMER_INDEL", "CG_NON_HMER_INDEL"]
    merged = R.merge_blacklists([pd.Series(["PASS", "FAIL", "FAIL"]), pd.Series(["PASS", "FAIL1", "PASS"])])
    assert list(merged) == ["PASS;PASS", "FAIL;FAIL1", "FAIL;PASS"]


def test_validate_data():
    # test_variant_filtering_utils.py:62-77
    R.validate_data(np.array([[0, 1], [1, 2]]))
    with pytest.raises(AssertionError):
        R.validate_data(np.array([[0, 1], [1, np.nan]]))
    R.validate_data(pd.Series([0, 1]))
    with pytest.raises(AssertionError):
        R.validate_data(pd.Series([1, np.nan]))


def test_get_gt_from_pl_idx_table():
    # filtering/tests/unit/test_multiallelics.py:140-160
    want = {0: (0, 0), 1: (0, 1), 2: (1, 1), 3: (0, 2), 4: (1, 2), 5: (2, 2), 6: (0, 3), 9: (3, 3), 55: (0, 10)}
    for idx, gt in want.items():
        assert R.get_gt_from_pl_idx(idx) == gt


def _golden():
    z = np.load(os.path.join(GOLD, "transformer_single_sample.npz"))
    return bytes(z["vcf_text"]), [str(c) for c in z["customs"]], z["features_ref"], json.loads(str(z["categories"]))


def test_mirror_transformer_matches_reference_golden():
    """oracle loader + host mirror of get_transformer == the reference module's fit_transform."""
    text, customs, feats_ref, cats = _golden()
    vf = OracleVariantFile(text)
    df = R.harness_float_columns(R.get_vcf_df(vf, None, customs))
    tr = T.get_transformer(VcfType.SINGLE_SAMPLE, [c.lower() for c in customs])
    with pd.option_context("future.infer_string", False):
        x = tr.fit_transform(df).to_numpy(dtype=np.float64)
    assert x.shape == feats_ref.shape
    assert np.array_equal(x, feats_ref)
    for name, trans, _ in tr.transformers_:
        last = trans.steps[-1][1] if hasattr(trans, "steps") else trans
        if hasattr(last, "categories_"):
            assert [str(c) for c in last.categories_[0]] == cats[name]


def test_oracle_typed_decode_rules():
    """Hand-written snippets with the typed values the VCF 4.2 spec + htslib float32 rule give."""
    hdr = "\n".join([
        "##fileformat=VCFv4.2",
        '##INFO=<ID=DP,Number=1,Type=Integer,Description="d">', '##INFO=<ID=AF,Number=A,Type=Float,Description="a">',
        '##INFO=<ID=DB,Number=0,Type=Flag,Description="f">', '##INFO=<ID=LCR,Number=1,Type=String,Description="s">',
        '##INFO=<ID=X_IC,Number=A,Type=String,Description="s">',
        '##FORMAT=<ID=GT,Number=1,Type=String,Description="g">', '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="d">',
        '##FORMAT=<ID=AD,Number=R,Type=Integer,Description="a">', '##FORMAT=<ID=PL,Number=G,Type=Integer,Description="p">',
        "##contig=<ID=c1,length=1000>", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS"])
    recs = ["c1\t5\t.\tA\tC,G\t0.1\tq10;s50\tDP=7;AF=0.1,.;DB;LCR=TRUE;X_IC=ins,NA\tGT:DP:AD:PL\t1|2:.:3,.,4:1,2",
            "c1\t9\trs9\tAT\t.\t.\t.\t.\tGT\t./."]
    vf = OracleVariantFile((hdr + "\n" + "\n".join(recs) + "\n").encode())
    r0, r1 = list(vf)
    assert r0.qual == float(np.float32(0.1)) and r0.qual != 0.1
    assert r0.info["DP"] == 7 and r0.info["AF"] == (float(np.float32(0.1)), None) and r0.info["DB"] is True
    assert r0.info["LCR"] == "TRUE" and r0.info["X_IC"] == ("ins", "NA")
    assert r0.sample["GT"] == (1, 2) and r0.sample["DP"] is None and r0.sample["AD"] == (3, None, 4)
    assert r0.sample["PL"] == (1, 2) and r0.alleles == ("A", "C", "G") and r0.filter_keys == ["q10", "s50"]
    assert r1.qual is None and r1.alleles == ("AT",) and r1.id == "rs9" and r1.filter_keys == []
    assert r1.sample["GT"] == (None, None) and r1.info == {}
    df = R.get_vcf_df(vf, None, ["LCR"])
    # FORMAT wins over INFO for a shared name (vcftools.py:69-86): DP of record 0 is the sample's missing DP
    assert pd.isna(df["dp"].iloc[0]) and list(df["indel"]) == [False, False]
    assert df.index[0] == ("c1", 5) and df["filter"].iloc[0] == "q10;s50" and df["filter"].iloc[1] == ""


def test_writer_rules():
    hdr = ["##fileformat=VCFv4.2", "##contig=<ID=c1,length=10>", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"]
    vf = OracleVariantFile(("\n".join(hdr) + "\nc1\t1\t.\tA\tC\t5\tPASS\tX=1\nc1\t2\t.\tA\tC\t5\tq10\t.\n"
                            "c1\t3\t.\tA\tC\t5\t.\tTREE_SCORE=1;Y=2\n").encode())
    recs = list(vf)
    line, keys = R.write_record(recs[0], 12.5, 30.0, overwrite_qual=False, blacklist_value=None)
    assert keys == ["LOW_SCORE"] and line.split("\t")[6:8] == ["LOW_SCORE", "X=1;TREE_SCORE=12.5"]
    line, keys = R.write_record(recs[1], 30.0, 30.0, overwrite_qual=True, blacklist_value="PASS;CG_NON_HMER_INDEL")
    assert keys == ["q10", "LOW_SCORE"] and line.split("\t")[5] == "30"
    assert line.split("\t")[7] == "TREE_SCORE=30;BLACKLST=CG_NON_HMER_INDEL"
    line, keys = R.write_record(recs[2], 45.123456789, 30.0, overwrite_qual=False, blacklist_value="PASS")
    assert keys == ["PASS"] and line.split("\t")[7] == "TREE_SCORE=45.1235;Y=2"
    out = R.edited_header_lines(hdr, with_model=True, with_blacklist=True)
    assert out[-1].startswith("#CHROM") and sum(l.startswith("##FILTER=<ID=LOW_SCORE") for l in out) == 1
    assert any(l.startswith("##INFO=<ID=TREE_SCORE,Number=1,Type=Float") for l in out)
    assert any(l.startswith("##INFO=<ID=BLACKLST,Number=.,Type=String") for l in out)


def test_mirror_cnv_transformer_matches_reference_golden():
    """CNV flavour (+ region_annotations): mirror == the reference module's fit_transform."""
    z = np.load(os.path.join(GOLD, "transformer_cnv.npz"))
    text, customs, feats_ref = bytes(z["vcf_text"]), [str(c) for c in z["customs"]], z["features_ref"]
    df = R.harness_float_columns(R.get_vcf_df(OracleVariantFile(text), None, customs))
    tr = T.get_transformer(VcfType.CNV, ["region_annotations"])
    with pd.option_context("future.infer_string", False):
        x = tr.fit_transform(df).to_numpy(dtype=np.float64)
    assert x.shape == feats_ref.shape == (400, 19)
    assert np.array_equal(x, feats_ref)


def test_oracle_reader_scalar_versus_tuple_rule_known_answers():
    """The few facts about pysam's typed decode that the reference's own (in-code) unit tests assert,
    re-checked on the oracle reader -- they pin the Number=1 -> scalar / otherwise -> tuple rule and the
    Integer / Float / String typing (everything else about that boundary stays "parity unpinned"):
      * INFO Integer Number=1 reads back as an int:  ``first_record.info["CNV_TOTAL_READS"] == 2``
        (ugbio_cnv/tests/unit/test_analyze_cnv_breakpoint_reads.py:232-234; header analyze_cnv_breakpoint_reads.py:443-445)
      * INFO Integer Number=. reads back as a tuple even with one element: ``record.info["SVLEN"] == (n,)``
        (tests/unit/test_combine_cnv_vcf_utils.py:656; header cnmops_utils.py:68)
      * INFO String Number=. -> tuple of str: ``record.info["CNV_SOURCE"] == ("CNVpytor",)``
        (tests/unit/test_combine_cnv_vcf_utils.py:222,268,349; header cnv_vcf_consts.py:13-19)
      * INFO Float Number=1 -> float: ``records[0].info["GAP_PERCENTAGE"] == 1.0`` / ``0.5``
        (tests/unit/test_combine_cnmops_cnvpytor_cnv_calls_unit.py:109-216; header cnv_vcf_consts.py:48-54)
    """
    from oracle.vcf_reader import OracleVariantFile

    text = ("##fileformat=VCFv4.2\n"
            '##INFO=<ID=CNV_TOTAL_READS,Number=1,Type=Integer,Description="x">\n'
            '##INFO=<ID=CNV_DUP_READS,Number=1,Type=Integer,Description="x">\n'
            '##INFO=<ID=SVLEN,Number=.,Type=Integer,Description="CNV length">\n'
            '##INFO=<ID=CNV_SOURCE,Number=.,Type=String,Description="tool">\n'
            '##INFO=<ID=GAP_PERCENTAGE,Number=1,Type=Float,Description="fraction">\n'
            "##contig=<ID=chr1,length=1000000>\n"
            "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
            "chr1\t2500\t.\tN\t<DEL>\t.\tPASS\tCNV_TOTAL_READS=2;CNV_DUP_READS=1;SVLEN=2000;CNV_SOURCE=CNVpytor;GAP_PERCENTAGE=0.5\n"
            "chr1\t9000\t.\tN\t<DUP>\t.\tPASS\tSVLEN=100,200;CNV_SOURCE=cn.mops,cnvpytor;GAP_PERCENTAGE=1\n")
    first, second = list(OracleVariantFile(text.encode()))
    assert first.info["CNV_TOTAL_READS"] == 2 and isinstance(first.info["CNV_TOTAL_READS"], int)
    assert first.info["CNV_DUP_READS"] == 1
    assert first.info["SVLEN"] == (2000,)
    assert first.info["CNV_SOURCE"] == ("CNVpytor",)
    assert first.info["GAP_PERCENTAGE"] == 0.5 and isinstance(first.info["GAP_PERCENTAGE"], float)
    assert second.info["SVLEN"] == (100, 200) and second.info["CNV_SOURCE"] == ("cn.mops", "cnvpytor")
    assert second.info["GAP_PERCENTAGE"] == 1.0


@pytest.mark.parametrize("flavour,key", [(VcfType.DEEP_VARIANT, "features_deep_variant"), (VcfType.JOINT, "features_joint")])
def test_mirror_transformer_flavours_match_reference_golden(flavour, key):
    """deep_variant / joint_callset: oracle loader + host mirror == the reference module's own fit_transform
    (tests/golden/transformer_flavours.npz, scripts/make_golden_flavours.py; transformers.py:221-245,278)."""
    z = np.load(os.path.join(GOLD, "transformer_flavours.npz"))
    customs = [str(c) for c in z["customs"]]
    df = R.harness_float_columns(R.get_vcf_df(OracleVariantFile(bytes(z["vcf_text"])), None, customs))
    tr = T.get_transformer(flavour, [c.lower() for c in customs])
    with pd.option_context("future.infer_string", False):
        x = tr.fit_transform(df).to_numpy(dtype=np.float64)
    assert x.shape == z[key].shape and np.array_equal(x, z[key], equal_nan=True)
