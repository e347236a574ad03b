"""GPU: the CNV flavour of the transformer (transformers.py:291-313 + region_annotations,
:108-123,334): svtype / cnv_source fixed encodings, max(cn, copynumber) with nulls skipped, svlen
first element, region-annotation subset codes -- against the oracle running the transformer mirror
(whose CNV entries are pinned to the reference by tests/test_oracle_golden.py)."""
import numpy as np
import pandas as pd
import pytest

from oracle import ref_pipeline as R
from oracle.vcf_reader import OracleVariantFile
from variantcalling_b200 import lib
from variantcalling_b200 import model_compiler as MC
from variantcalling_b200 import transformers as T
from variantcalling_b200.tprep_constants import VcfType
from variantcalling_b200.vcf_header import VcfHeader

pytestmark = pytest.mark.gpu

FLOAT1 = ["pytorQ0", "pytorP2", "pytorRD", "pytorP1", "pytorP3", "GAP_PERCENTAGE", "CNV_DUP_FRAC", "CNV_DEL_FRAC"]
INT1 = ["CNV_DUP_READS", "CNV_DEL_READS", "JALIGN_DUP_SUPPORT", "JALIGN_DEL_SUPPORT", "JALIGN_DUP_SUPPORT_STRONG",
        "JALIGN_DEL_SUPPORT_STRONG"]
CUSTOM = ["SVTYPE"] + FLOAT1 + INT1 + ["SVLEN", "CN", "CopyNumber", "CNV_SOURCE", "REGION_ANNOTATIONS"]
REGIONS = ["Telomere_Centromere", "Clusters", "Coverage-Mappability"]


def make_cnv_vcf(n=400, seed=3):
    rng = np.random.default_rng(seed)
    hdr = ["##fileformat=VCFv4.2", '##INFO=<ID=SVTYPE,Number=1,Type=String,Description="t">']
    hdr += [f'##INFO=<ID={t},Number=1,Type=Float,Description="f">' for t in FLOAT1]
    hdr += [f'##INFO=<ID={t},Number=1,Type=Integer,Description="i">' for t in INT1]
    hdr += ['##INFO=<ID=SVLEN,Number=.,Type=Integer,Description="l">',
            '##INFO=<ID=CopyNumber,Number=1,Type=Float,Description="c">',
            '##INFO=<ID=CNV_SOURCE,Number=.,Type=String,Description="s">',
            '##INFO=<ID=REGION_ANNOTATIONS,Number=.,Type=String,Description="r">',
            '##FORMAT=<ID=GT,Number=1,Type=String,Description="g">', '##FORMAT=<ID=CN,Number=1,Type=Integer,Description="n">',
            "##contig=<ID=chr1,length=248956422>", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1"]
    lines = []
    pos = 1000
    for i in range(n):
        pos += int(rng.integers(1000, 100000))
        info = [f"SVTYPE={['DEL', 'DUP', 'NEUTRAL'][int(rng.integers(0, 3))]}"]
        for t in FLOAT1:
            if not (t.startswith("pytor") and rng.random() < 0.2):  # default_filler columns may be absent
                info.append(f"{t}={rng.uniform(0, 5):.3f}")
        for t in INT1:
            info.append(f"{t}={int(rng.integers(0, 60))}")
        info.append("SVLEN=" + ",".join(str(int(rng.integers(1000, 500000))) for _ in range(int(rng.integers(1, 3)))))
        u = rng.random()
        if u < 0.7:
            info.append(f"CopyNumber={rng.uniform(0, 6):.2f}")
        cn = "." if (u < 0.7 and rng.random() < 0.3) else str(int(rng.integers(0, 7)))
        info.append("CNV_SOURCE=" + ["cn.mops", "cnvpytor"][int(rng.integers(0, 2))])
        if rng.random() < 0.6:
            k = int(rng.integers(1, 4))
            info.append("REGION_ANNOTATIONS=" + ",".join(rng.permutation(REGIONS)[:k]))
        lines.append(f"chr1\t{pos}\t.\tN\t<CNV>\t{rng.uniform(1, 90):.1f}\t.\t{';'.join(info)}\tGT:CN\t0/1:{cn}")
    return hdr, lines


@pytest.fixture(scope="module")
def cnv():
    hdr, lines = make_cnv_vcf()
    vf = OracleVariantFile(("\n".join(hdr) + "\n" + "\n".join(lines) + "\n").encode())
    df = R.harness_float_columns(R.get_vcf_df(vf, None, CUSTOM))
    tr = T.get_transformer(VcfType.CNV, ["region_annotations"])
    with pd.option_context("future.infer_string", False):
        x = tr.fit_transform(df).to_numpy(dtype=np.float64)
    import os

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "transformer_cnv.npz"))
    assert np.array_equal(x, gold["features_ref"]), "mirror differs from the reference CNV transformer golden"
    from sklearn.ensemble import RandomForestClassifier

    y = (x[:, 1] + x[:, 16] > 5).astype(int)
    model = RandomForestClassifier(n_estimators=25, max_depth=4, random_state=42, n_jobs=1).fit(x, y)
    return dict(hdr=hdr, lines=lines, vf=vf, tr=tr, x=x, model=model)


def test_cnv_features_and_scores_match_oracle(gpu_ctx, cnv):
    assert cnv["x"].shape[1] == 19 and set(np.unique(cnv["x"][:, 0])) == {0.0, 1.0, 2.0}
    plan = MC.compile_plan(VcfHeader("\n".join(cnv["hdr"]) + "\n"), cnv["tr"], cnv["model"], CUSTOM)
    assert plan.feature_names[-3:] == ["copynumber", "cnv_source", "region_annotations"]
    text = ("\n".join(cnv["lines"]) + "\n").encode()
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.reserve(len(text) + 64, len(cnv["lines"]) + 8, 1)
    for mode in ("generic", "learned"):
        if mode == "learned":
            gpu_ctx.set_key_order(*lib.learn_key_order(text))
        res = gpu_ctx.filter_batch(text)
        feats = gpu_ctx.debug_features(res["n_records"]).T
        want = cnv["x"].astype(np.float32)
        bad = np.argwhere(feats != want)
        assert bad.size == 0, f"[{mode}] {len(bad)} mismatches, first {bad[0]}: {feats[tuple(bad[0])]} vs {want[tuple(bad[0])]}"
    exp = R.filter_variants(cnv["vf"], cnv["model"], cnv["tr"], custom_annotations=CUSTOM)
    np.testing.assert_allclose(res["probs"], exp["probs"], atol=1e-5, rtol=0)
    assert np.array_equal(res["low_score"].astype(bool), np.array(["LOW_SCORE" in f.split(";") for f in exp["filters"]]))


@pytest.mark.parametrize("old,new", [("SVTYPE=DEL", "SVTYPE=INV"), ("CNV_SOURCE=cn.mops", "CNV_SOURCE=cn.mops,cnvpytor"),
                                     ("CNV_SOURCE=cnvpytor", "CNV_SOURCE=other"),
                                     ("REGION_ANNOTATIONS=Clusters", "REGION_ANNOTATIONS=Clusters,Clusters"),
                                     ("REGION_ANNOTATIONS=Clusters", "REGION_ANNOTATIONS=Exome")])
def test_cnv_inputs_the_reference_raises_on(gpu_ctx, cnv, old, new):
    idx = next(i for i, l in enumerate(cnv["lines"]) if old + ";" in l + ";" or l.split("\t")[7].endswith(old) or old + "," in l)
    lines = list(cnv["lines"][:60])
    src = cnv["lines"][idx]
    key = old.split("=")[0]
    fields = src.split("\t")
    info = [kv if not kv.startswith(key + "=") else new for kv in fields[7].split(";")]
    fields[7] = ";".join(info)
    lines[10] = "\t".join(fields[:1] + [lines[10].split("\t")[1]] + fields[2:])
    text = ("\n".join(lines) + "\n").encode()
    with pytest.raises(Exception):  # noqa: B017
        R.filter_variants(OracleVariantFile(("\n".join(cnv["hdr"]) + "\n").encode() + text), cnv["model"], cnv["tr"],
                          custom_annotations=CUSTOM)
    gpu_ctx.load_plan(MC.compile_plan(VcfHeader("\n".join(cnv["hdr"]) + "\n"), cnv["tr"], cnv["model"], CUSTOM).blob)
    gpu_ctx.reserve(1 << 20, 1024, 1)
    with pytest.raises(lib.UgvcDataError):
        gpu_ctx.filter_batch(text)
    assert gpu_ctx.last_data_error()[0] == 10
