"""GPU: a DeepVariant-style call set (the schema of the reference's real header fixture) with the
``deep_variant`` transformer flavour (transformers.py:241-245): features bit-identical to the reference
transformer mirror on the oracle frame, FILTER identical, scores within 1e-5; CLI output line for line."""
import gzip
import pickle

import numpy as np
import pandas as pd
import pytest

from oracle import ref_pipeline as R
from oracle.vcf_reader import OracleVariantFile
from tests import dv_data, util
from variantcalling_b200 import bgzf_io, lib
from variantcalling_b200 import filter_variants_pipeline as fvp
from variantcalling_b200 import model_compiler as MC
from variantcalling_b200 import transformers as T
from variantcalling_b200.tprep_constants import VcfType
from variantcalling_b200.vcf_header import VcfHeader

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dv():
    ds = dv_data.generate(4000, seed=17)
    ds["vf"] = OracleVariantFile(ds["header_text"].encode() + ds["text"])
    df = R.get_vcf_df(ds["vf"], None, ds["customs"])
    tr = T.get_transformer(VcfType.DEEP_VARIANT, [c.lower() for c in ds["customs"]])
    with pd.option_context("future.infer_string", False):
        x = tr.fit_transform(R.harness_float_columns(df)).to_numpy(dtype=np.float64)
    ds.update(df=df, tr=tr, x=x)
    return ds


@pytest.mark.parametrize("kind", ["lr", "gb_small"])
def test_features_filter_and_scores(gpu_ctx, dv, kind):
    model = util.fit_model(kind, dv["x"], dv["labels"])
    plan = MC.compile_plan(VcfHeader(dv["header_text"]), dv["tr"], model, dv["customs"])
    assert "vaf" in plan.feature_names and "qual" not in plan.feature_names
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.reserve(len(dv["text"]) + 1024, len(dv["lines"]) + 16, 1)
    exp = R.filter_variants(dv["vf"], model, dv["tr"], custom_annotations=dv["customs"])
    want_low = np.array(["LOW_SCORE" in f.split(";") for f in exp["filters"]])
    for mode in ("generic", "learned"):
        gpu_ctx.set_key_order(*(lib.learn_key_order(dv["text"]) if mode == "learned" else ("", "")))
        res = gpu_ctx.filter_batch(dv["text"], 30.0)
        feats = gpu_ctx.debug_features(res["n_records"]).T
        bad = np.argwhere(feats != dv["x"].astype(np.float32))
        assert bad.size == 0, f"[{mode}] first mismatch {bad[0]} ({plan.feature_names[bad[0][1]]})"
        assert np.array_equal(res["low_score"].astype(bool), want_low)
        np.testing.assert_allclose(res["probs"], exp["probs"], atol=1e-5, rtol=0)


def test_joint_callset_flavour(gpu_ctx, dv):
    """VcfType.JOINT: the common transform list alone (transformers.py:278; no `vaf`, no `qual`) on the same records --
    features, FILTER and scores against the oracle, with and without a learned key order."""
    tr = T.get_transformer(VcfType.JOINT, [c.lower() for c in dv["customs"]])
    with pd.option_context("future.infer_string", False):
        x = tr.fit_transform(R.harness_float_columns(dv["df"])).to_numpy(dtype=np.float64)
    model = util.fit_model("gb_small", x, dv["labels"])
    plan = MC.compile_plan(VcfHeader(dv["header_text"]), tr, model, dv["customs"])
    assert "vaf" not in plan.feature_names and "qual" not in plan.feature_names and x.shape[1] == plan.n_features
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.reserve(len(dv["text"]) + 1024, len(dv["lines"]) + 16, 1)
    exp = R.filter_variants(dv["vf"], model, tr, custom_annotations=dv["customs"])
    want_low = np.array(["LOW_SCORE" in f.split(";") for f in exp["filters"]])
    for mode in ("generic", "learned"):
        gpu_ctx.set_key_order(*(lib.learn_key_order(dv["text"]) if mode == "learned" else ("", "")))
        res = gpu_ctx.filter_batch(dv["text"], 30.0)
        feats = gpu_ctx.debug_features(res["n_records"]).T
        bad = np.argwhere(feats != x.astype(np.float32))
        assert bad.size == 0, f"[{mode}] first mismatch {bad[0]} ({plan.feature_names[bad[0][1]]})"
        assert np.array_equal(res["low_score"].astype(bool), want_low)
        np.testing.assert_allclose(res["probs"], exp["probs"], atol=1e-5, rtol=0)


def test_cli_on_a_deepvariant_file(dv, tmp_path):
    model = util.fit_model("rf", dv["x"], dv["labels"])
    vcf, mpath, out = str(tmp_path / "dv.vcf.gz"), str(tmp_path / "m.pkl"), str(tmp_path / "o.vcf.gz")
    bgzf_io.write_vcf_gz(vcf, dv["header"], dv["lines"])
    with open(mpath, "wb") as fh:
        pickle.dump({"xgb": model, "transformer": dv["tr"]}, fh)
    argv = ["--input_file", vcf, "--model_file", mpath, "--output_file", out, "--blacklist_cg_insertions"]
    for c in dv["customs"]:
        argv += ["--custom_annotations", c]
    fvp.run(argv)
    exp = R.filter_variants(dv["vf"], model, dv["tr"], custom_annotations=dv["customs"], blacklist_cg=True)
    text = gzip.open(out).read().decode().split("\n")[:-1]
    assert [ln for ln in text if ln.startswith("#")] == exp["header"]
    assert [ln for ln in text if not ln.startswith("#")] == exp["lines"]


@pytest.mark.parametrize("flavour,key", [(VcfType.DEEP_VARIANT, "features_deep_variant"), (VcfType.JOINT, "features_joint")])
def test_flavours_against_the_reference_transformers_own_output(gpu_ctx, flavour, key):
    """Device features bit-identical to what the REFERENCE's get_transformer(flavour).fit_transform produced on the same
    records (tests/golden/transformer_flavours.npz, scripts/make_golden_flavours.py)."""
    import os

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "transformer_flavours.npz"))
    text, customs = bytes(z["vcf_text"]), [str(c) for c in z["customs"]]
    at = text.index(b"#CHROM")
    at = text.index(b"\n", at) + 1
    header_text, records = text[:at].decode(), text[at:]
    df = R.harness_float_columns(R.get_vcf_df(OracleVariantFile(text), None, customs))
    tr = T.get_transformer(flavour, [c.lower() for c in customs])
    with pd.option_context("future.infer_string", False):
        x = tr.fit_transform(df).to_numpy(dtype=np.float64)
    labels = (np.arange(x.shape[0]) % 3 == 0).astype(int)
    plan = MC.compile_plan(VcfHeader(header_text), tr, util.fit_model("lr", x, labels), customs)
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.reserve(len(records) + 1024, x.shape[0] + 16, 1)
    for mode in ("generic", "learned"):
        gpu_ctx.set_key_order(*(lib.learn_key_order(records) if mode == "learned" else ("", "")))
        res = gpu_ctx.filter_batch(records, 30.0)
        feats = gpu_ctx.debug_features(res["n_records"]).T
        want = z[key].astype(np.float32)
        assert feats.shape == want.shape
        bad = np.argwhere(feats != want)
        assert bad.size == 0, f"[{mode}] first mismatch {bad[0]} ({plan.feature_names[bad[0][1]]})"
