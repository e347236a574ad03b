"""GPU parity: the CUDA path through the C ABI vs the oracle (reference restatement) on the
same seeded VCF text.  FILTER decision bit-identical, features bit-identical (fp32), scores
within 1e-5 (BASELINE.json north_star tolerance)."""
import numpy as np
import pytest

from oracle import ref_pipeline as R
from tests import util
from variantcalling_b200 import lib
from variantcalling_b200 import model_compiler as MC
from variantcalling_b200.vcf_header import VcfHeader

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def ds():
    d = util.make_dataset(n_records=4000, n_custom=5)
    d["df"], d["tr"], d["x"] = util.fit_transformer(d)
    return d


WRONG_ORDER = "X_RM;X_LM;DP;AC;NOPE;QD;AF!;SOR;AN"  # deliberately scrambled / partly bogus schedule


@pytest.mark.parametrize("kind,order", [("lr", "none"), ("gb_small", "learned"), ("rf", "wrong"), ("gb", "learned"),
                                        ("gb_small", "none")])
def test_filter_batch_matches_oracle(gpu_ctx, ds, kind, order):
    model = util.fit_model(kind, ds["x"], ds["labels"])
    plan = MC.compile_plan(VcfHeader(ds["header_text"]), ds["tr"], model, ds["customs"])
    gpu_ctx.load_plan(plan.blob)
    if order == "learned":  # K1's schedule-driven path
        info, fmt = lib.learn_key_order(ds["text"])
        assert info.startswith("AC;AF;AN") and fmt == "GT:AD:DP:GQ:PL"
        gpu_ctx.set_key_order(info, fmt)
    elif order == "wrong":  # a bad schedule must only cost speed, never change results
        gpu_ctx.set_key_order(WRONG_ORDER, "GT:DP")
    gpu_ctx.reserve(len(ds["text"]) + 1024, len(ds["lines"]) + 16, 1)
    gpu_ctx.counts_reset()
    res = gpu_ctx.filter_batch(ds["text"], 30.0)
    n = res["n_records"]
    assert n == len(ds["lines"])
    # features: bit-identical to the reference transformer output cast to fp32
    feats = gpu_ctx.debug_features(n).T
    want = ds["x"].astype(np.float32)
    assert feats.shape == want.shape
    bad = np.argwhere(feats != want)
    assert bad.size == 0, f"feature mismatch at {bad[:5]}: {feats[tuple(bad[0])]} vs {want[tuple(bad[0])]}"
    # oracle scores
    exp = R.filter_variants(ds["vf"], model, ds["tr"], custom_annotations=ds["customs"], decision_threshold=30.0)
    np.testing.assert_allclose(res["probs"], exp["probs"], atol=TOL, rtol=0)
    np.testing.assert_allclose(res["qual"], exp["quals"], atol=1e-4, rtol=1e-6)
    low = np.array(["LOW_SCORE" in f.split(";") for f in exp["filters"]])
    assert np.array_equal(res["low_score"].astype(bool), low), "FILTER decision differs from the oracle"
    c = gpu_ctx.counts()
    assert c["n_records"] == n and c["n_low_score"] == int(low.sum()) and c["n_pass"] == n - int(low.sum())
    # writer support: offsets point at the right columns
    ri, ls = res["recinfo"], res["line_start"]
    for i in (0, 1, n // 2, n - 1):
        line = ds["lines"][i]
        cols = line.split("\t")
        assert ri["pos"][i] == int(cols[1])
        start = lambda k: len("\t".join(cols[:k])) + 1  # noqa: E731
        assert (ri["qual_off"][i], ri["filter_off"][i], ri["info_off"][i], ri["format_off"][i]) == (
            start(5), start(6), start(7), start(8))
        assert ds["text"][ls[i]:ls[i + 1] - 1].decode() == line


@pytest.mark.parametrize("n_class", [2, 3])
def test_xgboost_json_model_path(gpu_ctx, ds, n_class):
    """MODEL_XGB (x < t, fp32 margins in tree order, fp32 sigmoid / softmax) against the NumPy
    restatement of xgboost's predictor on an xgboost-format JSON document."""
    from oracle import xgb_predictor as XP

    y = ds["labels"] if n_class == 2 else np.where(ds["x"][:, 2] > 0, 2, ds["labels"])
    gb = util.fit_model("gb_small" if n_class == 2 else "gb3", ds["x"], y)
    doc = XP.sklearn_gb_to_xgb_json(gb)
    plan = MC.compile_plan(VcfHeader(ds["header_text"]), ds["tr"], doc, ds["customs"])
    assert plan.model_kind == MC.MODEL_XGB and plan.n_classes == n_class
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.set_key_order(*lib.learn_key_order(ds["text"]))
    gpu_ctx.reserve(len(ds["text"]) + 1024, len(ds["lines"]) + 16, 1)
    res = gpu_ctx.filter_batch(ds["text"], 30.0)
    want = XP.predict_proba(doc, ds["x"].astype(np.float32))
    np.testing.assert_allclose(res["probs"], want, atol=2e-6, rtol=0)
    _, quals, _ = R.score_math(want)
    np.testing.assert_allclose(res["qual"], quals, atol=2e-4, rtol=0)
    far = np.abs(quals - 30.0) > 1e-3  # fp32 expf may differ by an ulp between libm and CUDA near the threshold
    assert np.array_equal(res["low_score"].astype(bool)[far], (quals <= 30.0)[far])
    if n_class == 2:  # same trees, same prior: sklearn's fp64 evaluation agrees to fp32 accuracy
        assert np.abs(gb.predict_proba(ds["x"]) - want).max() < 1e-5


def test_multinomial_logistic_regression(gpu_ctx, ds):
    y = np.where(ds["x"][:, 2] > 0, 2, ds["labels"])
    model = util.fit_model("lr", ds["x"], y)
    assert model.coef_.shape[0] == 3
    plan = MC.compile_plan(VcfHeader(ds["header_text"]), ds["tr"], model, ds["customs"])
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.reserve(len(ds["text"]) + 1024, len(ds["lines"]) + 16, 1)
    res = gpu_ctx.filter_batch(ds["text"], 30.0)
    exp = R.filter_variants(ds["vf"], model, ds["tr"], custom_annotations=ds["customs"])
    np.testing.assert_allclose(res["probs"], exp["probs"], atol=TOL, rtol=0)
    assert np.array_equal(res["low_score"].astype(bool), np.array(["LOW_SCORE" in f.split(";") for f in exp["filters"]]))
