"""GPU: the model-apply step alone (ugvc_predict_features / variantcalling_b200.model_apply) against
the fitted estimators' own predict_proba / predict, and apply_model against the oracle's."""
import numpy as np
import pandas as pd
import pytest

from oracle import ref_pipeline as R
from oracle import xgb_predictor as XP
from tests import util
from variantcalling_b200 import model_apply as MA

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def data():
    ds = util.make_dataset(n_records=6000, n_custom=4, seed=31)
    df, tr, x = util.fit_transformer(ds)
    y3 = np.where(x[:, 2] > 0, 2, ds["labels"])
    return dict(ds=ds, df=df, tr=tr, x=x, y=ds["labels"], y3=y3)


@pytest.mark.parametrize("kind,three", [("lr", False), ("lr", True), ("gb_small", False), ("gb3", True), ("rf", False), ("rf", True)])
def test_predict_proba_equals_the_estimator(data, kind, three):
    model = util.fit_model(kind, data["x"], data["y3"] if three else data["y"])
    clf = MA.GpuClassifier(model, max_rows=2500)  # smaller than the matrix: exercises the chunk loop
    x32 = data["x"].astype(np.float32)
    want = model.predict_proba(x32.astype(np.float64) if kind == "lr" else x32)
    got = clf.predict_proba(x32)
    assert got.shape == want.shape and got.dtype == np.float64
    np.testing.assert_allclose(got, want, atol=1e-12 if kind != "lr" else 1e-9, rtol=0)
    assert np.array_equal(clf.predict(x32), model.classes_[np.argmax(want, axis=1)])
    with pytest.raises(ValueError, match="NaN or infinity"):
        clf.predict_proba(np.where(np.arange(x32.size).reshape(x32.shape) == 7, np.nan, x32))
    with pytest.raises(ValueError, match="features"):
        clf.predict_proba(x32[:, :-1])
    clf.close()


def test_xgboost_json_model_in_fp32(data):
    gb = util.fit_model("gb_small", data["x"], data["y"])
    doc = XP.sklearn_gb_to_xgb_json(gb)
    clf = MA.GpuClassifier(doc, n_features=data["x"].shape[1])
    x32 = data["x"].astype(np.float32)
    want = XP.predict_proba(doc, x32)
    np.testing.assert_allclose(clf.predict_proba(x32), want, atol=2e-6, rtol=0)  # fp32 expf: an ulp between libraries
    clf.close()


def test_apply_model_matches_the_oracle(data):
    model = util.fit_model("gb_small", data["x"], data["y"])
    with pd.option_context("future.infer_string", False):
        pred, prob = MA.apply_model(R.harness_float_columns(data["df"]), model, data["tr"])
        want_pred, want_prob = R.apply_model(data["df"], model, data["tr"])
    np.testing.assert_allclose(prob, want_prob, atol=1e-6, rtol=0)  # the estimator sees fp64 features, K3 fp32 ones
    assert (pred == want_pred).mean() > 0.9999


def test_c_abi_contract(gpu_ctx, data):
    from variantcalling_b200 import lib
    from variantcalling_b200 import model_compiler as MC

    model = util.fit_model("lr", data["x"], data["y"])
    gpu_ctx.load_plan(MC.compile_plan_model_only(model, data["x"].shape[1]).blob)
    gpu_ctx.reserve(4096, 1000, 1)
    x32 = data["x"].astype(np.float32)
    res = gpu_ctx.predict_features(x32[:1000], threshold=25.0)
    phreds = -10 * np.log10(model.predict_proba(x32[:1000].astype(np.float64)) + 1e-10)
    qual = np.clip(30 + phreds[:, 0] - phreds[:, 1], 0, None)
    np.testing.assert_allclose(res["qual"], qual, atol=1e-6, rtol=0)
    assert np.array_equal(res["low_score"].astype(bool), res["qual"] <= 25.0)
    assert gpu_ctx.predict_features(x32[:0])["n_records"] == 0
    with pytest.raises(lib.UgvcError):
        gpu_ctx.predict_features(x32[:1025])                      # more rows than reserved (1000 rounds up to 1024)
    bad = x32[:10].copy()
    bad[3, 5] = np.inf
    with pytest.raises(lib.UgvcDataError):
        gpu_ctx.predict_features(bad)
    assert gpu_ctx.last_data_error()[:2] == (3, 5)


def test_featuremap_predict_record_with_xgb(data, tmp_path):
    """featuremap_xgb_prediction.predict_record_with_xgb (:301-323): columns picked by the booster's feature names, object
    columns label-encoded per call (sorted distinct strings), nulls -> 0, probability of class 1 -- against the
    restated xgboost predictor on the frame prepared the reference's way (sklearn's LabelEncoder itself)."""
    import json

    from sklearn.preprocessing import LabelEncoder

    rng = np.random.default_rng(8)
    n = 3000
    with pd.option_context("future.infer_string", False):  # harness: object columns, as the reference's pandas builds them
        frame = pd.DataFrame({
            "x_qual_mean": rng.normal(30, 5, n), "alt_reads": rng.integers(1, 40, n),
            "ref_allele": rng.choice(list("ACGT"), n).astype(object),  # object columns, as the reference's pandas builds them
            "alt_allele": rng.choice(["A", "C", "G", "T", "AT", "10", "9"], n).astype(object), "is_cycle_skip": rng.random(n) < 0.3,
            "vaf": rng.random(n), "unused": rng.normal(size=n), "st_mixed": rng.integers(0, 5, n).astype(float),
        })
    frame.loc[rng.random(n) < 0.1, "vaf"] = np.nan
    frame.loc[rng.random(n) < 0.05, "st_mixed"] = np.nan
    features = ["alt_reads", "x_qual_mean", "alt_allele", "ref_allele", "is_cycle_skip", "vaf", "st_mixed"]
    ref = frame[features].copy()           # the reference's preparation, with sklearn's own encoder
    for col in ref.select_dtypes(include=["object", "category"]).columns:
        ref.loc[:, col] = LabelEncoder().fit_transform(ref[col].astype(str))
    ref = ref.fillna(0).infer_objects(copy=False)
    x = ref.to_numpy(dtype=np.float32)
    y = (x[:, 0] + 3 * x[:, 2] - 10 * x[:, 5] + rng.normal(0, 2, n) > 12).astype(int)
    doc = XP.sklearn_gb_to_xgb_json(util.fit_model("gb_small", x.astype(np.float64), y))
    doc["learner"]["feature_names"] = features
    want = XP.predict_proba(doc, x)[:, 1]
    path = str(tmp_path / "model.json")
    with open(path, "w") as fh:
        json.dump(doc, fh)
    got = MA.predict_record_with_xgb(frame, path)
    assert got.dtype == np.float32 and got.shape == (n,)
    np.testing.assert_allclose(got, want, atol=2e-6, rtol=0)
    assert np.array_equal(MA.predict_record_with_xgb(frame.iloc[:0], doc), np.zeros(0, dtype=np.float32))
    sub = frame.iloc[:50]                  # the codes depend on the frame: a subset is encoded on its own strings
    ref_sub = sub[features].copy()
    for col in ref_sub.select_dtypes(include=["object", "category"]).columns:
        ref_sub.loc[:, col] = LabelEncoder().fit_transform(ref_sub[col].astype(str))
    want_sub = XP.predict_proba(doc, ref_sub.fillna(0).infer_objects(copy=False).to_numpy(dtype=np.float32))[:, 1]
    np.testing.assert_allclose(MA.predict_record_with_xgb(sub, doc), want_sub, atol=2e-6, rtol=0)
    with pytest.raises(ValueError, match="feature_names"):
        MA.predict_record_with_xgb(frame, {"learner": {k: v for k, v in doc["learner"].items() if k != "feature_names"}})
