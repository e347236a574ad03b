"""The model-apply step on its own: ``predict_proba`` / ``predict`` of a fitted model on the GPU (K3).

``apply_model(input_df, model, transformer)`` mirrors
``ugbio_filtering/variant_filtering_utils.py:95-125`` (transform in chunks on the host, null check,
then the model -- here one K3 launch per chunk instead of ``model.predict`` + ``model.predict_proba``),
and :class:`GpuClassifier` gives the same step to the other model-apply tools the reference has
(``featuremap_xgb_prediction.predict_record_with_xgb``, ``srsnv_inference_utils``: a feature frame and
an xgboost / sklearn classifier).  Supported models: what ``model_compiler`` lowers (sklearn
LogisticRegression / GradientBoostingClassifier / RandomForestClassifier, xgboost JSON or Booster).
Probabilities come back in the trainer's precision (fp64 for sklearn, fp32 values for xgboost).
``predict_record_with_xgb`` is the model step of the featuremap tool
(``ugbio_featuremap/featuremap_xgb_prediction.py:258-262,301-323``) with the same signature: the frame its
``df_vcf_manual_aggregation`` built, the model file, the probability of class 1 back.
There is no CPU path.
"""
from __future__ import annotations

import json

import numpy as np
import pandas as pd

from variantcalling_b200 import lib
from variantcalling_b200 import model_compiler as MC

MAX_CHUNK_SIZE = 1000000  # variant_filtering_utils.py:18


class GpuClassifier:
    def __init__(self, model, n_features: int | None = None, device: int = 0, max_rows: int = 1 << 20):
        if n_features is None:
            n_features = getattr(model, "n_features_in_", None)
        if n_features is None:
            raise ValueError("n_features is needed for a model that does not record n_features_in_")
        self.plan = MC.compile_plan_model_only(model, int(n_features))
        self.classes_ = np.asarray(self.plan.classes if len(self.plan.classes) else np.arange(self.plan.n_classes))
        self.ctx = lib.Context(device)
        self.ctx.load_plan(self.plan.blob)
        self.ctx.enable_phreds(2)  # K3 keeps the fp64 class probabilities
        self.max_rows = 0
        self._reserve(max_rows)

    def _reserve(self, rows: int):
        if rows > self.max_rows:
            self.max_rows = int(rows)
            self.ctx.reserve(4096, self.max_rows, 1)

    def close(self):
        self.ctx.close()

    def predict_proba(self, x) -> np.ndarray:
        x = x.to_numpy() if hasattr(x, "to_numpy") else np.asarray(x)
        if x.ndim != 2 or x.shape[1] != self.plan.n_features:
            raise ValueError(f"X has {x.shape[1] if x.ndim == 2 else '?'} features, the model expects {self.plan.n_features}")
        if not np.isfinite(np.asarray(x, dtype=np.float64)).all():
            raise ValueError("Input X contains NaN or infinity")  # sklearn's check_array message, in short
        out = np.empty((x.shape[0], self.plan.n_classes), dtype=np.float64)
        step = min(self.max_rows, MAX_CHUNK_SIZE)
        for b in range(0, x.shape[0], step):
            part = np.ascontiguousarray(x[b:b + step], dtype=np.float32)
            self.ctx.predict_features(part)
            out[b:b + part.shape[0]] = self.ctx.collect_phreds(0, part.shape[0])
        return out

    def predict(self, x) -> np.ndarray:
        return self.classes_[np.argmax(self.predict_proba(x), axis=1)]


def _validate_data(data) -> None:
    """variant_filtering_utils.py:128-143"""
    arr = data if isinstance(data, np.ndarray) else pd.DataFrame(data).to_numpy()
    if arr.ndim == 1 or arr.shape[1] <= 1:
        assert pd.isna(arr).sum() == 0, "data vector contains null"  # noqa: S101
    else:
        for c in range(arr.shape[1]):
            assert pd.isna(arr[:, c]).sum() == 0, f"Data matrix contains null in column {c}"  # noqa: S101


def apply_model(input_df: pd.DataFrame, model, transformer, classifier: GpuClassifier | None = None):
    """-> (predictions, probabilities) like variant_filtering_utils.apply_model (:95-125)."""
    chunks = np.arange(0, input_df.shape[0], MAX_CHUNK_SIZE, dtype=int)
    chunks = np.concatenate((chunks, [input_df.shape[0]]))
    parts = [transformer.transform(input_df.iloc[chunks[i]:chunks[i + 1]]) for i in range(len(chunks) - 1)]
    x_test = pd.concat(parts)
    _validate_data(x_test)
    own = classifier is None
    clf = classifier or GpuClassifier(model, n_features=x_test.shape[1], max_rows=min(MAX_CHUNK_SIZE, max(1, x_test.shape[0])))
    try:
        probabilities = clf.predict_proba(x_test)
        predictions = clf.classes_[np.argmax(probabilities, axis=1)]
    finally:
        if own:
            clf.close()
    return predictions, probabilities


# ------------------------------------------------------------------ featuremap_xgb_prediction.py:258-262,301-323
def set_categorial_columns(df: pd.DataFrame) -> None:
    """In place, like the reference's helper: every object / category column becomes the rank of ``str(value)`` among
    the sorted distinct strings of THIS frame (``LabelEncoder().fit_transform(df[col].astype(str))``)."""
    for col in df.select_dtypes(include=["object", "category", "string"]).columns:
        _, codes = np.unique(df[col].astype(str).to_numpy(dtype=object).astype(str), return_inverse=True)
        df[col] = codes.astype(np.int64)


def load_xgb_document(xgb_model) -> dict:
    """The model file of ``XGBClassifier.load_model`` (JSON or UBJSON), a parsed document, or raw bytes of either."""
    if isinstance(xgb_model, dict):
        return xgb_model
    raw = xgb_model
    if isinstance(xgb_model, str):
        with open(xgb_model, "rb") as fh:
            raw = fh.read()
    try:
        return json.loads(raw)
    except (UnicodeDecodeError, json.JSONDecodeError):
        from variantcalling_b200 import ubjson

        return ubjson.loads(bytes(raw))


def predict_record_with_xgb(df_variants: pd.DataFrame, xgb_model, classifier: GpuClassifier | None = None) -> np.ndarray:
    """-> probability of class "1" for every row (the reference's ``df_probabilities["1"].to_numpy()``).
    Columns = the booster's ``feature_names``; object columns label-encoded per call, nulls filled with 0; then K3."""
    doc = load_xgb_document(xgb_model)
    features = doc["learner"].get("feature_names") or []
    if not features:
        raise ValueError("the model document carries no feature_names (it was not trained on a data frame)")
    x = df_variants[list(features)].copy()
    set_categorial_columns(x)
    x = x.fillna(0)
    own = classifier is None
    clf = classifier or GpuClassifier(doc, n_features=len(features), max_rows=min(MAX_CHUNK_SIZE, max(1, x.shape[0])))
    try:
        if x.shape[0] == 0:
            return np.zeros(0, dtype=np.float32)
        probabilities = clf.predict_proba(x.to_numpy(dtype=np.float32))
    finally:
        if own:
            clf.close()
    return probabilities[:, 1].astype(np.float32)  # xgboost's probabilities are fp32 values
