"""CPU: the UBJSON reader (xgboost's binary model flavour) against an independent encoder written
from the specification, and the xgboost-format lowering fed with the binary document."""
import struct

import numpy as np
import pytest

from variantcalling_b200 import model_compiler as MC
from variantcalling_b200 import ubjson


def enc_len(n):
    return (b"U" + struct.pack(">B", n)) if n < 256 else (b"l" + struct.pack(">i", n)) if n < 2**31 else (b"L" + struct.pack(">q", n))


def enc(v, typed_arrays=True):  # noqa: C901, PLR0911
    if v is None:
        return b"Z"
    if v is True:
        return b"T"
    if v is False:
        return b"F"
    if isinstance(v, int):
        for m, fmt, lo, hi in (("i", ">b", -128, 127), ("U", ">B", 0, 255), ("I", ">h", -2**15, 2**15 - 1), ("l", ">i", -2**31, 2**31 - 1)):
            if lo <= v <= hi:
                return m.encode() + struct.pack(fmt, v)
        return b"L" + struct.pack(">q", v)
    if isinstance(v, float):
        return b"D" + struct.pack(">d", v)
    if isinstance(v, str):
        raw = v.encode()
        return b"S" + enc_len(len(raw)) + raw
    if isinstance(v, (list, tuple)):
        if typed_arrays and v and all(isinstance(x, float) for x in v):
            return b"[$d#" + b"L" + struct.pack(">q", len(v)) + np.asarray(v, dtype=">f4").tobytes()
        if typed_arrays and v and all(isinstance(x, int) and not isinstance(x, bool) for x in v):
            return b"[$l#" + enc_len(len(v)) + np.asarray(v, dtype=">i4").tobytes()
        if typed_arrays:
            return b"[#" + enc_len(len(v)) + b"".join(enc(x, typed_arrays) for x in v)
        return b"[" + b"".join(enc(x, typed_arrays) for x in v) + b"]"
    if isinstance(v, dict):
        body = b"".join(enc_len(len(k.encode())) + k.encode() + enc(x, typed_arrays) for k, x in v.items())
        return (b"{#" + enc_len(len(v)) + body) if typed_arrays else (b"{" + body + b"}")
    raise TypeError(type(v))


@pytest.mark.parametrize("typed", [True, False])
def test_round_trip_of_every_value_kind(typed):
    doc = {"a": [1, 2, 300, -70000], "b": [0.5, -1.25, 3.0], "c": {"x": None, "y": True, "z": False, "s": "héllo", "n": 2**40},
           "d": [], "e": [[1, 2], {"k": "v"}, "t", 1.5], "f": 3.141592653589793, "long": "x" * 300}
    got = ubjson.loads(enc(doc, typed))
    assert got == doc
    assert ubjson.loads(b"N" + enc({"k": 1})) == {"k": 1}                       # no-op bytes are skipped
    assert ubjson.loads(b"[$U#U\x03\x01\x02\x03") == [1, 2, 3] and ubjson.loads(b"C" + b"q") == "q"
    assert ubjson.loads(b"H" + enc_len(4) + b"12.5") == 12.5
    with pytest.raises(ubjson.UbjsonError):
        ubjson.loads(enc(doc, typed)[:-3])
    with pytest.raises(ubjson.UbjsonError):
        ubjson.loads(b"?")


def test_binary_xgboost_document_lowers_like_the_json_one():
    import json

    from oracle import xgb_predictor as XP
    from tests import util

    ds = util.make_dataset(n_records=1500, n_custom=2, seed=8)
    _, _tr, x = util.fit_transformer(ds)
    doc = XP.sklearn_gb_to_xgb_json(util.fit_model("gb_small", x, ds["labels"]))
    doc = json.loads(json.dumps(doc))  # plain python containers
    as_json = MC.compile_plan_model_only(json.dumps(doc).encode(), x.shape[1])
    as_ubj = MC.compile_plan_model_only(enc(doc), x.shape[1])
    assert as_ubj.blob == as_json.blob and as_ubj.model_kind == MC.MODEL_XGB


def test_model_files_of_either_flavour_load_as_the_same_document(tmp_path):
    """model_apply.load_xgb_document: what XGBClassifier.load_model reads -- a .json or a .ubj file, raw bytes of either,
    or the parsed document (the model step of the featuremap tool, featuremap_xgb_prediction.py:301-323)."""
    import json

    from variantcalling_b200 import model_apply as MA

    doc = {"learner": {"feature_names": ["a", "b"], "objective": {"name": "binary:logistic"},
                       "learner_model_param": {"base_score": "5E-1", "num_class": "0", "num_feature": "2"},
                       "gradient_booster": {"name": "gbtree", "model": {"trees": [
                           {"left_children": [1, -1, -1], "right_children": [2, -1, -1], "split_indices": [0, 0, 0],
                            "split_conditions": [0.5, -0.25, 0.75], "default_left": [0, 0, 0]}], "tree_info": [0]}}}}
    pj, pu = tmp_path / "m.json", tmp_path / "m.ubj"
    pj.write_text(json.dumps(doc))
    pu.write_bytes(enc(doc))
    for source in (str(pj), str(pu), pj.read_bytes(), pu.read_bytes(), doc):
        got = MA.load_xgb_document(source)
        assert got["learner"]["feature_names"] == ["a", "b"]
        assert [float(v) for v in got["learner"]["gradient_booster"]["model"]["trees"][0]["split_conditions"]] == [0.5, -0.25, 0.75]
