"""GPU parity: the CUDA path through the C ABI vs the oracle (reference restatement) on the
same seeded VCF text.  FILTER decision bit-identical, features bit-identical (fp32), scores
within 1e-5 (BASELINE.json north_star tolerance)."""
import os

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


def test_device_sigmoid_equals_the_restatement(gpu_ctx):
    """K3's fp32 sigmoid (exponential rounded once from fp64) == the restatement's, bit for bit, on a margin sweep."""
    from oracle import xgb_predictor as XP

    rng = np.random.default_rng(5)
    m = np.concatenate([rng.normal(scale=3.0, size=200_000), np.linspace(-20, 20, 4001)]).astype(np.float32)
    p1, e = np.empty_like(m), np.empty_like(m)
    rc = gpu_ctx.lib.ugvc_test_device_sigmoid(gpu_ctx.h, lib._ptr(m), m.size, lib._ptr(p1), lib._ptr(e))  # noqa: SLF001
    if rc != 0 and os.environ.get("UGVC_LIB_PATH"):
        pytest.skip("a device-code hook: not part of the host emulation")
    assert rc == 0
    want_e = XP._expf(-m)  # noqa: SLF001
    bad_e = np.flatnonzero(e != want_e)
    want_p = (np.float32(1.0) / (np.float32(1.0) + want_e)).astype(np.float32)
    bad_p = np.flatnonzero(p1 != want_p)
    assert bad_e.size == 0 and bad_p.size == 0, (
        f"exp differs at {bad_e.size} margins (first {[(float(m[i]), e[i].view(np.uint32), want_e[i].view(np.uint32)) for i in bad_e[:3]]}), "
        f"sigmoid at {bad_p.size} (first {[(float(m[i]), p1[i].view(np.uint32), want_p[i].view(np.uint32)) for i in bad_p[:3]]})")


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
    bad = np.argwhere(res["probs"] != want)
    if bad.size:  # diagnostics: are the features the same, and what does the restatement give on the device's features?
        feats = gpu_ctx.debug_features(res["n_records"]).T
        same_feats = bool(np.array_equal(feats, ds["x"].astype(np.float32)))
        again = XP.predict_proba(doc, feats)
        print(f"features identical: {same_feats}; restatement on device features == device probs: "
              f"{bool(np.array_equal(again, res['probs']))}; margins (logit of p1) device vs oracle at the first bad rows: "
              f"{[(float(np.log(res['probs'][b[0], 1] / res['probs'][b[0], 0])), float(np.log(want[b[0], 1] / want[b[0], 0]))) for b in bad[:3]]}")
    assert bad.size == 0, (  # bit for bit
        f"fp32 probabilities differ from the xgboost restatement at {len(bad)} places, first {bad[:4].tolist()}: "
        f"{[(res['probs'][tuple(b)].view(np.uint32), want[tuple(b)].view(np.uint32)) for b in bad[:4]]}")
    _, quals, _ = R.score_math(want)
    np.testing.assert_allclose(res["qual"], quals, atol=2e-4, rtol=0)
    # FILTER bit-identical on every record: both sides take the fp32 sigmoid's exponential correctly rounded
    assert np.array_equal(res["low_score"].astype(bool), quals <= 30.0)
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


def test_cfg5_precision_recall_identical_to_oracle(gpu_ctx, ds):
    """BASELINE.json configs[4]: accuracy of the filtered call set against the synthetic truth.  With
    the FILTER column identical, the concordance counts (tp / fp / fn after filtering) and hence
    precision / recall are identical; this pins that end of the contract on the truth labels the
    generator emits."""
    model = util.fit_model("gb_small", ds["x"], ds["labels"])
    plan = MC.compile_plan(VcfHeader(ds["header_text"]), ds["tr"], model, ds["customs"])
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.set_key_order(*lib.learn_key_order(ds["text"]))
    gpu_ctx.reserve(len(ds["text"]) + 1024, len(ds["lines"]) + 16, 1)
    res = gpu_ctx.filter_batch(ds["text"], 30.0)
    exp = R.filter_variants(ds["vf"], model, ds["tr"], custom_annotations=ds["customs"])
    truth = ds["labels"].astype(bool)

    def stats(kept):
        tp, fp, fn = int((kept & truth).sum()), int((kept & ~truth).sum()), int((~kept & truth).sum())
        return tp, fp, fn, round(tp / max(1, tp + fp), 5), round(tp / max(1, tp + fn), 5)

    gpu = stats(~res["low_score"].astype(bool))
    ora = stats(np.array(["LOW_SCORE" not in f.split(";") for f in exp["filters"]]))
    assert gpu == ora and gpu[3] > 0.5 and gpu[4] > 0.5


def test_full_size_properties_batching_invariance(gpu_ctx):
    """Size-independent properties at a BASELINE-scale batch (cfg 2: logistic regression, 2 M device
    generated records): per-record results do not depend on how the text is cut into batches, the
    pass/fail counters add up, and a second pass reproduces the first bit for bit."""
    import torch

    n, n_custom = 2_000_000, 0
    small = util.make_dataset(n_records=3000, n_custom=n_custom, seed=1984)
    _, tr, x = util.fit_transformer(small)
    model = util.fit_model("lr", x, small["labels"])
    plan = MC.compile_plan(VcfHeader(lib.synth_header(n_custom)), tr, model, [])
    gpu_ctx.load_plan(plan.blob)
    cap = n + 1024
    gpu_ctx.reserve(900 << 20, cap, 1)
    buf = torch.empty(900 << 20, dtype=torch.uint8, device="cuda")
    nbytes = gpu_ctx.synth_device(7, 1_000_000, n, 50_000_000, n_custom, buf.data_ptr(), buf.numel() - 64)
    head = bytes(buf[: 1 << 20].cpu().numpy())
    gpu_ctx.set_key_order(*lib.learn_key_order(head[: head.rfind(b"\n") + 1]))
    low = torch.empty(cap, dtype=torch.uint8, device="cuda")
    probs = torch.empty((cap, 2), dtype=torch.float32, device="cuda")
    qual = torch.empty(cap, dtype=torch.float64, device="cuda")
    ls = torch.empty(cap + 1, dtype=torch.int64, device="cuda")
    nrec = torch.zeros(1, dtype=torch.int64, device="cuda")

    def run(off, nb, rec0):
        gpu_ctx.filter_device(buf.data_ptr() + off, nb, 30.0, low.data_ptr() + rec0, probs.data_ptr() + rec0 * 8,
                              qual.data_ptr() + rec0 * 8, cap - rec0, d_n_records=nrec.data_ptr(),
                              d_line_start=ls.data_ptr())
        gpu_ctx.device_status()
        return int(nrec.item())

    gpu_ctx.counts_reset()
    assert run(0, nbytes, 0) == n
    c1 = gpu_ctx.counts()
    starts = ls[: n + 1].clone()
    ref_low, ref_qual, ref_probs = low[:n].clone(), qual[:n].clone(), probs[:n].clone()
    assert c1["n_records"] == n and c1["n_low_score"] + c1["n_pass"] == n
    assert c1["n_low_score"] == int(ref_low.sum().item()) and 0 < c1["n_low_score"] < n
    # second pass: bit-identical
    low.zero_(); qual.zero_(); probs.zero_()  # noqa: E702
    run(0, nbytes, 0)
    assert torch.equal(low[:n], ref_low) and torch.equal(qual[:n], ref_qual) and torch.equal(probs[:n], ref_probs)
    # seven uneven batches cut at line starts (16-byte alignment is part of the device-text contract)
    cuts = [0]
    host_starts = starts.cpu().numpy()
    for frac in (0.07, 0.2, 0.21, 0.5, 0.77, 0.9):
        r = int(n * frac)
        while host_starts[r] % 16:
            r += 1
        cuts.append(r)
    cuts.append(n)
    low.zero_(); qual.zero_(); probs.zero_()  # noqa: E702
    gpu_ctx.counts_reset()
    for a, b in zip(cuts[:-1], cuts[1:]):
        off, end = int(starts[a].item()), int(starts[b].item())
        assert run(off, end - off, a) == b - a
    assert torch.equal(low[:n], ref_low) and torch.equal(qual[:n], ref_qual) and torch.equal(probs[:n], ref_probs)
    assert gpu_ctx.counts() == c1
