"""GPU: edge cases of the typed decode and of the transformer policies, against the golden
vectors of the reference's own transformer and against the oracle's error behaviour."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import ref_pipeline as R
from oracle.vcf_reader import OracleVariantFile
from tests import util
from variantcalling_b200 import lib
from variantcalling_b200 import model_compiler as MC
from variantcalling_b200 import transformers as T
from variantcalling_b200.tprep_constants import VcfType
from variantcalling_b200.vcf_header import VcfHeader

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def split_text(text: bytes):
    lines = text.decode().split("\n")
    hdr = [l for l in lines if l.startswith("#")]
    recs = [l for l in lines if l and not l.startswith("#")]
    return "\n".join(hdr) + "\n", ("\n".join(recs) + "\n").encode(), recs


@pytest.fixture(scope="module")
def golden():
    z = np.load(os.path.join(GOLD, "transformer_single_sample.npz"))
    text, customs = bytes(z["vcf_text"]), [str(c) for c in z["customs"]]
    vf = OracleVariantFile(text)
    df = R.harness_float_columns(R.get_vcf_df(vf, None, customs))
    tr = T.get_transformer(VcfType.SINGLE_SAMPLE, [c.lower() for c in customs])
    with pd.option_context("future.infer_string", False):
        x = tr.fit_transform(df).to_numpy(dtype=np.float64)
    rng = np.random.default_rng(3)
    y = (x[:, 23] + rng.normal(size=x.shape[0]) * 5 > 15).astype(int)
    return dict(text=text, customs=customs, feats_ref=z["features_ref"], vf=vf, tr=tr, x=x, y=y)


def test_features_equal_reference_transformer_golden(gpu_ctx, golden):
    """K1+K2 output == ugbio_filtering.transformers' own fit_transform (fp32 cast), incl. the
    hand-written edge rows (float32 rounding, '.' values, FORMAT-overrides-INFO, short PL, dropped
    trailing FORMAT sub-fields, N / lower-case motifs)."""
    model = util.fit_model("lr", golden["x"], golden["y"])
    hdr, body, recs = split_text(golden["text"])
    plan = MC.compile_plan(VcfHeader(hdr), golden["tr"], model, golden["customs"])
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.reserve(len(body) + 64, len(recs) + 8, 1)
    want = golden["feats_ref"].astype(np.float32)
    for mode in ("generic", "learned"):
        if mode == "learned":
            gpu_ctx.set_key_order(*lib.learn_key_order(body))
        res = gpu_ctx.filter_batch(body)
        feats = gpu_ctx.debug_features(res["n_records"]).T
        bad = np.argwhere(feats != want)
        assert bad.size == 0, f"[{mode}] {len(bad)} mismatches, first {bad[0]}: {feats[tuple(bad[0])]} vs {want[tuple(bad[0])]}"
    exp = R.filter_variants(golden["vf"], model, golden["tr"], custom_annotations=golden["customs"])
    assert np.array_equal(res["low_score"].astype(bool), np.array(["LOW_SCORE" in f.split(";") for f in exp["filters"]]))
    np.testing.assert_allclose(res["probs"], exp["probs"], atol=1e-5, rtol=0)


BAD_ROWS = {
    "af_missing_element": ("AF=0.500", "AF=."),                  # (None,) -> NaN -> _validate_data asserts
    "unknown_css_category": ("X_CSS=non-skip", "X_CSS=weird"),    # OrdinalEncoder: unknown category
    "unknown_indel_class": ("X_IC=NA", "X_IC=dup"),               # KeyError in ins_del_encode
    "mq0c_single_value": ("MQ0C=0,0", "MQ0C=0"),                  # ragged doublet
    "custom_unknown_value": ("XC=3", "XC=3;LCR=MAYBE"),           # OrdinalEncoder: unknown category
    "infinite_imputed_value": ("SOR=0.693", "SOR=1e39"),          # float32 overflow -> inf: SimpleImputer's input check
    "infinite_literal": ("MQ=60.00", "MQ=-inf"),                  # ValueError: Input X contains infinity
}


@pytest.mark.parametrize("case", sorted(BAD_ROWS) + ["pl_too_wide", "alt_missing", "qual_missing", "qual_infinite"])
def test_inputs_the_reference_raises_on_raise_here_too(gpu_ctx, golden, case):
    model = util.fit_model("lr", golden["x"], golden["y"])
    hdr, body, recs = split_text(golden["text"])
    row = recs[-8]  # first hand-written edge row (well formed)
    if case in BAD_ROWS:
        a, b = BAD_ROWS[case]
        assert a in row
        bad = row.replace(a, b, 1)
    elif case == "pl_too_wide":
        bad = row.rsplit(":", 1)[0] + ":100,0,200,300"
    elif case == "qual_infinite":  # passthrough column: the model's own input check raises on inf
        cols = row.split("\t"); cols[5] = "1e40"; bad = "\t".join(cols)  # noqa: E702
    elif case == "alt_missing":
        cols = row.split("\t"); cols[4] = "."; bad = "\t".join(cols)  # noqa: E702
    else:
        cols = row.split("\t"); cols[5] = "."; bad = "\t".join(cols)  # noqa: E702
    new_recs = recs[:50] + [bad] + recs[50:100]
    text = (hdr + "\n".join(new_recs) + "\n").encode()
    with pytest.raises(Exception):  # noqa: B017  the reference path raises (ValueError/KeyError/AssertionError)
        R.filter_variants(OracleVariantFile(text), model, golden["tr"], custom_annotations=golden["customs"])
    plan = MC.compile_plan(VcfHeader(hdr), golden["tr"], model, golden["customs"])
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.set_key_order(*lib.learn_key_order(("\n".join(recs) + "\n").encode()))
    gpu_ctx.reserve(1 << 20, 4096, 1)
    with pytest.raises(lib.UgvcDataError):
        gpu_ctx.filter_batch(("\n".join(new_recs) + "\n").encode())
    rec, _col, reason = gpu_ctx.last_data_error()
    assert rec == 50 and reason != 0
    # and the same batch without the bad row is fine
    ok = gpu_ctx.filter_batch(("\n".join(recs[:100]) + "\n").encode())
    assert ok["n_records"] == 100


def test_empty_batch_and_batch_contract(gpu_ctx, golden):
    model = util.fit_model("lr", golden["x"], golden["y"])
    hdr, body, recs = split_text(golden["text"])
    gpu_ctx.load_plan(MC.compile_plan(VcfHeader(hdr), golden["tr"], model, golden["customs"]).blob)
    gpu_ctx.reserve(1 << 20, 4096, 2)
    assert gpu_ctx.filter_batch(b"")["n_records"] == 0
    with pytest.raises(lib.UgvcError):
        gpu_ctx.filter_batch(recs[0].encode())  # no trailing newline
    one = gpu_ctx.filter_batch((recs[0] + "\n").encode())
    assert one["n_records"] == 1
    with pytest.raises(lib.UgvcError):
        gpu_ctx.filter_batch(b"x" * ((1 << 20) + 100) + b"\n")  # larger than reserved
    # pipelined lanes give the same answer as the blocking call
    full = gpu_ctx.filter_batch(body)
    n = full["n_records"]
    half = body[: full["line_start"][n // 2]]
    rest = body[full["line_start"][n // 2]:]
    a, b = np.frombuffer(half, np.uint8), np.frombuffer(rest, np.uint8)
    gpu_ctx.submit(0, a, a.size)
    gpu_ctx.submit(1, b, b.size)
    o0, o1 = gpu_ctx.alloc_outputs(n), gpu_ctx.alloc_outputs(n)
    n0 = gpu_ctx.collect(0, o0, n)
    n1 = gpu_ctx.collect(1, o1, n)
    assert n0 + n1 == n
    assert np.array_equal(np.concatenate((o0["qual"][:n0], o1["qual"][:n1])), full["qual"])
    assert np.array_equal(np.concatenate((o0["low_score"][:n0], o1["low_score"][:n1])), full["low_score"])


def test_float_literals_match_strtod_float32(gpu_ctx, golden):
    """QUAL is a passthrough feature: K1's parse must equal float32(strtod(text)) bit for bit."""
    model = util.fit_model("lr", golden["x"], golden["y"])
    hdr, body, recs = split_text(golden["text"])
    rng = np.random.default_rng(12)
    lits = ["0", "0.0", "1", "16777217", "16777216", "16777215.5", "0.1", "1e5", "1E-3", "1.5e+2", "+2.5", "7.",
            "000123.4500", "33.333333333", "1e22", "123456789012345", "0.000001234", "9007199254740992",
            "3.4028234e38", "1e-5", "4.35", "2.675", "0.30000000000000004", "1.0000001192092896",
            "8388608.5", "8388609.5", "0.1234567890123456789", "123456789.123456789e-5", "1e-30", "1e-45",
            "2.2250738585072014e-308", "17976931348623157e292", "1.17549435e-38", "7.00649232e-46"]
    for _ in range(3000):
        digits = rng.integers(1, 18)
        m = "".join(str(d) for d in rng.integers(0, 10, size=digits))
        point = rng.integers(0, digits + 1)
        s = (m[:point] or "0") + ("." + m[point:] if point < digits else "")
        if rng.random() < 0.3:
            s += "e" + str(int(rng.integers(-30, 20)))
        lits.append(s)
    row = recs[-8].split("\t")
    new = []
    for s in lits:
        r = list(row)
        r[5] = s
        new.append("\t".join(r))
    gpu_ctx.load_plan(MC.compile_plan(VcfHeader(hdr), golden["tr"], model, golden["customs"]).blob)
    gpu_ctx.reserve(4 << 20, 8192, 1)
    with np.errstate(over="ignore"):
        want = np.array([np.float32(float(s)) for s in lits], dtype=np.float32)
    # a literal beyond float32 parses to inf, which the batch then refuses like the reference's input checks do
    # ("Input X contains infinity"); the feature matrix is still there to compare the parse itself
    with pytest.raises(lib.UgvcDataError):
        gpu_ctx.filter_batch(("\n".join(new) + "\n").encode())
    first_inf = int(np.flatnonzero(np.isinf(want))[0])
    assert gpu_ctx.last_data_error()[0] == first_inf and lits[first_inf] == "17976931348623157e292"
    qual = gpu_ctx.debug_features(len(lits))[21]
    bad = np.flatnonzero(qual.view(np.uint32) != want.view(np.uint32))
    assert bad.size == 0, f"{bad.size} literals differ, e.g. {lits[bad[0]]!r}: {qual[bad[0]]!r} vs {want[bad[0]]!r}"


def test_device_generator_text_parity(gpu_ctx):
    """The CUDA generator's records (bench input) through the oracle vs the GPU path."""
    import torch

    n, n_custom = 6000, 40
    ds = util.make_dataset(n_records=3000, n_custom=n_custom, seed=1984)
    _, tr, x = util.fit_transformer(ds)
    model = util.fit_model("gb_small", x, ds["labels"])
    hdr = lib.synth_header(n_custom)
    plan = MC.compile_plan(VcfHeader(hdr), tr, model, ds["customs"])
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.reserve(8 << 20, n + 128, 1)
    buf = torch.empty(8 << 20, dtype=torch.uint8, device="cuda")
    nbytes = gpu_ctx.synth_device(20260922, 49_994_000, n, 50_000_000, n_custom, buf.data_ptr(), buf.numel() - 64)
    text = bytes(buf[:nbytes].cpu().numpy())
    assert text.count(b"\n") == n and text.startswith(b"chrY\t")
    res = gpu_ctx.filter_batch(text)
    vf = OracleVariantFile(hdr.encode() + text)
    exp = R.filter_variants(vf, model, tr, custom_annotations=ds["customs"])
    assert np.array_equal(gpu_ctx.debug_features(n).T, exp["features"].astype(np.float32))
    assert np.array_equal(res["low_score"].astype(bool), np.array(["LOW_SCORE" in f.split(";") for f in exp["filters"]]))
    np.testing.assert_allclose(res["probs"], exp["probs"], atol=1e-5, rtol=0)
    pos = res["recinfo"]["pos"]
    assert np.all(np.diff(pos) >= 0), "generator positions are sorted within a contig"
    # the first records of the job sit on chr1
    nb2 = gpu_ctx.synth_device(20260922, 0, 10, 50_000_000, n_custom, buf.data_ptr(), buf.numel() - 64)
    assert bytes(buf[:nb2].cpu().numpy()).startswith(b"chr1\t")
