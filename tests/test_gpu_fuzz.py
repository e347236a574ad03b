"""GPU: differential fuzz of K1/K2 against the oracle loader + transformer.  Records are rewritten at
random -- INFO keys shuffled, unknown keys and flags inserted, optional keys dropped, numbers
re-spelled (exponents, signs, leading zeros, long fractions), FORMAT columns permuted -- so the
schedule-driven path, its fall-back to the generic key lookup and the exact number parser all
see inputs that no generator-shaped test gives them.  Features must stay bit-identical."""
import numpy as np
import pandas as pd
import pytest

from oracle import ref_pipeline as R
from oracle.vcf_reader import OracleVariantFile
from tests import util
from variantcalling_b200 import lib
from variantcalling_b200 import model_compiler as MC
from variantcalling_b200.vcf_header import VcfHeader

pytestmark = pytest.mark.gpu

DROPPABLE = {"SOR", "FS", "QD", "MQ", "AN", "ExcessHet", "XC", "X_GCC", "DP", "AF", "AC", "MLEAC", "MLEAF", "HAPCOMP",
             "VARIANT_TYPE", "X_HIL", "X_IL", "X_HIN"}
FLOAT_KEYS = {"SOR", "FS", "QD", "MQ", "X_GCC", "BaseQRankSum", "MQRankSum", "ReadPosRankSum", "ExcessHet"}


def respell(rng, text):
    """Another spelling of the same decimal value (same float32 after strtod)."""
    try:
        v = float(text)
    except ValueError:
        return text
    r = rng.integers(0, 6)
    if r == 0:
        return "%.6e" % v if float(np.float32(float("%.6e" % v))) == float(np.float32(v)) else text
    if r == 1 and not text.startswith(("-", "+")):
        return "+" + text
    if r == 2 and not text.startswith(("-", "+")):
        return "00" + text
    if r == 3 and "." in text and "e" not in text.lower():
        return text + "0000000000000"
    if r == 4 and "." in text and "e" not in text.lower():
        return text + "e0"
    return text


def mutate(rng, line):
    c = line.split("\t")
    info = c[7].split(";")
    out = []
    for kv in info:
        k, _, v = kv.partition("=")
        if k in DROPPABLE and rng.random() < 0.08:
            continue
        if k in FLOAT_KEYS and rng.random() < 0.5:
            kv = k + "=" + respell(rng, v)
        out.append(kv)
        if rng.random() < 0.05:
            out.append(["FOO=1,2,x", "DB", "LONGKEYNAME_THAT_IS_NOT_A_TAG_ANYWHERE=3", "X_LMX=ACGT", "A=", "DPX=7"][rng.integers(0, 6)])
    if rng.random() < 0.4:
        rng.shuffle(out)
    c[7] = ";".join(out) if out else "."
    if rng.random() < 0.3:  # permute FORMAT sub-fields (keys and values together)
        keys, vals = c[8].split(":"), c[9].split(":")
        order = rng.permutation(len(keys))
        c[8], c[9] = ":".join(keys[i] for i in order), ":".join(vals[i] for i in order)
    if rng.random() < 0.3:
        c[5] = respell(rng, c[5])
    return "\t".join(c)


@pytest.mark.parametrize("seed", [1, 2])
def test_fuzzed_records_features_bit_identical(gpu_ctx, seed):
    rng = np.random.default_rng(seed)
    ds = util.make_dataset(n_records=5000, n_custom=8, seed=40 + seed)
    _, tr, x = util.fit_transformer(ds)
    model = util.fit_model("lr", x, ds["labels"])
    lines = [mutate(rng, ln) for ln in ds["lines"]]
    text = ("\n".join(lines) + "\n").encode()
    vf = OracleVariantFile(ds["header_text"].encode() + text)
    df = R.harness_float_columns(R.get_vcf_df(vf, None, ds["customs"]))
    with pd.option_context("future.infer_string", False):
        want = tr.transform(df).to_numpy(dtype=np.float64).astype(np.float32)
    plan = MC.compile_plan(VcfHeader(ds["header_text"]), tr, model, ds["customs"])
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.reserve(len(text) + 64, len(lines) + 8, 1)
    for mode in ("generic", "learned", "learned-from-unmutated"):
        if mode == "learned":
            gpu_ctx.set_key_order(*lib.learn_key_order(text))
        elif mode == "learned-from-unmutated":
            gpu_ctx.set_key_order(*lib.learn_key_order(ds["text"]))
        res = gpu_ctx.filter_batch(text)
        got = gpu_ctx.debug_features(res["n_records"]).T
        bad = np.argwhere(got != want)
        assert bad.size == 0, (f"[{mode}] {len(bad)} mismatches; first at record {bad[0][0]} column {bad[0][1]} "
                               f"({plan.feature_names[bad[0][1]]}): {got[tuple(bad[0])]} vs {want[tuple(bad[0])]}\n"
                               f"{lines[bad[0][0]]}")
