"""GPU: the K1 tile kernel (k1_tok: one delimiter pass per 32 KiB tile, one lane per INFO field) against the generic
per-record parser (UGVC_K1_LEGACY=1), which defines the semantics (vcftools.py:63-89,196-214 through pysam's
typing rules).  The inputs aim at the tile bookkeeping rather than at the values: every alignment of the lines
against the 64-byte scan spans and the tile border, lines of a few dozen bytes (several windows per tile, then
more newlines than a tile holds), lines around and beyond the 4 KiB overlap, ';' outside INFO, empty fields,
extra sample columns, control characters next to tabs, and numeric tokens around the edges of the short decoder.
Slot words, recinfo rows, scores and data errors must be identical; clean data must stay off the slow list."""
import os

import numpy as np
import pytest

from tests import util
from variantcalling_b200 import lib
from variantcalling_b200 import model_compiler as MC
from variantcalling_b200.vcf_header import VcfHeader

pytestmark = pytest.mark.gpu

TILE = 32768


@pytest.fixture(scope="module")
def case():
    ds = util.make_dataset(n_records=2500, n_custom=4, seed=321)
    _, tr, x = util.fit_transformer(ds)
    model = util.fit_model("gb_small", x, ds["labels"])
    plan = MC.compile_plan(VcfHeader(ds["header_text"]), tr, model, ds["customs"])
    return ds, plan


def run(ctx, plan, text, learn_from, legacy):
    if legacy:
        os.environ["UGVC_K1_LEGACY"] = "1"
    else:
        os.environ.pop("UGVC_K1_LEGACY", None)
    try:
        ctx.load_plan(plan.blob)
        ctx.reserve(len(text) + 4096, text.count(b"\n") + 16, 1)
        ctx.set_key_order(*lib.learn_key_order(learn_from))
        try:
            out = ctx.filter_batch(text)
        except lib.UgvcDataError:
            return {"error": ctx.last_data_error()}
        n = out["n_records"]
        res = {"n": n, "raw": ctx.debug_raw(n).copy(), "recinfo": np.asarray(out["recinfo"]).copy(),
               "low": out["low_score"].copy(), "qual": out["qual"].copy(), "slow": ctx.slow_records(0)}
        return res
    finally:
        os.environ.pop("UGVC_K1_LEGACY", None)


def same(a, b, what):
    assert ("error" in a) == ("error" in b), (what, a.get("error"), b.get("error"))
    if "error" in a:
        assert a["error"] == b["error"], what
        return
    assert a["n"] == b["n"], what
    bad = np.argwhere(a["raw"] != b["raw"])
    assert bad.size == 0, f"{what}: {len(bad)} slot words differ, first (slot, record) = {bad[0]}: {a['raw'][tuple(bad[0])]:#x} vs {b['raw'][tuple(bad[0])]:#x}"
    assert a["recinfo"].tobytes() == b["recinfo"].tobytes(), what
    assert np.array_equal(a["low"], b["low"]) and np.array_equal(a["qual"], b["qual"]), what


def both(ctx, plan, text, learn_from, what):
    fast = run(ctx, plan, text, learn_from, legacy=False)
    slow = run(ctx, plan, text, learn_from, legacy=True)
    same(fast, slow, what)
    return fast


def pad_first(lines, extra):
    """The first record grows by `extra` bytes (an INFO key nobody needs), which shifts every later line."""
    c = lines[0].split("\t")
    c[7] += ";ZZPAD=" + "x" * max(0, extra - 7)
    return ["\t".join(c)] + lines[1:]


def test_clean_data_stays_on_the_tile_kernel_at_every_alignment(gpu_ctx, case):
    ds, plan = case
    lines = ds["lines"][:600]
    learn = ("\n".join(pad_first(lines, 8)) + "\n").encode()
    first = len(lines[0]) + 1
    # where the second line starts: every offset of a 64-byte span, and line borders on / next to the tile border
    shifts = list(range(8, 8 + 64)) + [TILE - first - 1 + d for d in (-1, 0, 1, 2)] + [2 * TILE - first - 1 + d for d in (0, 1)]
    for extra in shifts:
        text = ("\n".join(pad_first(lines, extra)) + "\n").encode()
        res = both(gpu_ctx, plan, text, learn, f"shift {extra}")
        if extra < 4000:
            assert res["slow"] == 0, (extra, res["slow"])
        else:
            assert res["slow"] <= 1  # only the padded record itself (longer than the overlap) may be handed over


def test_short_lines_many_windows_and_more_newlines_than_a_tile_holds(gpu_ctx, case):
    ds, plan = case
    learn = ds["text"]
    short = []
    for ln in ds["lines"][:2400]:
        c = ln.split("\t")
        info = [kv for kv in c[7].split(";") if kv.split("=")[0] in ("DP", "X_CSS", "X_IC", "X_LM", "X_RM", "MQ0C", "SCL", "SCR")]  # what the encoders insist on
        c[7] = ";".join(info) if info else "."
        short.append("\t".join(c))                       # ~110 bytes: several windows per tile
    tiny = ["\t".join(ln.split("\t")[:7] + ["X_IC=NA;X_LM=A;X_RM=C;X_CSS=non-skip;MQ0C=1,1;SCL=0,0;SCR=0,0"] + ln.split("\t")[8:])
            for ln in ds["lines"][:2400]]                    # the shortest line the encoders accept
    res = both(gpu_ctx, plan, ("\n".join(short) + "\n").encode(), learn, "short lines")
    assert res["n"] == 2400
    res = both(gpu_ctx, plan, ("\n".join(tiny) + "\n").encode(), learn, "tiny lines")
    assert res["n"] == 2400
    mixed = []
    for i in range(1200):
        mixed += [short[i], tiny[i], ds["lines"][i]]
    both(gpu_ctx, plan, ("\n".join(mixed) + "\n").encode(), learn, "mixed lines")


def test_more_newlines_than_a_tile_holds(gpu_ctx, case):
    """Lines of ~25 bytes (over a thousand per tile) under the model-less plan of the index pass: the tile kernel hands
    whole tiles to the generic parser, which finds the lines through the line starts the tile kernel wrote."""
    ds, _ = case
    plan = MC.compile_plan_no_model(VcfHeader(ds["header_text"]))
    rng = np.random.default_rng(3)
    lines = []
    for i, ln in enumerate(ds["lines"][:2400]):
        c = ln.split("\t")
        lines.append("\t".join(c[:2] + [".", c[3], c[4], "9", ".", "DP=%d" % rng.integers(1, 99)]))
        if i % 400 == 399:
            lines += ds["lines"][i - 30:i]               # a stretch of ordinary lines in between
    text = ("\n".join(lines) + "\n").encode()
    fast = run(gpu_ctx, plan, text, ds["text"], legacy=False)
    slow = run(gpu_ctx, plan, text, ds["text"], legacy=True)
    assert "error" not in slow and fast["n"] == slow["n"] == len(lines)
    assert fast["recinfo"].tobytes() == slow["recinfo"].tobytes()
    assert fast["slow"] > 1000                           # whole tiles were handed over


def test_long_lines_around_the_overlap(gpu_ctx, case):
    ds, plan = case
    lines = list(ds["lines"][:400])
    for at, size in ((3, 3800), (40, 3995), (41, 4005), (90, 4200), (150, 9000), (151, 70000), (300, 33000)):
        c = lines[at].split("\t")
        c[7] += ";ZZPAD=" + "y" * size
        lines[at] = "\t".join(c)
    both(gpu_ctx, plan, ("\n".join(lines) + "\n").encode(), ds["text"], "long lines")


def mutate_frame(rng, ln):
    c = ln.split("\t")
    r = rng.integers(0, 16)
    if r == 0:
        c[2] = "rs1;rs2"                                 # ';' in ID
    elif r == 1:
        c[6] = "LowQual;q10"                             # ';' in FILTER
    elif r == 2:
        c[7] += ";"                                      # empty last field
    elif r == 3:
        c[7] = c[7].replace(";", ";;", 1)                # empty field
    elif r == 4:
        c[7] = ";" + c[7]
    elif r == 5:
        c += [c[9], c[9]]                                # more samples
    elif r == 6:
        c[7] += ";ZZ=a\x0bb"                             # vertical tab in a value
    elif r == 7:
        c[7] += ";ZZ=a\x08b"
    elif r == 8:
        c = c[:8]                                        # no FORMAT / sample
    elif r == 9:
        c = c[:9]                                        # FORMAT without a sample
    elif r == 10:
        c[9] = ":".join(c[9].split(":")[:3])             # trailing sub-fields dropped
    elif r == 11:
        c[7] = c[7] + ";" + c[7].split(";")[0]           # a key twice
    elif r == 12:
        c[9] += ";x"                                     # ';' beyond INFO
    elif r == 13:
        c[7] = c[7].replace("=", "==", 1)
    elif r == 14:
        c = c[:5]                                        # columns missing
    return "\t".join(c)


def test_odd_frames(gpu_ctx, case):
    ds, plan = case
    rng = np.random.default_rng(5)
    for rounds in range(3):
        lines = [mutate_frame(rng, ln) if rng.random() < 0.3 else ln for ln in ds["lines"][:1500]]
        # a line the reference raises on decides on both sides (same record, same reason); drop it and go on
        for attempt in range(400):
            text = ("\n".join(lines) + "\n").encode()
            fast = run(gpu_ctx, plan, text, ds["text"], legacy=False)
            slow = run(gpu_ctx, plan, text, ds["text"], legacy=True)
            same(fast, slow, f"odd frames round {rounds}, attempt {attempt}")
            if "error" not in slow:
                break
            del lines[slow["error"][0]]
        assert "error" not in slow


NUMS = ["-", "-.", ".", "1.", ".5", "-0", "-0.0", "00012", "1.5.2", "12345678", "123456789", "-12345678", "1234567.8", "-1234.567",
        "0.000001", "9999999.9", "99999.99", "-99999.99", "1e3", "1E-2", "+5", "5,", ",5", "0x10", "1_000", " 7", "7 ", "1..2", "--1",
        "-1.", "16777217", "0.1", "0.25", "4294967296", "-2147483648", "2147483647", "nan", "inf", "-inf", "1.0000001", "33.333334"]


def test_numeric_tokens_at_the_edges_of_the_short_decoder(gpu_ctx, case):
    ds, plan = case
    rng = np.random.default_rng(11)
    for tok in NUMS:
        lines = []
        for ln in ds["lines"][:60]:
            c = ln.split("\t")
            info = c[7].split(";")
            k = rng.integers(0, 4)
            if k == 0:
                c[5] = tok                                # QUAL
            elif k == 1:
                info = [("MQ=" + tok) if kv.startswith("MQ=") else kv for kv in info]         # Float
            elif k == 2:
                info = [("DP=" + tok) if kv.startswith("DP=") else kv for kv in info]         # Integer
            else:
                sub = c[9].split(":")
                sub[2] = tok                              # FORMAT/DP
                c[9] = ":".join(sub)
            c[7] = ";".join(info)
            lines.append("\t".join(c))
        text = ("\n".join(lines) + "\n").encode()
        # one bad literal raises on the first record that carries it: compare record by record then
        fast = run(gpu_ctx, plan, text, ds["text"], legacy=False)
        slow = run(gpu_ctx, plan, text, ds["text"], legacy=True)
        same(fast, slow, f"token {tok!r}")
        if "error" in slow:
            for ln in lines[:12]:
                one = (ln + "\n").encode()
                same(run(gpu_ctx, plan, one, ds["text"], legacy=False), run(gpu_ctx, plan, one, ds["text"], legacy=True),
                     f"token {tok!r} alone")
