"""GPU: BGZF inflate on the device (csrc/inflate.cuh, ugvc_bgzf_inflate_device / ugvc_submit_bgzf):
byte-identical with zlib on every block type, and the filter results of compressed input equal those
of the plain text."""
import struct
import zlib

import numpy as np
import pytest

from tests import util
from variantcalling_b200 import bgzf_io, lib
from variantcalling_b200 import model_compiler as MC
from variantcalling_b200.vcf_header import VcfHeader

pytestmark = pytest.mark.gpu


def bgzf_block(payload: bytes, data: bytes) -> bytes:
    """One BGZF block around a raw DEFLATE payload (SAM spec 4.1)."""
    bsize = 18 + len(payload) + 8
    return (b"\x1f\x8b\x08\x04" + b"\0" * 6 + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) + payload
            + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def raw_deflate(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return c.compress(data) + c.flush()


def bgzf_bytes(data: bytes, chunk=0xff00, **kw) -> bytes:
    return b"".join(bgzf_block(raw_deflate(data[i:i + chunk], **kw), data[i:i + chunk]) for i in range(0, len(data), chunk))


@pytest.fixture(scope="module")
def ds():
    return util.make_dataset(n_records=6000, n_custom=6, seed=23)


@pytest.fixture()
def ctx(gpu_ctx, ds):
    plan = MC.compile_plan_no_model(VcfHeader(ds["header_text"]))
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.reserve(8 << 20, 20000, 2)
    return gpu_ctx


@pytest.mark.parametrize("kind", ["level1", "level6", "level9", "stored", "fixed", "huffman_only", "rle", "tiny_blocks"])
def test_inflate_equals_zlib_on_every_block_type(ctx, ds, kind):
    text = ds["text"]
    kw = {"level1": dict(level=1), "level6": dict(level=6), "level9": dict(level=9), "stored": dict(level=0),
          "fixed": dict(level=6, strategy=zlib.Z_FIXED), "huffman_only": dict(level=6, strategy=zlib.Z_HUFFMAN_ONLY),
          "rle": dict(level=6, strategy=zlib.Z_RLE), "tiny_blocks": dict(level=6)}[kind]
    comp = bgzf_bytes(text, chunk=700 if kind == "tiny_blocks" else 0xff00, **kw) + bgzf_io.BGZF_EOF
    got = ctx.inflate_bgzf(comp)
    assert got.tobytes() == text
    assert ctx.inflate_bgzf(comp, want_text=False) == len(text)


def test_binary_and_degenerate_inputs(ctx):
    rng = np.random.default_rng(1)
    noise = rng.integers(0, 256, size=200_000, dtype=np.uint8).tobytes()      # incompressible: stored blocks inside level 6
    runs = (b"A" * 70_000) + bytes(range(256)) * 300 + b"\n"                       # long matches, distance 1 and 256
    for data in (noise, runs, b"x", b""):
        comp = bgzf_bytes(data) + bgzf_io.BGZF_EOF
        assert ctx.inflate_bgzf(comp).tobytes() == data
    assert ctx.inflate_bgzf(b"").size == 0
    # what our own writer produces (threaded zlib, 0xff00-byte blocks)
    import os
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        w = bgzf_io.BgzfWriter(os.path.join(d, "x.gz"), level=6, n_threads=2)
        w.write(runs)
        w.close()
        assert ctx.inflate_bgzf(open(os.path.join(d, "x.gz"), "rb").read()).tobytes() == runs


def test_corrupt_blocks_are_refused(ctx, ds):
    comp = bytearray(bgzf_bytes(ds["text"][:50_000]))
    with pytest.raises(lib.UgvcError, match="not a BGZF block"):
        ctx.inflate_bgzf(bytes(comp[1:]))
    bad = bytearray(comp)
    bad[40] ^= 0x55                                                              # inside the first DEFLATE payload
    with pytest.raises(lib.UgvcError):
        ctx.inflate_bgzf(bytes(bad))
    with pytest.raises(lib.UgvcError, match="larger than the reserved"):
        ctx.inflate_bgzf(bgzf_bytes(b"y" * (9 << 20)))
    assert ctx.inflate_bgzf(bytes(comp)).tobytes() == ds["text"][:50_000]      # the context is still usable


def test_filtering_compressed_input_equals_filtering_the_text(gpu_ctx, ds):
    _, tr, x = util.fit_transformer(ds)
    model = util.fit_model("gb_small", x, ds["labels"])
    plan = MC.compile_plan(VcfHeader(ds["header_text"]), tr, model, ds["customs"])
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.set_key_order(*lib.learn_key_order(ds["text"]))
    gpu_ctx.reserve(len(ds["text"]) + 4096, len(ds["lines"]) + 16, 2)
    want = gpu_ctx.filter_batch(ds["text"], 30.0)
    comp = np.frombuffer(bgzf_bytes(ds["text"], level=6), dtype=np.uint8)
    n_max = len(ds["lines"]) + 1
    for lane in (0, 1):
        gpu_ctx.submit_bgzf(lane, comp, comp.size, 30.0)
        out = gpu_ctx.alloc_outputs(n_max, want_recinfo=True)
        n = gpu_ctx.collect(lane, out, n_max)
        got = gpu_ctx.trim_outputs(out, n)
        assert n == want["n_records"]
        for k in ("low_score", "probs", "qual", "line_start"):
            assert np.array_equal(got[k], want[k]), k
        assert got["recinfo"].tobytes() == want["recinfo"].tobytes()
