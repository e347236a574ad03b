"""GPU: the device-side file path (ugvc_filter_bgzf: BGZF inflate -> K1..K3 -> record writer -> BGZF deflate) writes
exactly the text the host writer (ugvc_splice_records) produces from the host-buffer path's results, as valid BGZF
blocks, with line offsets that point at the records; ranges that start inside a block; the fallback signal."""
import ctypes as C
import gzip
import struct
import zlib

import numpy as np
import pytest

from tests import util
from variantcalling_b200 import bgzf_io, lib
from variantcalling_b200 import model_compiler as MC
from variantcalling_b200.vcf_header import VcfHeader

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def job():
    ds = util.make_dataset(n_records=6000, n_custom=6, seed=91)
    _, tr, x = util.fit_transformer(ds)
    model = util.fit_model("gb_small", x, ds["labels"])
    plan = MC.compile_plan(VcfHeader(ds["header_text"]), tr, model, ds["customs"])
    return ds, plan


def host_text(ctx, text: bytes, flags: int) -> bytes:
    """The host-buffer path + the host writer on the same records."""
    L = lib.load_library()
    res = ctx.filter_batch(text, 30.0)
    n = res["n_records"]
    buf = np.frombuffer(text, dtype=np.uint8)
    out = np.empty(buf.size + n * 128 + 1024, dtype=np.uint8)
    out_ls = np.empty(n + 1, dtype=np.int64)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
    code = table = off = None
    if flags & lib.FILE_BLACKLIST_CG:
        code = (res["recinfo"]["flags"] & 1).astype(np.int32)
        strings = [b"PASS", b"CG_NON_HMER_INDEL"]
        table = np.frombuffer(b"".join(strings), dtype=np.uint8)
        off = np.array([0, 4, 4 + len(strings[1])], dtype=np.int64)
    nb = L.ugvc_splice_records(p(buf), p(res["line_start"]), p(res["recinfo"]), p(res["low_score"]), p(res["qual"]), n,
                               int(bool(flags & lib.FILE_OVERWRITE_QUAL)), 1, p(code), p(table), p(off), None, 0, p(out),
                               out.size, p(out_ls), 4)
    assert nb > 0
    return out[:nb].tobytes(), res


def check_blocks(comp: bytes, csizes) -> bytes:
    at, parts = 0, []
    for cs in csizes:
        blk = comp[at:at + int(cs)]
        assert blk[:4] == b"\x1f\x8b\x08\x04" and struct.unpack_from("<H", blk, 16)[0] + 1 == len(blk)
        data = zlib.decompress(blk[18:-8], -15)
        crc, isize = struct.unpack_from("<II", blk, len(blk) - 8)
        assert isize == len(data) and crc == (zlib.crc32(data) & 0xFFFFFFFF)
        parts.append(data)
        at += int(cs)
    assert at == len(comp)
    return b"".join(parts)


@pytest.mark.parametrize("flags", [0, lib.FILE_BLACKLIST_CG, lib.FILE_OVERWRITE_QUAL | lib.FILE_BLACKLIST_CG])
def test_device_file_path_equals_host_writer(gpu_ctx, job, flags):
    ds, plan = job
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.set_key_order(*lib.learn_key_order(ds["text"]))
    gpu_ctx.reserve(len(ds["text"]) + (1 << 20), len(ds["lines"]) + 64, 1)
    want, res = host_text(gpu_ctx, ds["text"], flags)
    comp = np.frombuffer(bgzf_io.compress_bytes(ds["text"]), dtype=np.uint8)
    got = gpu_ctx.filter_bgzf(comp, 0, len(ds["text"]), 30.0, flags, len(ds["lines"]) + 64)
    assert got is not None and got["n_records"] == res["n_records"]
    out = check_blocks(got["bgzf"].tobytes(), got["block_csize"])
    assert out == want, "the device writer's text differs from the host writer's"
    assert gzip.decompress(got["bgzf"].tobytes()) == want  # a plain gzip reader agrees
    ls = got["line_start"]
    lines = want.split(b"\n")
    for i in (0, 1, len(lines) // 2, res["n_records"] - 1):
        assert want[ls[i]:ls[i + 1] - 1] == lines[i]
    assert np.array_equal(got["low_score"], res["low_score"])
    assert all(int(c) <= 65536 for c in got["block_csize"]) and len(got["block_csize"]) == (len(want) + lib.DEF_CHUNK - 1) // lib.DEF_CHUNK


def test_range_inside_blocks(gpu_ctx, job):
    """A contig of an indexed file starts and ends inside BGZF blocks: skip_head / take_bytes select the lines."""
    ds, plan = job
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.set_key_order(*lib.learn_key_order(ds["text"]))
    gpu_ctx.reserve(len(ds["text"]) + (1 << 20), len(ds["lines"]) + 64, 1)
    lines = ds["lines"]
    a, b = 1500, 4200
    head = ("\n".join(lines[:a]) + "\n").encode()
    mid = ("\n".join(lines[a:b]) + "\n").encode()
    comp = bgzf_io.compress_bytes(ds["text"])
    # blocks hold 0xff00 bytes each: the range starts in block len(head) // 0xff00
    blk = 0xff00
    first = len(head) // blk
    sizes, at = [], 0
    while at < len(comp):
        sizes.append(struct.unpack_from("<H", comp, at + 16)[0] + 1)
        at += sizes[-1]
    c0 = sum(sizes[:first])
    part = np.frombuffer(comp[c0:], dtype=np.uint8)
    want, _ = host_text(gpu_ctx, mid, 0)
    got = gpu_ctx.filter_bgzf(part, len(head) - first * blk, len(mid), 30.0, 0, len(lines) + 64)
    assert got is not None and got["n_records"] == b - a
    assert gzip.decompress(got["bgzf"].tobytes()) == want


def test_fallback_when_the_general_writer_is_needed(gpu_ctx, job):
    ds, plan = job
    gpu_ctx.load_plan(plan.blob)
    gpu_ctx.set_key_order(*lib.learn_key_order(ds["text"]))
    gpu_ctx.reserve(len(ds["text"]) + (1 << 20), len(ds["lines"]) + 64, 1)
    lines = list(ds["lines"][:400])
    cols = lines[7].split("\t")
    cols[7] = cols[7] + ";TREE_SCORE=12.5"  # already filtered once: the key has to be replaced, not appended
    lines[7] = "\t".join(cols)
    text = ("\n".join(lines) + "\n").encode()
    comp = np.frombuffer(bgzf_io.compress_bytes(text), dtype=np.uint8)
    assert gpu_ctx.filter_bgzf(comp, 0, len(text), 30.0, 0, 512) is None
    # the same records without that line go through
    del lines[7]
    text = ("\n".join(lines) + "\n").encode()
    comp = np.frombuffer(bgzf_io.compress_bytes(text), dtype=np.uint8)
    got = gpu_ctx.filter_bgzf(comp, 0, len(text), 30.0, 0, 512)
    assert got is not None and gzip.decompress(got["bgzf"].tobytes()) == host_text(gpu_ctx, text, 0)[0]
