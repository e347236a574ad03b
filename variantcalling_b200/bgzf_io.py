"""BGZF / tabix containers around the hot path (host side).

The reference gets these from htslib: ``pysam.VariantFile`` reads and writes the
bgzip'ed VCF (``filter_variants_pipeline.py:106,115``) and ``bcftools index -t``
writes the ``.tbi`` (``:231``).  Neither is in this image, so the byte formats are
implemented here: block (de)compression is the multi-threaded C++ in
``csrc/hostio.cpp`` (zlib), the tabix index is built/parsed with NumPy.

Formats: BGZF = gzip members of <= 64 KiB with a ``BC`` extra field (SAM spec 4.1);
tabix index (``.tbi``) per the tabix format note: magic ``TBI\\1``, per-reference
binning index (``reg2bin``, 5 levels over 2^29) plus a 16 kb linear index of
virtual file offsets ``coffset << 16 | uoffset``.
"""
from __future__ import annotations

import ctypes as C
import os
import gzip
import struct

import numpy as np

from variantcalling_b200 import lib as _lib

BGZF_BLOCK_DATA = 0xFF00
TBI_SHIFT, TBI_LEVELS = 14, 5


def _check(rc: int, what: str):
    if rc:
        raise OSError(f"{what} failed (ugvc code {rc})")


# --------------------------------------------------------------------------- BGZF
def uncompressed_size(path: str) -> int:
    n = _lib.load_library().ugvc_bgzf_uncompressed_size(path.encode())
    if n < 0:
        raise OSError(f"{path}: not a readable BGZF file (ugvc code {n})")
    return int(n)


def inflate(path: str, voff_begin: int = 0, voff_end: int = 0, capacity: int | None = None, n_threads: int = 0,
            out: np.ndarray | None = None) -> np.ndarray:
    """Inflate [voff_begin, voff_end) (0,0 = whole file) into a uint8 array."""
    L = _lib.load_library()
    if out is None:
        if capacity is None:
            capacity = uncompressed_size(path) if voff_end == 0 and voff_begin == 0 else None
        if capacity is None:
            # a virtual-offset range: ask the reader itself -- with no room it only walks the block headers of
            # [voff_begin, voff_end) and reports the size (one contig of an indexed call set, not the whole file)
            probe = C.c_size_t()
            rc = L.ugvc_bgzf_inflate_file(path.encode(), voff_begin, voff_end, None, 0, C.byref(probe), 1)
            if rc not in (0, _lib.UGVC_E_ARG):
                _check(rc, f"inflate {path}")
            capacity = int(probe.value)
        out = np.empty(capacity + 64, dtype=np.uint8)
    n = C.c_size_t()
    rc = L.ugvc_bgzf_inflate_file(path.encode(), voff_begin, voff_end, out.ctypes.data_as(C.c_void_p), out.size,
                                  C.byref(n), n_threads)
    _check(rc, f"inflate {path}")
    return out[: n.value]


def range_info(path: str, voff_begin: int, voff_end: int) -> tuple[int, int, int, int]:
    """(c_begin, c_end, skip_head, take_bytes) of the blocks holding a virtual-offset range (ugvc_bgzf_range_info)."""
    c0, c1, skip, take = C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_uint64()
    rc = _lib.load_library().ugvc_bgzf_range_info(path.encode(), voff_begin, voff_end, C.byref(c0), C.byref(c1), C.byref(skip),
                                                  C.byref(take))
    _check(rc, f"range of {path}")
    return int(c0.value), int(c1.value), int(skip.value), int(take.value)


def first_block_text(comp: np.ndarray, skip_head: int) -> bytes:
    """The uncompressed bytes of the first BGZF block of `comp` from skip_head on (a sample of the records that follow)."""
    import struct
    import zlib

    raw = memoryview(comp)
    xlen = struct.unpack_from("<H", raw, 10)[0]
    bsize = struct.unpack_from("<H", raw, 16)[0] + 1
    return zlib.decompress(bytes(raw[12 + xlen: bsize - 8]), -15)[skip_head:]


# the empty BGZF block that ends a file (SAM spec 4.1.2)
BGZF_EOF = bytes([0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0,
                  0, 0, 0, 0, 0, 0, 0, 0])


def compress_bytes(data, level: int = 6, n_threads: int = 0) -> bytes:
    """BGZF-compress a byte string in memory (whole 0xff00-byte blocks, no EOF block); zlib releases
    the GIL, so the blocks are deflated on a thread pool.  Used to build compressed host buffers for
    ``ugvc_submit_bgzf`` (bench, tests); files go through :class:`BgzfWriter`."""
    import struct
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    view = memoryview(data)
    n = len(view)

    def one(i: int) -> bytes:
        chunk = bytes(view[i:i + BGZF_BLOCK_DATA])
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 8)
        payload = c.compress(chunk) + c.flush()
        if 18 + len(payload) + 8 > 65536:  # incompressible: store
            c = zlib.compressobj(0, zlib.DEFLATED, -15, 8)
            payload = c.compress(chunk) + c.flush()
        bsize = 18 + len(payload) + 8
        return (b"\x1f\x8b\x08\x04" + b"\0" * 6 + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) + payload
                + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))

    with ThreadPoolExecutor(max_workers=n_threads or min(64, (os.cpu_count() or 8))) as pool:
        return b"".join(pool.map(one, range(0, n, BGZF_BLOCK_DATA), chunksize=16))


def count_lines(text: np.ndarray, n_threads: int = 0) -> int:
    """Number of newline bytes in a uint8 array (threaded memchr; NumPy's compare + count is one core)."""
    buf = np.ascontiguousarray(text)
    n = _lib.load_library().ugvc_count_byte(buf.ctypes.data_as(C.c_void_p) if buf.size else None, buf.size, 10, n_threads)
    if n < 0:
        _check(int(n), "count_lines")
    return int(n)


class BgzfWriter:
    """Append-only BGZF writer that remembers every block's compressed size, so
    virtual offsets of written bytes can be computed for the tabix index."""

    def __init__(self, path: str, level: int = 6, n_threads: int = 0):
        self.path, self.level, self.n_threads = path, level, n_threads
        self.coffset = 0                 # compressed bytes written so far
        self.uoffset = 0                 # uncompressed bytes written so far
        self._started = False
        # piecewise map uncompressed offset -> (block coffset, block uoffset)
        self.block_u: list[np.ndarray] = []
        self.block_c: list[np.ndarray] = []

    def write(self, data: np.ndarray | bytes, last: bool = False):
        buf = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data
        L = _lib.load_library()
        n_blocks = (buf.size + BGZF_BLOCK_DATA - 1) // BGZF_BLOCK_DATA
        csz = np.zeros(max(1, n_blocks), dtype=np.uint32)
        total, nb = C.c_uint64(), C.c_size_t()
        rc = L.ugvc_bgzf_deflate_to_file(self.path.encode(), b"a" if self._started else b"w",
                                         buf.ctypes.data_as(C.c_void_p) if buf.size else None, buf.size, self.level,
                                         1 if last else 0, self.n_threads, C.byref(total),
                                         csz.ctypes.data_as(C.c_void_p), csz.size, C.byref(nb))
        _check(rc, f"deflate to {self.path}")
        self._started = True
        if n_blocks:
            cs = csz[:n_blocks].astype(np.int64)
            c_start = self.coffset + np.concatenate(([0], np.cumsum(cs)[:-1]))
            u_start = self.uoffset + np.arange(n_blocks, dtype=np.int64) * BGZF_BLOCK_DATA
            self.block_c.append(c_start)
            self.block_u.append(u_start)
        self.coffset += int(total.value)
        self.uoffset += int(buf.size)

    def write_compressed(self, blocks: np.ndarray, block_csize: np.ndarray, n_uncompressed: int, chunk: int):
        """Append BGZF blocks compressed elsewhere (the device encoder: `chunk` uncompressed bytes per block, the last
        one shorter) and keep the offset map the index needs."""
        with open(self.path, "ab" if self._started else "wb") as fh:
            fh.write(memoryview(np.ascontiguousarray(blocks)))
        self._started = True
        n_blocks = int(block_csize.size)
        if n_blocks:
            cs = block_csize.astype(np.int64)
            self.block_c.append(self.coffset + np.concatenate(([0], np.cumsum(cs)[:-1])))
            self.block_u.append(self.uoffset + np.arange(n_blocks, dtype=np.int64) * chunk)
            self.coffset += int(cs.sum())
        self.uoffset += int(n_uncompressed)

    def close(self):
        if not self._started:
            self.write(np.zeros(0, np.uint8), last=True)
        else:
            L = _lib.load_library()
            total = C.c_uint64()
            rc = L.ugvc_bgzf_deflate_to_file(self.path.encode(), b"a", None, 0, self.level, 1, 1, C.byref(total),
                                             None, 0, None)
            _check(rc, f"close {self.path}")
            self.coffset += int(total.value)

    def virtual_offsets(self, uoffsets: np.ndarray) -> np.ndarray:
        """Virtual file offsets of uncompressed stream offsets (of bytes already written)."""
        if not self.block_u:
            return np.zeros_like(uoffsets, dtype=np.uint64)
        bu = np.concatenate(self.block_u)
        bc = np.concatenate(self.block_c)
        idx = np.searchsorted(bu, uoffsets, side="right") - 1
        return (bc[idx].astype(np.uint64) << np.uint64(16)) | (uoffsets - bu[idx]).astype(np.uint64)


def read_header_text(path: str) -> str:
    """Header lines of a bgzip'ed / plain VCF (reads only as far as needed)."""
    with open(path, "rb") as fh:
        gz = fh.read(2) == b"\x1f\x8b"
    opener = gzip.open if gz else open
    lines = []
    with opener(path, "rb") as fh:
        for raw in fh:
            if not raw.startswith(b"#"):
                break
            lines.append(raw.decode())
    return "".join(lines)


# --------------------------------------------------------------------------- tabix
def reg2bin(beg: np.ndarray, end: np.ndarray) -> np.ndarray:
    """UCSC binning (tabix spec): beg 0-based inclusive, end exclusive; vectorised."""
    end = end - 1
    out = np.zeros(beg.shape, dtype=np.int64)
    done = np.zeros(beg.shape, dtype=bool)
    for shift, offset in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        same = ((beg >> shift) == (end >> shift)) & ~done
        out[same] = offset + (beg[same] >> shift)
        done |= same
    return out


def tbi_section(beg: np.ndarray, end: np.ndarray, voff_start: np.ndarray, voff_end: np.ndarray) -> bytes:
    """The binning index + linear index of ONE reference sequence (records sorted by position): the bytes that follow
    the names in a ``.tbi``.  beg/end are 0-based half-open; voff_* the virtual offsets of each record's first byte
    and one past its last byte."""
    if len(beg) == 0:
        return struct.pack("<i", 0) + struct.pack("<i", 0)
    b, e = np.asarray(beg, dtype=np.int64), np.asarray(end, dtype=np.int64)
    vs, ve = np.asarray(voff_start, dtype=np.uint64), np.asarray(voff_end, dtype=np.uint64)
    out = []
    bins = reg2bin(b, np.maximum(e, b + 1))
    # chunks: runs of consecutive records with the same bin
    change = np.flatnonzero(np.concatenate(([True], bins[1:] != bins[:-1])))
    run_bin = bins[change]
    run_beg = vs[change]
    run_end = ve[np.concatenate((change[1:] - 1, [bins.size - 1]))]
    order = np.argsort(run_bin, kind="stable")
    run_bin, run_beg, run_end = run_bin[order], run_beg[order], run_end[order]
    uniq, first = np.unique(run_bin, return_index=True)
    counts = np.diff(np.concatenate((first, [run_bin.size])))
    out.append(struct.pack("<i", uniq.size))
    # every bin is (u32 bin, i32 n_chunks, n_chunks x (u64 begin, u64 end)): lay all of them out at once.
    # Bin k's header starts at 8-byte word k + 2 * first[k]; run r (of bin k) follows at word (k + 1) + 2 r.
    words = np.zeros(uniq.size + 2 * run_bin.size, dtype="<u8")
    halves = words.view("<u4")
    head = np.arange(uniq.size, dtype=np.int64) + 2 * first
    halves[2 * head] = uniq.astype(np.uint32)
    halves[2 * head + 1] = counts.astype(np.uint32)
    at = np.repeat(np.arange(uniq.size, dtype=np.int64), counts) + 1 + 2 * np.arange(run_bin.size, dtype=np.int64)
    words[at] = run_beg
    words[at + 1] = run_end
    out.append(words.tobytes())
    # linear index: smallest virtual offset of any record overlapping each 16 kb window
    n_win = int((np.maximum(e, b + 1).max() - 1) >> TBI_SHIFT) + 1
    lin = np.full(n_win, np.iinfo(np.uint64).max, dtype=np.uint64)
    w0 = b >> TBI_SHIFT
    w1 = (np.maximum(e, b + 1) - 1) >> TBI_SHIFT
    if np.all(w0[1:] >= w0[:-1]) and np.all(vs[1:] >= vs[:-1]):  # sorted file: the first record of a window is its minimum
        win, first_rec = np.unique(w0, return_index=True)
        lin[win] = vs[first_rec]
    else:
        np.minimum.at(lin, w0, vs)
    span = np.flatnonzero(w1 > w0)
    for i in span:  # records crossing window borders are rare (long REF alleles)
        lin[w0[i] + 1: w1[i] + 1] = np.minimum(lin[w0[i] + 1: w1[i] + 1], vs[i])
    # empty windows take the offset of the next filled one, as htslib back-fills its linear index
    # (hts_idx_finish: offset[l] = offset[l + 1] from the end); the last window always holds a record
    filled = lin != np.iinfo(np.uint64).max
    nxt = np.minimum.accumulate(np.where(filled, np.arange(n_win), n_win - 1)[::-1])[::-1]
    lin = lin[nxt]
    out.append(struct.pack("<i", n_win))
    out.append(lin.astype("<u8").tobytes())
    return b"".join(out)


def tbi_assemble(contig_names: list[str], sections: list[bytes]) -> bytes:
    """Uncompressed ``.tbi`` payload from the per-sequence sections (tbi_section), in file order."""
    names = b"".join(n.encode() + b"\0" for n in contig_names)
    return b"".join([b"TBI\x01", struct.pack("<8i", len(contig_names), 2, 1, 2, 0, ord("#"), 0, len(names)), names, *sections])


def build_tbi(contig_names: list[str], contig_of: np.ndarray, beg: np.ndarray, end: np.ndarray,
              voff_start: np.ndarray, voff_end: np.ndarray) -> bytes:
    """Uncompressed ``.tbi`` payload for records sorted by (contig block, position).

    contig_of[i] indexes contig_names; beg/end are 0-based half-open; voff_* are the
    virtual offsets of each record's first byte and one past its last byte.
    """
    sections = []
    for ci in range(len(contig_names)):
        sel = np.flatnonzero(contig_of == ci)
        sections.append(tbi_section(beg[sel], end[sel], voff_start[sel], voff_end[sel]))
    return tbi_assemble(contig_names, sections)


def write_tbi(path: str, payload: bytes):
    w = BgzfWriter(path, level=6, n_threads=1)
    w.write(payload, last=True)


def read_tbi(path: str, linear: bool = False):
    """{contig: (min chunk begin voff, max chunk end voff)} in file order, from a ``.tbi``.  linear=True: also
    {contig: the distinct virtual offsets of its linear index, ascending} -- each is where a record starts, so a
    contig can be cut there into pieces that are read (and filtered) independently."""
    with gzip.open(path, "rb") as fh:
        data = fh.read()
    if data[:4] != b"TBI\x01":
        raise OSError(f"{path}: not a tabix index")
    (n_ref,) = struct.unpack_from("<i", data, 4)
    n_ref = max(0, n_ref)
    L = _lib.load_library()
    buf = np.frombuffer(data, dtype=np.uint8)
    lo, hi = np.empty(n_ref, np.uint64), np.empty(n_ref, np.uint64)
    l_off, l_cnt = np.empty(n_ref, np.int64), np.empty(n_ref, np.int32)
    got, nm_off, nm_len = C.c_int32(), C.c_int64(), C.c_int32()
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    _check(L.ugvc_tbi_summary(p(buf), buf.size, n_ref, p(lo), p(hi), p(l_off), p(l_cnt), C.byref(got), C.byref(nm_off),
                               C.byref(nm_len)), f"{path}: malformed tabix index")
    names = data[nm_off.value:nm_off.value + nm_len.value].split(b"\0")[:n_ref]
    out, lin = {}, {}
    for r in range(n_ref):
        if int(hi[r]) == 0 and int(lo[r]) == 0xFFFFFFFFFFFFFFFF:
            continue  # no records on this contig
        name = names[r].decode()
        out[name] = (int(lo[r]), int(hi[r]))
        if linear:
            io = np.unique(np.frombuffer(data, dtype="<u8", count=int(l_cnt[r]), offset=int(l_off[r])))
            lin[name] = io[(io > lo[r]) & (io < hi[r])]
    return (out, lin) if linear else out


def write_vcf_gz(path: str, header_lines: list[str], record_lines: list[str], n_threads: int = 0):
    """Write a bgzip'ed VCF plus its ``.tbi`` from text lines (tests / fixtures)."""
    hdr = ("\n".join(header_lines) + "\n").encode()
    body = ("\n".join(record_lines) + ("\n" if record_lines else "")).encode()
    w = BgzfWriter(path, n_threads=n_threads)
    w.write(hdr)
    w.write(body)
    w.close()
    if not record_lines:
        write_tbi(path + ".tbi", build_tbi([], np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64),
                                           np.zeros(0, np.uint64), np.zeros(0, np.uint64)))
        return
    names: list[str] = []
    contig_of, beg, end = [], [], []
    for ln in record_lines:
        c, pos, _id, ref, _rest = ln.split("\t", 4)
        if not names or names[-1] != c:
            if c in names:
                raise ValueError("records of a contig must be contiguous")
            names.append(c)
        contig_of.append(len(names) - 1)
        beg.append(int(pos) - 1)
        end.append(int(pos) - 1 + len(ref))
    lens = np.array([len(ln.encode()) + 1 for ln in record_lines], dtype=np.int64)
    u_start = len(hdr) + np.concatenate(([0], np.cumsum(lens)[:-1]))
    u_end = u_start + lens
    vs = w.virtual_offsets(u_start)
    # one past the record's last byte, expressed inside the block that holds that byte
    ve = w.virtual_offsets(u_end - 1) + np.uint64(1)
    write_tbi(path + ".tbi", build_tbi(names, np.array(contig_of), np.array(beg), np.array(end), vs, ve))
