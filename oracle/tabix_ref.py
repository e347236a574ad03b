"""ORACLE (test infrastructure, not product code).

An independent reader of the two containers the tool writes, straight from their published specifications
(SAM/BAM spec section 4.1 "The BGZF compression format", tabix spec "The tabix index file format"):
``bcftools index -t`` / htslib readers are what the reference relies on (filter_variants_pipeline.py:106,231) and
neither is installed here.  Written without looking at variantcalling_b200/bgzf_io.py: the product's .tbi and
.vcf.gz are parsed with this module in the tests, so a mistake shared by the product's writer and the product's own
reader cannot hide.

  read_bgzf_blocks(path)          -> [(file offset, block size, uncompressed size)], checks magic / BSIZE / CRC32 / ISIZE / EOF
  TabixIndex(path)                -> header fields, per reference: bins {bin: [(beg voff, end voff)]}, linear index
  TabixIndex.query(...)           -> records overlapping a region, found the way the spec prescribes
                                     (reg2bins -> chunks, linear-index lower bound, scan of the candidate chunks)
"""
from __future__ import annotations

import gzip
import struct
import zlib

EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def read_bgzf_blocks(path: str, verify: bool = True) -> list[tuple[int, int, int]]:
    raw = open(path, "rb").read()
    out, at = [], 0
    while at < len(raw):
        id1, id2, cm, flg, _mtime, _xfl, _os, xlen = struct.unpack_from("<BBBBIBBH", raw, at)
        assert (id1, id2, cm) == (31, 139, 8) and flg & 4, f"not a BGZF member at {at}"
        x, bsize = at + 12, None
        while x < at + 12 + xlen:
            si1, si2, slen = struct.unpack_from("<BBH", raw, x)
            if (si1, si2, slen) == (66, 67, 2):
                bsize = struct.unpack_from("<H", raw, x + 4)[0] + 1
            x += 4 + slen
        assert bsize is not None, f"no BC subfield at {at}"
        crc, isize = struct.unpack_from("<II", raw, at + bsize - 8)
        assert isize <= 65536
        if verify:
            data = zlib.decompress(raw[at + 12 + xlen: at + bsize - 8], -15)
            assert len(data) == isize and (zlib.crc32(data) & 0xFFFFFFFF) == crc, f"bad block at {at}"
        out.append((at, bsize, isize))
        at += bsize
    assert raw[-28:] == EOF_BLOCK, "no EOF marker block"
    return out


def reg2bins(beg: int, end: int) -> list[int]:
    """Bins that may hold features overlapping [beg, end) (0-based, half open) -- the C code of the tabix spec."""
    end -= 1
    bins = [0]
    for shift, offset in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        bins.extend(range(offset + (beg >> shift), offset + (end >> shift) + 1))
    return bins


def reg2bin(beg: int, end: int) -> int:
    end -= 1
    for shift, offset in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return offset + (beg >> shift)
    return 0


class TabixIndex:
    def __init__(self, path: str):
        data = gzip.open(path, "rb").read()
        assert data[:4] == b"TBI\x01"
        (n_ref, self.format, self.col_seq, self.col_beg, self.col_end, self.meta, self.skip, l_nm) = struct.unpack_from("<8i", data, 4)
        at = 36
        names = data[at:at + l_nm].split(b"\0")
        assert names[-1] == b"" and len(names) == n_ref + 1
        self.names = [n.decode() for n in names[:-1]]
        at += l_nm
        self.bins, self.linear = [], []
        for _ in range(n_ref):
            n_bin = struct.unpack_from("<i", data, at)[0]
            at += 4
            bins = {}
            for _b in range(n_bin):
                b, n_chunk = struct.unpack_from("<Ii", data, at)
                at += 8
                chunks = [struct.unpack_from("<QQ", data, at + 16 * k) for k in range(n_chunk)]
                at += 16 * n_chunk
                assert b not in bins
                bins[b] = chunks
            n_intv = struct.unpack_from("<i", data, at)[0]
            at += 4
            self.linear.append(list(struct.unpack_from(f"<{n_intv}Q", data, at)))
            at += 8 * n_intv
            self.bins.append(bins)
        self.n_no_coor = struct.unpack_from("<Q", data, at)[0] if at + 8 <= len(data) else None
        assert at == len(data) or at + 8 == len(data)

    def candidate_chunks(self, name: str, beg: int, end: int) -> list[tuple[int, int]]:
        tid = self.names.index(name)
        lin = self.linear[tid]
        win = beg >> 14
        min_off = lin[win] if win < len(lin) else (lin[-1] if lin else 0)
        chunks = []
        for b in reg2bins(beg, end):
            for cb, ce in self.bins[tid].get(b, []):
                if ce > min_off:
                    chunks.append((max(cb, min_off) if False else cb, ce))
        chunks.sort()
        return chunks


def read_virtual(path_bytes: bytes, blocks: dict[int, bytes], voff_begin: int, voff_end: int) -> bytes:
    """Uncompressed bytes of [voff_begin, voff_end); `blocks` caches inflated blocks by file offset."""
    out = []
    coff, uoff = voff_begin >> 16, voff_begin & 0xFFFF
    cend, uend = voff_end >> 16, voff_end & 0xFFFF
    while coff < cend or (coff == cend and uoff < uend):
        if coff not in blocks:
            xlen = struct.unpack_from("<H", path_bytes, coff + 10)[0]
            bsize = struct.unpack_from("<H", path_bytes, coff + 16)[0] + 1
            blocks[coff] = (zlib.decompress(path_bytes[coff + 12 + xlen: coff + bsize - 8], -15), bsize)
        data, bsize = blocks[coff]
        if coff == cend:
            out.append(data[uoff:uend])
            break
        out.append(data[uoff:])
        coff += bsize
        uoff = 0
    return b"".join(out)


def query(vcf_path: str, index: TabixIndex, name: str, beg: int, end: int) -> list[bytes]:
    """VCF data lines of `name` overlapping [beg, end) (0-based half open), the way a tabix reader finds them."""
    raw = open(vcf_path, "rb").read()
    cache: dict[int, tuple[bytes, int]] = {}
    hits, seen = [], set()
    for cb, ce in index.candidate_chunks(name, beg, end):
        for line in read_virtual(raw, cache, cb, ce).split(b"\n"):
            if not line or line.startswith(b"#"):
                continue
            cols = line.split(b"\t", 8)
            if cols[0].decode() != name:
                continue
            pos0 = int(cols[1]) - 1
            rec_end = pos0 + max(1, len(cols[3]))
            # htslib: a record ends at INFO/END when the tag is there and beyond POS (rlen = END - pos)
            for kv in (cols[7].split(b";") if len(cols) > 7 else []):  # noqa: PLR2004
                if kv.startswith(b"END=") and kv[4:].isdigit():
                    rec_end = int(kv[4:]) if int(kv[4:]) > pos0 else rec_end
                    break
            if pos0 < end and rec_end > beg and (cols[1], line) not in seen:
                seen.add((cols[1], line))
                hits.append((pos0, line))
    hits.sort(key=lambda t: t[0])
    return [ln for _, ln in hits]
