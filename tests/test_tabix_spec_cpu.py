"""CPU: the .vcf.gz / .tbi the product writes, read by an independent spec-based parser (oracle/tabix_ref.py):
every BGZF block is well formed (BSIZE, CRC32, ISIZE, EOF marker), the index header says VCF, every record sits in
the bin the spec's reg2bin assigns, the linear index never points past a record of its window, and region queries
through bins + linear index return exactly the records a scan of the file finds."""
import gzip

import numpy as np
import pytest

from oracle import tabix_ref as TR
from variantcalling_b200 import bgzf_io, synth


@pytest.fixture(scope="module")
def written(tmp_path_factory):
    d = tmp_path_factory.mktemp("tbx")
    spec = synth.SynthSpec(n_records=6000, n_custom=3, seed=23,
                           contigs={"chr1": 3_000_000, "chr2": 900_000, "chrM": 16_000, "chr9": 400_000})
    header, lines, _ = synth.generate(spec)
    path = str(d / "calls.vcf.gz")
    bgzf_io.write_vcf_gz(path, header, lines, n_threads=3)
    return path, header, lines


def test_every_block_is_a_valid_bgzf_member(written):
    path, header, lines = written
    blocks = TR.read_bgzf_blocks(path)
    total = sum(b[2] for b in blocks)
    assert total == len(("\n".join(header) + "\n" + "\n".join(lines) + "\n").encode())
    assert blocks[-1][2] == 0 and all(b[1] <= 65536 for b in blocks)


def test_index_header_bins_and_linear_index_follow_the_spec(written):
    path, _header, lines = written
    idx = TR.TabixIndex(path + ".tbi")
    assert (idx.format, idx.col_seq, idx.col_beg, idx.col_end, idx.meta, idx.skip) == (2, 1, 2, 0, ord("#"), 0)
    by_contig = {}
    for ln in lines:
        by_contig.setdefault(ln.split("\t", 1)[0], []).append(ln)
    assert idx.names == list(by_contig)
    raw = open(path, "rb").read()
    for tid, name in enumerate(idx.names):
        want_bins = {}
        for ln in by_contig[name]:
            c = ln.split("\t", 5)
            b0 = int(c[1]) - 1
            want_bins.setdefault(TR.reg2bin(b0, b0 + max(1, len(c[3]))), []).append(ln)
        assert set(idx.bins[tid]) == set(want_bins)
        cache = {}
        for b, chunks in idx.bins[tid].items():
            got = b"".join(TR.read_virtual(raw, cache, cb, ce) for cb, ce in chunks).decode().split("\n")
            got = [g for g in got if g]
            assert all(g in want_bins[b] or TR.reg2bin(int(g.split("\t")[1]) - 1, int(g.split("\t")[1]) - 1 + max(1, len(g.split("\t")[3]))) != b
                       for g in got)  # a chunk may span records of other bins, never miss one of its own
            assert set(want_bins[b]) <= set(got)
        # linear index: offset of window w is at or before the first record overlapping it
        lin = idx.linear[tid]
        first_voff = {}
        for b, chunks in idx.bins[tid].items():
            for cb, _ce in chunks:
                for ln in TR.read_virtual(raw, cache, cb, _ce).decode().split("\n"):
                    if ln:
                        c = ln.split("\t", 5)
                        b0 = int(c[1]) - 1
                        for w in range(b0 >> 14, ((b0 + max(1, len(c[3])) - 1) >> 14) + 1):
                            first_voff[w] = min(first_voff.get(w, 1 << 62), cb)
        assert len(lin) == max(first_voff) + 1
        for w, v in first_voff.items():
            assert lin[w] <= v or lin[w] >> 16 == v >> 16


def test_region_queries_equal_a_scan(written):
    path, _header, lines = written
    idx = TR.TabixIndex(path + ".tbi")
    rng = np.random.default_rng(3)
    recs = [(ln.split("\t", 1)[0], int(ln.split("\t", 2)[1]) - 1, ln) for ln in lines]
    for name in idx.names:
        mine = [(p, ln) for c, p, ln in recs if c == name]
        top = max(p for p, _ in mine) + 10
        for _ in range(25):
            a = int(rng.integers(0, top))
            b = a + int(rng.integers(1, 200_000))
            want = [ln.encode() for p, ln in mine if p < b and p + max(1, len(ln.split("\t")[3])) > a]
            assert TR.query(path, idx, name, a, b) == want, (name, a, b)
    # whole-contig query returns the contig
    assert TR.query(path, idx, "chrM", 0, 1 << 29) == [ln.encode() for c, _p, ln in recs if c == "chrM"]
    assert gzip.open(path).read().count(b"\n") == len(lines) + sum(1 for _ in open("/dev/null")) + len(_header)
