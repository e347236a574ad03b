"""CPU: BGZF / tabix containers and the record splicer (host C++) against Python's gzip and the
oracle's writer rules (filter_variants_pipeline.py:188-228)."""
import ctypes as C
import gzip

import numpy as np
import pytest

from oracle import ref_pipeline as R
from oracle.vcf_reader import OracleVariantFile
from variantcalling_b200 import bgzf_io, lib, synth


@pytest.fixture(scope="module")
def small(tmp_path_factory):
    d = tmp_path_factory.mktemp("io")
    spec = synth.SynthSpec(n_records=3000, n_custom=3, seed=11)
    header, lines, _ = synth.generate(spec)
    path = str(d / "in.vcf.gz")
    bgzf_io.write_vcf_gz(path, header, lines, n_threads=4)
    return dict(path=path, header=header, lines=lines)


def test_bgzf_roundtrip_and_eof_block(small):
    raw = open(small["path"], "rb").read()
    assert raw[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    text = gzip.open(small["path"]).read().decode()
    assert text == "\n".join(small["header"]) + "\n" + "\n".join(small["lines"]) + "\n"
    assert bgzf_io.uncompressed_size(small["path"]) == len(text)
    assert bgzf_io.inflate(small["path"], n_threads=3).tobytes().decode() == text
    assert bgzf_io.read_header_text(small["path"]) == "\n".join(small["header"]) + "\n"


def test_tabix_ranges_select_exactly_each_contig(small):
    idx = bgzf_io.read_tbi(small["path"] + ".tbi")
    by_contig = {}
    for ln in small["lines"]:
        by_contig.setdefault(ln.split("\t", 1)[0], []).append(ln)
    assert list(idx) == list(by_contig)
    for c, (vb, ve) in idx.items():
        raw = bgzf_io.inflate(small["path"], vb, ve, n_threads=2).tobytes().decode()
        assert raw.endswith("\n") and raw.split("\n")[:-1] == by_contig[c]


def test_reg2bin_known_values():
    b = np.array([0, 16383, 16384, 0, 1 << 28], dtype=np.int64)
    e = np.array([1, 16384, 16385, 1 << 29, (1 << 28) + 5], dtype=np.int64)
    assert list(bgzf_io.reg2bin(b, e)) == [4681, 4681, 4682, 0, 4681 + (1 << 14)]


def _recinfo_for(lines):
    ri = np.zeros(len(lines), dtype=lib.RECINFO_DTYPE)
    for i, ln in enumerate(lines):
        cols = ln.split("\t")
        start = lambda k: len("\t".join(cols[:k])) + 1  # noqa: E731
        n_alleles = 1 + (0 if cols[4] == "." else cols[4].count(",") + 1)
        ri[i] = (int(cols[1]), start(5), start(6), start(7), start(8), (len(cols[3]) << 8) | (n_alleles << 1))
    return ri


@pytest.mark.parametrize("overwrite_qual", [False, True])
def test_splicer_matches_oracle_writer(overwrite_qual):
    hdr = ["##fileformat=VCFv4.2", "##contig=<ID=c1,length=100>",
           "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS"]
    lines = ["c1\t1\t.\tA\tC\t5.5\tPASS\tX=1\tGT\t0/1", "c1\t2\t.\tA\tC\t.\tq10\t.\tGT\t1/1",
             "c1\t3\trs\tAT\tA\t7\t.\tTREE_SCORE=1;Y=2\tGT:DP\t0/1:3", "c1\t4\t.\tG\tGGC\t8\tq10;PASS\tBLACKLST=old;Z\tGT\t0/0",
             "c1\t5\t.\tG\tT\t9\tLOW_SCORE\tA=1", "c1\t6\t.\tG\tT\t9\t.\tA=1"]
    quals = np.array([12.5, 30.0, 45.123456789, 0.0, 29.999999, 1e-7])
    low = (quals <= 30.0).astype(np.uint8)
    bl_strings = [b"PASS;PASS", b"CG_NON_HMER_INDEL;PASS", b"PASS;COHORT_FP", b"CG_NON_HMER_INDEL;COHORT_FP"]
    codes = np.array([0, 2, 0, 3, 1, 0], dtype=np.int32)
    text = ("\n".join(lines) + "\n").encode()
    buf = np.frombuffer(text, dtype=np.uint8)
    ls = np.concatenate(([0], np.cumsum([len(l) + 1 for l in lines]))).astype(np.int64)
    ri = _recinfo_for(lines)
    table = np.frombuffer(b"".join(bl_strings), dtype=np.uint8)
    off = np.concatenate(([0], np.cumsum([len(s) for s in bl_strings]))).astype(np.int64)
    out = np.empty(len(text) + 1024, dtype=np.uint8)
    out_ls = np.empty(len(lines) + 1, dtype=np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    n = lib.load_library().ugvc_splice_records(p(buf), p(ls), p(ri), p(low), p(quals), len(lines), int(overwrite_qual),
                                               1, p(codes), p(table), p(off), None, 0, p(out), out.size, p(out_ls), 2)
    assert n > 0
    got = out[:n].tobytes().decode().split("\n")[:-1]
    vf = OracleVariantFile(("\n".join(hdr) + "\n" + "\n".join(lines) + "\n").encode())
    want = [R.write_record(rec, float(quals[i]), 30.0, overwrite_qual=overwrite_qual,
                           blacklist_value=bl_strings[codes[i]].decode())[0] for i, rec in enumerate(vf)]
    assert got == want
    assert [out[out_ls[i]:out_ls[i + 1] - 1].tobytes().decode() for i in range(len(lines))] == want


def test_splicer_without_model_only_fills_pass():
    lines = ["c1\t1\t.\tA\tC\t5\t.\tX=1", "c1\t2\t.\tA\tC\t5\tq10\t."]
    text = ("\n".join(lines) + "\n").encode()
    buf = np.frombuffer(text, dtype=np.uint8)
    ls = np.concatenate(([0], np.cumsum([len(l) + 1 for l in lines]))).astype(np.int64)
    ri = _recinfo_for(lines)
    out = np.empty(256, dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    n = lib.load_library().ugvc_splice_records(p(buf), p(ls), p(ri), None, None, 2, 0, 0, None, None, None, None, 0,
                                               p(out), out.size, None, 1)
    assert out[:n].tobytes().decode() == "c1\t1\t.\tA\tC\t5\tPASS\tX=1\nc1\t2\t.\tA\tC\t5\tq10\t.\n"


@pytest.mark.parametrize("overwrite_qual", [False, True])
def test_splicer_recalibrate_genotype_matches_oracle(overwrite_qual):
    """--recalibrate_genotype writer rules (filter_variants_pipeline.py:203-215): GQ / PL / GT of the
    first sample, no TREE_SCORE, QUAL = gq when overwriting."""
    hdr = ["##fileformat=VCFv4.2", "##contig=<ID=c1,length=100>",
           "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS\tS2"]
    lines = ["c1\t1\t.\tA\tC\t5.5\tPASS\tX=1\tGT:AD:DP:GQ:PL\t0/1:3,4:7:50:10,0,90\t1/1:0,9:9:20:90,20,0",
             "c1\t2\t.\tA\tC\t.\t.\t.\tGT:PL\t1|1:0,0,0\t.",
             "c1\t3\trs\tAT\tA\t7\tq10\tY=2\tGT:DP\t0/1:3\t./.",
             "c1\t4\t.\tA\tC,G\t7\t.\tY=2\tGT:GQ:PL\t1/2:9:9,8,7,6,5,4\t0/0"]
    rng = np.random.default_rng(4)
    probs = rng.dirichlet(np.ones(3), size=len(lines))
    probs[1] = [0.2, 0.2, 0.6]
    phreds, quals, gq = R.score_math(probs)
    low = (quals <= 30.0).astype(np.uint8)
    text = ("\n".join(lines) + "\n").encode()
    buf = np.frombuffer(text, dtype=np.uint8)
    ls = np.concatenate(([0], np.cumsum([len(l) + 1 for l in lines]))).astype(np.int64)
    ri = _recinfo_for(lines)
    out = np.empty(len(text) + 1024, dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    ph = np.ascontiguousarray(phreds)
    n = lib.load_library().ugvc_splice_records(p(buf), p(ls), p(ri), p(low), p(quals), len(lines), int(overwrite_qual),
                                               1, None, None, None, p(ph), 3, p(out), out.size, None, 2)
    assert n > 0
    got = out[:n].tobytes().decode().split("\n")[:-1]
    vf = OracleVariantFile(("\n".join(hdr) + "\n" + "\n".join(lines) + "\n").encode())
    want = [R.write_record(rec, float(quals[i]), 30.0, overwrite_qual=overwrite_qual, blacklist_value=None,
                           recal=(phreds[i], float(gq[i])))[0] for i, rec in enumerate(vf)]
    assert got == want
    assert all("TREE_SCORE" not in g for g in got) and got[3].split("\t")[8] == "GT:GQ:PL"


def test_output_pipeline_writer_thread_keeps_order_and_index(tmp_path):
    """_Splicer: splice on the caller's thread, BGZF deflate + index bookkeeping on its writer thread.
    Many small batches over three contigs -> the file holds every record in order and the .tbi selects
    each contig; a failing write surfaces on the caller's thread."""
    import gzip

    from variantcalling_b200 import filter_variants_pipeline as fvp
    from variantcalling_b200 import synth

    header, lines, _ = synth.generate(synth.SynthSpec(n_records=6000, n_custom=2, seed=9,
                                                      contigs={"chr1": 3_000_000, "chr2": 2_000_000, "chr3": 1_000_000}))
    path = str(tmp_path / "out.vcf.gz")
    sp = fvp._Splicer(path, 2)  # noqa: SLF001
    sp.write_header(header)
    rng = np.random.default_rng(0)
    want = []
    at = 0
    while at < len(lines):
        contig = lines[at].split("\t", 1)[0]
        stop = at
        while stop < len(lines) and stop - at < 700 and lines[stop].split("\t", 1)[0] == contig:
            stop += 1
        chunk = lines[at:stop]
        text = np.frombuffer(("\n".join(chunk) + "\n").encode(), dtype=np.uint8)
        quals = rng.uniform(0, 60, size=len(chunk))
        res = {"n_records": len(chunk), "recinfo": _recinfo_for(chunk), "low_score": (quals <= 30.0).astype(np.uint8),
               "qual": quals, "line_start": np.concatenate(([0], np.cumsum([len(x) + 1 for x in chunk]))).astype(np.int64)}
        sp.write_batch(contig, text, res, with_model=True, overwrite_qual=False, bl_code=None, bl_table=b"", bl_off=None)
        vf = OracleVariantFile(("\n".join(header) + "\n" + "\n".join(chunk) + "\n").encode())
        want += [R.write_record(rec, float(quals[i]), 30.0, overwrite_qual=False, blacklist_value=None)[0]
                 for i, rec in enumerate(vf)]
        at = stop
    sp.close(path)
    body = [ln for ln in gzip.open(path).read().decode().split("\n")[:-1] if not ln.startswith("#")]
    assert body == want
    idx = bgzf_io.read_tbi(path + ".tbi")
    assert list(idx) == ["chr1", "chr2", "chr3"]
    for c, (vb, ve) in idx.items():
        got = bgzf_io.inflate(path, vb, ve).tobytes().decode().split("\n")[:-1]
        assert got == [ln for ln in want if ln.split("\t", 1)[0] == c]
    assert sp.seconds["deflate"] > 0 and not sp._thread.is_alive()  # noqa: SLF001

    # an error on the writer thread is raised to the caller, and abort() stops the thread
    bad = fvp._Splicer(str(tmp_path / "no_such_dir" / "x.vcf.gz"), 1)  # noqa: SLF001
    chunk = lines[:10]
    text = np.frombuffer(("\n".join(chunk) + "\n").encode(), dtype=np.uint8)
    res = {"n_records": 10, "recinfo": _recinfo_for(chunk), "low_score": np.zeros(10, np.uint8), "qual": np.ones(10),
           "line_start": np.concatenate(([0], np.cumsum([len(x) + 1 for x in chunk]))).astype(np.int64)}
    bad.write_batch("chr1", text, res, with_model=True, overwrite_qual=False, bl_code=None, bl_table=b"", bl_off=None)
    with pytest.raises(OSError):
        bad.close(str(tmp_path / "no_such_dir" / "x.vcf.gz"))
    bad.abort()
    assert not bad._thread.is_alive()  # noqa: SLF001


def test_count_lines_equals_numpy():
    rng = np.random.default_rng(4)
    for n in (0, 1, 5, 1 << 20, (3 << 20) + 17):
        a = rng.integers(0, 32, size=n, dtype=np.uint8)
        for threads in (1, 3, 0):
            assert bgzf_io.count_lines(a, threads) == int(np.count_nonzero(a == 10))
    assert bgzf_io.count_lines(np.frombuffer(b"a\nb\n\n", dtype=np.uint8)) == 3
    assert bgzf_io.count_lines(np.arange(100, dtype=np.uint8)[::2]) == 1  # non-contiguous views are copied first


def test_info_end_positions_follow_htslib():
    """The index end of a record is INFO/END when the tag is there and beyond POS (htslib's rlen for VCF), found in
    the INFO column only."""
    from variantcalling_b200.filter_variants_pipeline import info_end_positions
    from variantcalling_b200.lib import RECINFO_DTYPE

    lines = [
        "chr1\t100\t.\tA\t<DEL>\t50\t.\tSVTYPE=DEL;END=5000;DP=3\tGT\t0/1",
        "chr1\t200\tEND=9\tA\tC\t50\t.\tDP=3\tGT\t0/1",                      # not in INFO
        "chr1\t300\t.\tA\t<NON_REF>\t.\t.\tEND=450\tGT\t0/0",
        "chr1\t400\t.\tAT\tA\t50\tPASS\tDP=1;XEND=7;END=350;END=9999\tGT:END=1\t0/1:5",  # END before POS: ignored by htslib
        "chr1\t500\t.\tA\tC\t50\t.\tDP=3;END=\tGT\t0/1",                      # no digits
        "chr1\t600\t.\tA\tC\t50\t.\tDP=3;END=700",                             # no FORMAT column
    ]
    text = np.frombuffer(("\n".join(lines) + "\n").encode(), dtype=np.uint8)
    n = len(lines)
    ls = np.zeros(n + 1, dtype=np.int64)
    ri = np.zeros(n, dtype=RECINFO_DTYPE)
    at = 0
    for i, ln in enumerate(lines):
        c = ln.split("\t")
        ls[i] = at
        ri["pos"][i] = int(c[1])
        ri["info_off"][i] = sum(len(x) + 1 for x in c[:7])
        ri["format_off"][i] = sum(len(x) + 1 for x in c[:8])
        at += len(ln) + 1
    ls[n] = at
    got = info_end_positions(text, ls, ri, n)
    assert got.tolist() == [5000, 0, 450, 350, 0, 700]
    beg = ri["pos"].astype(np.int64) - 1
    end = np.where(got > beg, got, beg + 1)
    assert end.tolist() == [5000, 200, 450, 400, 500, 700]
