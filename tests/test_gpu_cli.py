"""GPU: the drop-in CLI (`python ugvc filter_variants_pipeline ...`) end to end on a synthetic
bgzip'ed VCF (BASELINE.json configs[0], plumbing): output records equal the oracle's writer
output line for line, header edits match, the .tbi indexes the output."""
import gzip
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from oracle import ref_pipeline as R
from oracle.vcf_reader import OracleVariantFile
from tests import util
from variantcalling_b200 import bgzf_io
from variantcalling_b200 import filter_variants_pipeline as fvp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Blacklist:  # same attribute contract as ugbio_filtering.blacklist.Blacklist (blacklist.py:10-34)
    def __init__(self, blacklist, annotation, description=""):
        self.blacklist, self.annotation, self.description = blacklist, annotation, description
        self.selection_fcn = None  # VariantSelectionFunctions.ALL


@pytest.fixture(scope="module")
def job(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    # ~13 k records in chr1:1-5 Mb like the reference's system-test fixtures (SURVEY.md 8d cfg 1)
    ds = util.make_dataset(n_records=13000, n_custom=4, seed=5, region=("chr1", 1, 5_000_000))
    extra = util.make_dataset(n_records=800, n_custom=4, seed=6, contigs={"chr2": 242193529, "chr3": 198295559})
    lines = ds["lines"] + extra["lines"]
    header = ds["header"]
    ds["lines"], ds["vf"] = lines, OracleVariantFile(("\n".join(header) + "\n" + "\n".join(lines) + "\n").encode())
    ds["labels"] = np.concatenate((ds["labels"], extra["labels"]))
    _, tr, x = util.fit_transformer(ds)
    model = util.fit_model("gb_small", x, ds["labels"])
    vcf = str(d / "in.vcf.gz")
    bgzf_io.write_vcf_gz(vcf, header, lines)
    mpath = str(d / "model.pkl")
    with open(mpath, "wb") as fh:
        pickle.dump({"transformer": tr, "xgb": model, "xgb_recall_precision": None}, fh)
    pos = [int(l.split("\t")[1]) for l in ds["lines"][:4000:400]]
    bl = [Blacklist({("chr1", p) for p in pos[:6]}, "ILLUMINA_FP"), Blacklist({("chr1", pos[2]), ("chr9", 5)}, "COHORT_FP")]
    blpath = str(d / "bl.pkl")
    with open(blpath, "wb") as fh:
        pickle.dump(bl, fh)
    return dict(dir=d, vcf=vcf, model=mpath, bl=blpath, bl_objs=bl, ds=ds, tr=tr, model_obj=model)


def read_out(path):
    text = gzip.open(path).read().decode().split("\n")[:-1]
    return [l for l in text if l.startswith("#")], [l for l in text if not l.startswith("#")]


def test_cli_matches_oracle_with_model_blacklist_and_cg(job):
    out = str(job["dir"] / "out1.vcf.gz")
    argv = ["--input_file", job["vcf"], "--model_file", job["model"], "--output_file", out, "--blacklist", job["bl"],
            "--blacklist_cg_insertions"]
    for c in job["ds"]["customs"]:
        argv += ["--custom_annotations", c]
    totals = fvp.run(argv)
    bls = [(b.blacklist, b.annotation) for b in job["bl_objs"]]
    exp = R.filter_variants(job["ds"]["vf"], job["model_obj"], job["tr"], custom_annotations=job["ds"]["customs"],
                            blacklist_cg=True, position_blacklists=bls)
    hdr, recs = read_out(out)
    assert hdr == exp["header"]
    assert len(recs) == len(exp["lines"]) == totals["n_records"]
    bad = [i for i, (a, b) in enumerate(zip(recs, exp["lines"])) if a != b]
    assert not bad, f"{len(bad)} records differ, first: {recs[bad[0]]!r} vs {exp['lines'][bad[0]]!r}"
    assert sum("BLACKLST=" in r for r in recs) > 6
    # value counts like the reference's system test goldens (FILTER column)
    got_counts = {k: sum(r.split("\t")[6] == k for r in recs) for k in ("PASS", "LOW_SCORE")}
    want_counts = {k: sum(f == k for f in exp["filters"]) for k in ("PASS", "LOW_SCORE")}
    assert got_counts == want_counts and totals["n_low_score"] == sum("LOW_SCORE" in f for f in exp["filters"])
    # the index selects each contig of the output
    idx = bgzf_io.read_tbi(out + ".tbi")
    assert list(idx) == ["chr1", "chr2", "chr3"]
    for c, (vb, ve) in idx.items():
        got = bgzf_io.inflate(out, vb, ve).tobytes().decode().split("\n")[:-1]
        assert got == [r for r in recs if r.split("\t", 1)[0] == c]


def test_cli_overwrite_qual_threshold_and_limit_contigs(job):
    out = str(job["dir"] / "out2.vcf.gz")
    argv = ["--input_file", job["vcf"], "--model_file", job["model"], "--output_file", out, "--overwrite_qual_tag",
            "--decision_threshold", "12.5", "--limit_to_contigs", "chr3", "chr1", "chrNOPE"]
    for c in job["ds"]["customs"]:
        argv += ["--custom_annotations", c]
    fvp.run(argv)
    exp = R.filter_variants(job["ds"]["vf"], job["model_obj"], job["tr"], custom_annotations=job["ds"]["customs"],
                            decision_threshold=12.5, overwrite_qual_tag=True, limit_to_contigs=["chr3", "chr1", "chrNOPE"])
    _, recs = read_out(out)
    assert recs == exp["lines"] and recs[0].startswith("chr3\t")


def test_cli_without_model_fills_pass_and_marks_cg(job):
    out = str(job["dir"] / "out3.vcf.gz")
    fvp.run(["--input_file", job["vcf"], "--output_file", out, "--blacklist_cg_insertions"])
    exp = R.filter_variants(job["ds"]["vf"], None, None, blacklist_cg=True)
    hdr, recs = read_out(out)
    assert hdr == exp["header"] and recs == exp["lines"]
    assert not any("TREE_SCORE" in r for r in recs) and not any(r.split("\t")[6] == "." for r in recs)


def test_python_ugvc_entry_point_and_error_contract(job):
    out = str(job["dir"] / "out4.vcf.gz")
    cmd = [sys.executable, "ugvc", "filter_variants_pipeline", "--input_file", job["vcf"], "--model_file", job["model"],
           "--output_file", out]
    for c in job["ds"]["customs"]:
        cmd += ["--custom_annotations", c]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Variant filtering run: success" in r.stderr and os.path.exists(out + ".tbi")
    # missing index -> RuntimeError + "failed" line + non-zero exit (filter_variants_pipeline.py:101-104,235-240)
    lonely = str(job["dir"] / "lonely.vcf.gz")
    with open(job["vcf"], "rb") as a, open(lonely, "wb") as b:
        b.write(a.read())
    r = subprocess.run(cmd[:4] + [lonely] + cmd[5:], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "Variant filtering run: failed" in r.stderr and "does not exist" in r.stderr
    with pytest.raises(ValueError, match="Reference FASTA"):
        fvp.run(["--input_file", job["vcf"], "--output_file", out, "--treat_multiallelics"])


def test_cli_recalibrate_genotype_three_class_model(job):
    """--recalibrate_genotype with an exact-GT style 3-class model: GT / GQ / PL rewritten from the
    per-class phreds, no TREE_SCORE, QUAL = GQ when overwriting (filter_variants_pipeline.py:203-215)."""
    ds = job["ds"]
    _, tr, x = util.fit_transformer(ds)
    y3 = np.where(x[:, 2] > 0, 2, ds["labels"])  # hom-alt calls get class 2: labels 0 / 1 / 2
    model = util.fit_model("gb3", x, y3)
    assert len(model.classes_) == 3
    mpath = str(job["dir"] / "model3.pkl")
    with open(mpath, "wb") as fh:
        pickle.dump({"transformer": tr, "xgb": model}, fh)
    out = str(job["dir"] / "out5.vcf.gz")
    argv = ["--input_file", job["vcf"], "--model_file", mpath, "--output_file", out, "--recalibrate_genotype",
            "--overwrite_qual_tag"]
    for c in ds["customs"]:
        argv += ["--custom_annotations", c]
    fvp.run(argv)
    exp = R.filter_variants(ds["vf"], model, tr, custom_annotations=ds["customs"], recalibrate_genotype=True,
                            overwrite_qual_tag=True)
    _, recs = read_out(out)
    assert len(recs) == len(exp["lines"])
    bad = [i for i, (a, b) in enumerate(zip(recs, exp["lines"])) if a != b]
    # int() of a phred sits on an integer boundary for ~1 record in 10^5 when the device log10 and
    # NumPy's differ in the last ulp; everything else must match byte for byte
    assert len(bad) <= 1, f"{len(bad)} records differ, first: {recs[bad[0]]!r} vs {exp['lines'][bad[0]]!r}"
    assert not any("TREE_SCORE" in r for r in recs)
    gts = {}
    for r in recs:
        g = r.split("\t")[9].split(":")[0]
        gts[g] = gts.get(g, 0) + 1
    assert set(gts) <= {"0/0", "0/1", "1/1"} and len(gts) == 3


def test_cli_small_batches_long_lines_and_rescoring(job):
    """Several batches per contig (--batch_mb 1), a record longer than 64 KiB (saturated column
    offsets: the splicer recounts tabs), and re-filtering an already filtered file (TREE_SCORE /
    LOW_SCORE already declared and present: replaced in place, header lines not duplicated)."""
    ds = job["ds"]
    lines = list(ds["lines"][:6000])
    big = lines[100].split("\t")
    big[3] = "A" + "CGT" * 24000           # 72 kB REF allele
    big[4] = "A"
    lines[100] = "\t".join(big)
    vcf = str(job["dir"] / "in_long.vcf.gz")
    bgzf_io.write_vcf_gz(vcf, ds["header"], lines)
    out1 = str(job["dir"] / "out6.vcf.gz")
    argv = ["--input_file", vcf, "--model_file", job["model"], "--output_file", out1, "--batch_mb", "1"]
    for c in ds["customs"]:
        argv += ["--custom_annotations", c]
    fvp.run(argv)
    vf = OracleVariantFile(("\n".join(ds["header"]) + "\n" + "\n".join(lines) + "\n").encode())
    exp = R.filter_variants(vf, job["model_obj"], job["tr"], custom_annotations=ds["customs"])
    hdr1, recs1 = read_out(out1)
    assert hdr1 == exp["header"] and recs1 == exp["lines"]
    assert len(recs1[100]) > 72000 and "TREE_SCORE=" in recs1[100]
    # second pass over the filtered output with another threshold
    out2 = str(job["dir"] / "out7.vcf.gz")
    fvp.run(["--input_file", out1, "--model_file", job["model"], "--output_file", out2, "--decision_threshold", "5",
             "--batch_mb", "1"] + [a for c in ds["customs"] for a in ("--custom_annotations", c)])
    vf2 = OracleVariantFile(("\n".join(hdr1) + "\n" + "\n".join(recs1) + "\n").encode())
    exp2 = R.filter_variants(vf2, job["model_obj"], job["tr"], custom_annotations=ds["customs"], decision_threshold=5.0)
    hdr2, recs2 = read_out(out2)
    assert hdr2 == exp2["header"] == hdr1      # nothing added twice
    assert recs2 == exp2["lines"]
    assert all(r.count("TREE_SCORE=") == 1 for r in recs2)
    assert any(r.split("\t")[6] == "LOW_SCORE" for r in recs1) and sum("LOW_SCORE" in r.split("\t")[6] for r in recs2) >= \
        sum("LOW_SCORE" in r.split("\t")[6] for r in recs1)


def test_cli_device_groups_cut_contigs_at_index_entries(job):
    """--batch_mb 1: the device-side file path takes pieces of about 128 KiB of compressed input -- chr1 is cut at entries
    of the input's tabix linear index, chr2 and chr3 share a call -- and the output (records, index) is what one call
    per contig and the host writers produce."""
    ds = job["ds"]
    base = ["--input_file", job["vcf"], "--model_file", job["model"]] + [a for c in ds["customs"] for a in ("--custom_annotations", c)]
    outs = {}
    for tag, extra in (("pieces", ["--batch_mb", "1"]), ("whole", []), ("host", ["--host_io"])):
        outs[tag] = str(job["dir"] / f"out_groups_{tag}.vcf.gz")
        fvp.run(base + ["--output_file", outs[tag]] + extra)
    exp = R.filter_variants(ds["vf"], job["model_obj"], job["tr"], custom_annotations=ds["customs"])
    for tag, path in outs.items():
        hdr, recs = read_out(path)
        assert hdr == exp["header"] and recs == exp["lines"], tag
        idx = bgzf_io.read_tbi(path + ".tbi")
        assert list(idx) == ["chr1", "chr2", "chr3"], tag
        for contig, (vb, ve) in idx.items():
            got = bgzf_io.inflate(path, vb, ve).tobytes().decode().split("\n")[:-1]
            assert got == [l for l in exp["lines"] if l.split("\t", 1)[0] == contig], (tag, contig)
    # the pieces really were pieces: the linear index of the input offers cut points inside chr1
    _, lin = bgzf_io.read_tbi(job["vcf"] + ".tbi", linear=True)
    assert lin["chr1"].size > 4


def test_index_ends_records_at_info_end(job):
    """A header that declares INFO/END (gVCF blocks, symbolic alleles): htslib's tabix -- what `bcftools index -t`
    writes for the reference, filter_variants_pipeline.py:231 -- ends such a record at END, so a region query that
    overlaps only its tail finds it.  Read back with the independent spec reader of oracle/tabix_ref.py."""
    from oracle import tabix_ref as TR

    ds = job["ds"]
    header = [ln for ln in ds["header"] if not ln.startswith("#CHROM")]
    header += ['##INFO=<ID=END,Number=1,Type=Integer,Description="Stop position of the interval">', ds["header"][-1]]
    lines = list(ds["lines"])
    long_ones = {}
    for i in range(50, 6000, 1100):
        c = lines[i].split("\t")
        if c[0] != "chr1":
            continue
        end = int(c[1]) + 150_000 + i
        c[7] += f";END={end}"
        lines[i] = "\t".join(c)
        long_ones[i] = (int(c[1]), end)
    k = 75  # an END that is not beyond POS is ignored
    c = lines[k].split("\t")
    c[7] = "END=1;" + c[7]
    lines[k] = "\t".join(c)
    assert len(long_ones) >= 4  # noqa: PLR2004
    vcf, out = str(job["dir"] / "in_end.vcf.gz"), str(job["dir"] / "out_end.vcf.gz")
    bgzf_io.write_vcf_gz(vcf, header, lines)
    argv = ["--input_file", vcf, "--model_file", job["model"], "--output_file", out, "--host_io"]  # (the END-aware writer)
    for cu in ds["customs"]:
        argv += ["--custom_annotations", cu]
    fvp.run(argv)
    vf = OracleVariantFile(("\n".join(header) + "\n" + "\n".join(lines) + "\n").encode())
    exp = R.filter_variants(vf, job["model_obj"], job["tr"], custom_annotations=ds["customs"])
    _, recs = read_out(out)
    assert recs == exp["lines"]
    idx = TR.TabixIndex(out + ".tbi")
    for i, (pos, end) in long_ones.items():
        tail = TR.query(out, idx, "chr1", end - 10, end - 5)  # overlaps nothing but the tail of record i
        assert any(ln.split(b"\t")[1] == str(pos).encode() and f"END={end}".encode() in ln for ln in tail), (i, pos, end)
        assert not TR.query(out, idx, "chr1", end, end + 1) or all(
            int(ln.split(b"\t")[1]) - 1 < end + 1 for ln in TR.query(out, idx, "chr1", end, end + 1))
    # the record with END=1 keeps its REF-length end
    p75 = int(lines[k].split("\t")[1])
    assert any(int(ln.split(b"\t")[1]) == p75 for ln in TR.query(out, idx, "chr1", p75 - 1, p75))
