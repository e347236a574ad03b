"""GPU: --treat_multiallelics (and with --recalibrate_genotype) through the drop-in CLI against the
oracle's restatement of filter_variants_pipeline.py:145-228 (itself pinned against the reference's
own multi-allelic code, tests/test_multiallelics_cpu.py): every output record identical."""
import gzip
import pickle

import pytest

from oracle import ref_pipeline as R
from tests import multiallelic_data as MD
from tests import util
from variantcalling_b200 import bgzf_io
from variantcalling_b200 import filter_variants_pipeline as fvp

pytestmark = pytest.mark.gpu


def read_out(path):
    text = gzip.open(path).read().decode().split("\n")[:-1]
    return [ln for ln in text if ln.startswith("#")], [ln for ln in text if not ln.startswith("#")]


@pytest.fixture(scope="module", params=["rf", "lr", "gb3"])
def job(request, tmp_path_factory):
    d = tmp_path_factory.mktemp("multi_" + request.param)
    ds, tr, model, _split = util.make_multiallelic_case(21, request.param)
    vcf = str(d / "in.vcf.gz")
    bgzf_io.write_vcf_gz(vcf, ds["header"], ds["lines"])
    fasta = str(d / "ref.fa")
    with open(fasta, "w") as fh:
        fh.write(MD.fasta_text(ds["ref"]))
    mpath = str(d / "model.pkl")
    with open(mpath, "wb") as fh:
        pickle.dump({"transformer": tr, "xgb": model}, fh)
    return dict(dir=d, vcf=vcf, fasta=fasta, model=mpath, ds=ds, tr=tr, model_obj=model)


def argv_for(job, out, *extra):
    argv = ["--input_file", job["vcf"], "--model_file", job["model"], "--output_file", out, "--treat_multiallelics",
            "--ref_fasta", job["fasta"], *extra]
    for c in job["ds"]["customs"]:
        argv += ["--custom_annotations", c]
    return argv


def compare(recs, exp):
    assert len(recs) == len(exp["lines"])
    bad = [i for i, (a, b) in enumerate(zip(recs, exp["lines"])) if a != b]
    assert not bad, f"{len(bad)} records differ, first: {recs[bad[0]]!r} vs {exp['lines'][bad[0]]!r}"


def test_tree_score_and_filter_of_merged_records(job):
    out = str(job["dir"] / "o1.vcf.gz")
    totals = fvp.run(argv_for(job, out, "--blacklist_cg_insertions"))
    exp = R.filter_variants(job["ds"]["vf"], job["model_obj"], job["tr"], custom_annotations=job["ds"]["customs"],
                            treat_multiallelics=True, ref_fasta=job["ds"]["ref"], blacklist_cg=True)
    hdr, recs = read_out(out)
    assert hdr == exp["header"]
    compare(recs, exp)
    assert totals["n_low_score"] == sum("LOW_SCORE" in f for f in exp["filters"])
    assert sum(r.split("\t")[4].count(",") > 0 for r in recs) > 100  # multi-allelic records are in the output, unsplit


def test_recalibrated_genotypes_with_six_and_ten_pls(job):
    out = str(job["dir"] / "o2.vcf.gz")
    fvp.run(argv_for(job, out, "--recalibrate_genotype", "--overwrite_qual_tag", "--decision_threshold", "20"))
    exp = R.filter_variants(job["ds"]["vf"], job["model_obj"], job["tr"], custom_annotations=job["ds"]["customs"],
                            treat_multiallelics=True, ref_fasta=job["ds"]["ref"], recalibrate_genotype=True,
                            overwrite_qual_tag=True, decision_threshold=20.0)
    _, recs = read_out(out)
    compare(recs, exp)
    n_pl = [len(r.split("\t")[9].split(":")[-1].split(",")) for r in recs]
    assert 6 in n_pl and 10 in n_pl  # noqa: PLR2004


def test_reference_error_contract(job):
    # --ref_fasta is mandatory (filter_variants_pipeline.py:99-100)
    with pytest.raises(ValueError, match="Reference FASTA"):
        fvp.run(["--input_file", job["vcf"], "--model_file", job["model"], "--output_file", str(job["dir"] / "x.vcf.gz"),
                 "--treat_multiallelics"])
    # a contig without multi-allelic sites makes the reference fail in pd.concat (training_prep.py:261)
    d = job["dir"]
    ds = job["ds"]
    plain = [ln for ln in ds["lines"] if ln.split("\t")[4].count(",") == 0 and "*" not in ln.split("\t")[4]]
    vcf = str(d / "plain.vcf.gz")
    bgzf_io.write_vcf_gz(vcf, ds["header"], plain)
    argv = argv_for(dict(job, vcf=vcf), str(d / "y.vcf.gz"))
    with pytest.raises(ValueError, match="No objects to concatenate"):
        fvp.run(argv)
