"""GPU: BASELINE configs[4] in small -- filter a synthetic call set with the CLI, build the comparison
frame from the filtered VCF and the generator's truth labels (SURVEY.md 8d cfg 5: ``rtg vcfeval`` is not
available, the generator knows which calls are true), and compute the accuracy table on the GPU: the
``*.stats.csv`` rows must equal those the oracle computes from the oracle-filtered set."""
import gzip
import pickle

import numpy as np
import pandas as pd
import pytest

from oracle import concordance_ref as CR
from oracle import ref_pipeline as R
from tests import util
from variantcalling_b200 import bgzf_io
from variantcalling_b200 import concordance as PC
from variantcalling_b200 import filter_variants_pipeline as fvp

pytestmark = pytest.mark.gpu


def comparison_frame(record_lines, labels, n_missed, seed=0):
    """filter / tree_score / classify / indel / hmer_indel_length of the calls, plus ``n_missed`` truth
    variants that the call set lacks (false negatives: no score, no filter)."""
    cols = [ln.split("\t") for ln in record_lines]
    info = [dict(kv.partition("=")[::2] for kv in c[7].split(";")) for c in cols]
    indel = np.array([len({len(a) for a in [c[3], *c[4].split(",")]}) > 1 for c in cols])
    hil = np.array([0 if d.get("X_HIL", ".").split(",")[0] in (".", "") else int(d["X_HIL"].split(",")[0]) for d in info])
    frame = pd.DataFrame({
        "filter": [c[6] for c in cols],
        "tree_score": [float(np.float32(float(d["TREE_SCORE"]))) for d in info],  # what pysam reads back (float32)
        "classify": np.where(np.asarray(labels) == 1, "tp", "fp"), "indel": indel,
        "hmer_indel_length": np.where(indel, hil, 0), "hmer_indel_nuc": None})
    rng = np.random.default_rng(seed)
    missed = pd.DataFrame({"filter": "PASS", "tree_score": np.nan, "classify": "fn", "indel": rng.random(n_missed) < 0.4,
                           "hmer_indel_length": 0, "hmer_indel_nuc": None})
    out = pd.concat((frame, missed), ignore_index=True)
    out["classify_gt"] = out["classify"]
    return out


def test_stats_rows_of_gpu_filtered_and_oracle_filtered_sets_are_identical(tmp_path):
    ds = util.make_dataset(n_records=12000, n_custom=4, seed=55, region=("chr1", 1, 5_000_000))
    _, tr, x = util.fit_transformer(ds)
    model = util.fit_model("gb_small", x, ds["labels"])
    vcf, mpath, out = str(tmp_path / "in.vcf.gz"), str(tmp_path / "m.pkl"), str(tmp_path / "out.vcf.gz")
    bgzf_io.write_vcf_gz(vcf, ds["header"], ds["lines"])
    with open(mpath, "wb") as fh:
        pickle.dump({"xgb": model, "transformer": tr}, fh)
    argv = ["--input_file", vcf, "--model_file", mpath, "--output_file", out]
    for c in ds["customs"]:
        argv += ["--custom_annotations", c]
    fvp.run(argv)
    got_lines = [ln for ln in gzip.open(out).read().decode().split("\n")[:-1] if not ln.startswith("#")]
    want_lines = R.filter_variants(ds["vf"], model, tr, custom_annotations=ds["customs"])["lines"]
    gpu_frame = comparison_frame(got_lines, ds["labels"], n_missed=300)
    ref_frame = comparison_frame(want_lines, ds["labels"], n_missed=300)
    stats = PC.calc_accuracy_metrics(gpu_frame, "classify_gt")
    want = CR.calc_accuracy_metrics(ref_frame, "classify_gt")
    assert list(stats["group"]) == list(want["group"])
    for col in CR.METRIC_COLUMNS:
        np.testing.assert_array_equal(stats[col].to_numpy(dtype=float), want[col].to_numpy(dtype=float), err_msg=col)
    assert stats.loc[0, "tp"] > 1000 and stats.loc[0, "fn"] > 100 and 0 < stats.loc[0, "precision"] < 1
    # the csv the tool writes (';'-separated, evaluate_concordance.py:100-101) carries the same rows
    prefix = str(tmp_path / "eval")
    gpu_frame.to_csv(prefix + ".csv", index=False)
    from variantcalling_b200 import evaluate_concordance as EC

    EC.run(["--input_file", prefix + ".csv", "--output_prefix", prefix])
    csv = pd.read_csv(prefix + ".stats.csv", sep=";")
    for col in CR.METRIC_COLUMNS:
        np.testing.assert_allclose(csv[col].to_numpy(dtype=float), want[col].to_numpy(dtype=float), rtol=0, atol=1e-12)
