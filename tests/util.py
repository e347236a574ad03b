"""Shared builders for the tests: synthetic data -> oracle frame -> fitted transformer + model."""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd

from oracle import ref_pipeline as R
from oracle.vcf_reader import OracleVariantFile
from variantcalling_b200 import synth
from variantcalling_b200 import transformers as T
from variantcalling_b200.tprep_constants import VcfType

warnings.filterwarnings("ignore")


def make_dataset(n_records=3000, n_custom=5, seed=synth.DEFAULT_SEED, **kw):
    spec = synth.SynthSpec(n_records=n_records, n_custom=n_custom, seed=seed, **kw)
    header, lines, labels = synth.generate(spec)
    customs = synth.custom_annotation_names(n_custom)
    vf = OracleVariantFile(synth.vcf_text(header, lines))
    return dict(spec=spec, header=header, lines=lines, labels=labels, customs=customs, vf=vf,
                text=("\n".join(lines) + "\n").encode(), header_text="\n".join(header) + "\n")


def fit_transformer(ds):
    df = R.get_vcf_df(ds["vf"], None, ds["customs"])
    tr = T.get_transformer(VcfType.SINGLE_SAMPLE, [a.lower() for a in ds["customs"]])
    with pd.option_context("future.infer_string", False):
        x = tr.fit_transform(R.harness_float_columns(df)).to_numpy(dtype=np.float64)
    return df, tr, x


def fit_model(kind, x, y, seed=1984):
    from sklearn.ensemble import GradientBoostingClassifier, RandomForestClassifier
    from sklearn.linear_model import LogisticRegression

    np.random.seed(seed)
    if kind == "lr":
        m = LogisticRegression(max_iter=300)
    elif kind == "gb":
        # hyper-parameters of the reference's XGBClassifier (variant_filtering_utils.py:70-78)
        m = GradientBoostingClassifier(n_estimators=100, learning_rate=0.15, subsample=0.4, max_depth=6,
                                       random_state=0)
    elif kind == "gb_small":
        m = GradientBoostingClassifier(n_estimators=12, learning_rate=0.15, subsample=0.4, max_depth=4,
                                       random_state=0)
    elif kind == "rf":
        m = RandomForestClassifier(n_estimators=30, max_depth=6, random_state=0, n_jobs=1)
    elif kind == "gb3":
        m = GradientBoostingClassifier(n_estimators=10, learning_rate=0.15, max_depth=3, random_state=0)
    else:
        raise ValueError(kind)
    m.fit(x, y)
    return m


def make_multiallelic_case(seed=11, model_kind="rf", n_custom=3):
    """Data set with multi-allelic sites + spanning deletions (tests/multiallelic_data.py), a
    transformer fitted on the *split* frame (as the reference's training does, training_prep.py:220)
    and a 3-class genotype model (0/0, 0/1, 1/1) trained on the split rows."""
    from oracle import multiallelic_ref as MR
    from tests import multiallelic_data as MD

    ds = MD.generate(seed, n_custom=n_custom)
    ds["vf"] = OracleVariantFile(ds["header_text"].encode() + ds["text"])
    frames = []
    for contig in MD.CONTIGS:
        df = R.get_vcf_df(ds["vf"], contig, ds["customs"])
        frames.append(MR.process_multiallelic_spandel(df, ds["ref"][contig], ds["vf"].header))
    split = pd.concat(frames)
    tr = T.get_transformer(VcfType.SINGLE_SAMPLE, [a.lower() for a in ds["customs"]])
    with pd.option_context("future.infer_string", False):
        x = tr.fit_transform(R.harness_float_columns(split)).to_numpy(dtype=np.float64)
    y = np.array([sum(1 for g in gt if g) for gt in split["gt"]])
    # the split genotypes alone do not carry much signal: tie the label to a few annotations too
    rng = np.random.default_rng(seed)
    y = np.where(rng.random(len(y)) < 0.25, rng.integers(0, 3, size=len(y)), y)
    model = fit_model(model_kind, x, y)
    return ds, tr, model, split
