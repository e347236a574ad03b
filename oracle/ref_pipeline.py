"""ORACLE (test infrastructure, not product code).

CPU restatement of the reference's ``filter_variants_pipeline`` hot path, in
the reference's own shape (per-record Python dicts -> object DataFrame ->
sklearn ``ColumnTransformer`` -> ``predict_proba`` -> numpy fp64 score math ->
per-record FILTER / TREE_SCORE edits).  It exists to check the CUDA path and to
be timed as the CPU baseline; nothing in ``variantcalling_b200/`` imports it.

Follows (reference file:line):
  loader            ugbio_utils/src/core/ugbio_core/vcfbed/vcftools.py:16-217
  apply_model       ugbio_utils/src/filtering/ugbio_filtering/variant_filtering_utils.py:95-143
  phred             ugbio_utils/src/core/ugbio_core/math_utils.py:28-44
  score math        ugbio_utils/src/filtering/ugbio_filtering/filter_variants_pipeline.py:170-180
  header edits      .../filter_variants_pipeline.py:106-113
  writer rules      .../filter_variants_pipeline.py:188-228
  blacklists        ugbio_utils/src/filtering/ugbio_filtering/blacklist.py:36-101
  get_gt_from_pl_idx ugbio_utils/src/filtering/ugbio_filtering/multiallelics.py:257-277

PARITY STATUS: pinned for the transformer / phred / blacklist / PL-index
arithmetic (reference in-code known answers + golden vectors generated from
the reference's own ``transformers`` module, see ``tests/golden``); *unpinned*
for the pysam typed-decode step (``oracle/vcf_reader.py``) and for xgboost
inference -- those third-party engines are absent here and their reference
tests are git-LFS stubs.
"""
from __future__ import annotations

import time
from collections import defaultdict

import numpy as np
import pandas as pd

from oracle.vcf_reader import OracleHeader, OracleRecord, OracleVariantFile

MAX_CHUNK_SIZE = 1_000_000  # variant_filtering_utils.py:18

# vcftools.py:90-189 -- the loader's tag whitelist, in its order
LOADER_COLUMNS = (
    "GT PL DP AD MQ MMQ SOR AF DP_R DP_F AD_R AD_F TLOD VAF STRANDQ FPR GROUP TREE_SCORE VARIANT_TYPE DB "
    "AS_SOR AS_SORP FS VQR_VAL QD hiConfDeNovo loConfDeNovo GQ PGT PID PS AC AN BaseQRankSum ExcessHet "
    "MLEAC MLEAF MQRankSum ReadPosRankSum XC ID GNOMAD_AF gnomad.AF NLOD NALOD X_IC X_IL X_HIL X_HIN X_LM "
    "X_RM X_GCC X_CSS RPA RU STR AVERAGE_TREE_SCORE VQSLOD BLACKLST SCORE CALL BASE TVAF HighConfidence "
    "BG_AD BG_DP BG_SB DP4 INDEL IDV IMF VDB RPBZ MQBZ BQBZ MQSBZ NM SCBZ SGB MQ0F MinDP ADF ADR GP SYNC "
    "ML_PROB ASSEMBLED_HAPLOTYPES EXOME FILTERED_HAPS HAPCOMP HAPDOM HEC SB MQ0C SCL SCR NMC AFR"
).split()


def get_vcf_df(variant_file: OracleVariantFile, chromosome: str | None = None,
               custom_info_fields: list[str] | None = None) -> pd.DataFrame:
    """VCF -> object DataFrame.  vcftools.py:16-217 (sample_id=0, no scoring_field)."""
    header: OracleHeader = variant_file.header
    custom_info_fields = list(custom_info_fields or [])
    rows_iter = (
        defaultdict(
            lambda: None,
            list(x.info.items())
            + list(x.sample.items())
            + [
                ("QUAL", x.qual),
                ("CHROM", x.chrom),
                ("POS", x.pos),
                ("REF", x.ref),
                ("ID", x.id),
                ("ALLELES", x.alleles),
                ("FILTER", ";".join(str(y) for y in x.filter_keys)),
            ],
        )
        for x in variant_file.fetch(chromosome)
    )
    columns = list(LOADER_COLUMNS)
    for cf in custom_info_fields:
        if cf not in columns:
            columns.append(cf)
    known = list(header.info.keys()) + list(header.formats.keys())
    columns = [c for c in columns if c in known] + ["CHROM", "POS", "QUAL", "REF", "ALLELES", "FILTER", "ID"]
    # pandas-3 harness shim: keep str/None columns as object like pandas<3 did
    with pd.option_context("future.infer_string", False):
        df = pd.DataFrame([[r[c] for c in columns] for r in rows_iter], columns=[c.lower() for c in columns])
    if df.shape[0] == 0:
        return df
    df["indel"] = df["alleles"].apply(lambda a: len({len(y) for y in a}) > 1)
    df.index = pd.Index([(row[1]["chrom"], row[1]["pos"]) for row in df.iterrows()])
    if not df.columns.is_unique:
        raise ValueError("VCF columns are not unique")
    return df


def validate_data(data) -> None:
    """variant_filtering_utils.py:128-143 -- raise AssertionError on any null."""
    arr = data if isinstance(data, np.ndarray) else pd.DataFrame(data).to_numpy()
    if arr.ndim == 1 or arr.shape[1] <= 1:
        assert pd.isna(arr).sum() == 0, "data vector contains null"  # noqa: S101
    else:
        for c in range(arr.shape[1]):
            assert pd.isna(arr[:, c]).sum() == 0, f"Data matrix contains null in column {c}"  # noqa: S101


def harness_float_columns(df: pd.DataFrame) -> pd.DataFrame:
    """Harness shim for scikit-learn >= 1.6 (the reference pins 1.5.x): SimpleImputer now
    refuses to transform an int64 column when it was fitted on the same column as float64
    (a contig without any missing DP has an int column, the training frame had NaNs).  Integer
    columns are presented as float64 at both fit and transform; values are unchanged."""
    ints = [c for c in df.columns if pd.api.types.is_integer_dtype(df[c].dtype)]
    return df.astype({c: np.float64 for c in ints}) if ints else df


def transform_features(df: pd.DataFrame, transformer) -> pd.DataFrame:
    """Chunked ``transformer.transform``.  variant_filtering_utils.py:116-122."""
    df = harness_float_columns(df)
    bounds = np.concatenate((np.arange(0, df.shape[0], MAX_CHUNK_SIZE, dtype=int), [df.shape[0]]))
    with pd.option_context("future.infer_string", False):
        parts = [transformer.transform(df.iloc[bounds[i]: bounds[i + 1]]) for i in range(len(bounds) - 1)]
        x = pd.concat(parts)
    validate_data(x)
    return x


def apply_model(df: pd.DataFrame, model, transformer) -> tuple[np.ndarray, np.ndarray]:
    """variant_filtering_utils.py:95-125 (``predict`` and ``predict_proba`` both run)."""
    x = transform_features(df, transformer)
    x_in = x.to_numpy() if not hasattr(model, "feature_names_in_") else x
    predictions = model.predict(x_in)
    probabilities = model.predict_proba(x_in)
    return predictions, probabilities


def phred(p) -> np.ndarray:
    """math_utils.py:28-44."""
    return -10 * np.log10(np.array(p, dtype=float))


def score_math(scores: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(phreds, quals, gq) -- filter_variants_pipeline.py:168-180, incl. the per-row fill loop."""
    ml_lik = [list(x) for x in scores]
    likelihoods = np.zeros((len(ml_lik), max(len(r) for r in ml_lik)))
    for i, r in enumerate(ml_lik):
        likelihoods[i, : len(r)] = r
    phreds = phred(likelihoods + 1e-10)
    quals = np.clip(30 + phreds[:, 0] - np.min(phreds[:, 1:], axis=1), 0, None)
    order = np.argsort(phreds, axis=1)
    rows = np.arange(phreds.shape[0])
    gq = phreds[rows, order[:, 1]] - phreds[rows, order[:, 0]]
    return phreds, quals, gq


def blacklist_cg_insertions(df: pd.DataFrame) -> pd.Series:
    """blacklist.py:85-101 (tuple membership: an allele *equal* to GGC / CCG)."""
    hit = df["alleles"].apply(lambda a: "GGC" in a or "CCG" in a)
    return pd.Series("PASS", dtype=str, index=df.index).where(~hit, "CG_NON_HMER_INDEL")


def merge_blacklists(blacklists: list):
    """blacklist.py:61-82."""
    if len(blacklists) == 0:
        return None
    if len(blacklists) == 1:
        return blacklists[0]
    return blacklists[0].str.cat(blacklists[1:], sep=";", na_rep="PASS")


def apply_position_blacklist(df: pd.DataFrame, positions: set, annotation: str) -> pd.Series:
    """blacklist.py:36-55 with ``selection_fcn = ALL``."""
    common = set(df.index) & positions
    result = pd.Series("PASS", index=df.index, dtype=str)
    result.loc[list(common)] = annotation
    return result


def get_gt_from_pl_idx(idx: int) -> tuple:
    """multiallelics.py:257-277: triangular PL index -> (a, b)."""
    count, n_alleles = 0, 0
    while count < idx + 1:
        count += n_alleles
        n_alleles += 1
    max_allele = n_alleles - 1
    min_allele = idx - (count - n_alleles)
    return (min_allele - 1, max_allele - 1)


def format_float_g(value: float) -> str:
    """htslib prints INFO/QUAL floats (stored as float32) like ``%g``."""
    return "%g" % float(np.float32(value))


def edited_header_lines(header_lines: list[str], *, with_model: bool, with_blacklist: bool) -> list[str]:
    """filter_variants_pipeline.py:106-113 -- add FILTER/INFO meta lines if absent."""
    hdr = OracleHeader(header_lines)
    add = []
    if with_model and "LOW_SCORE" not in hdr.filters:
        add.append('##FILTER=<ID=LOW_SCORE,Description="Low decision tree score">')
    if with_blacklist and "BLACKLST" not in hdr.info:
        add.append('##INFO=<ID=BLACKLST,Number=.,Type=String,Description="blacklist">')
    if with_model and "TREE_SCORE" not in hdr.info:
        add.append('##INFO=<ID=TREE_SCORE,Number=1,Type=Float,Description="Filtering score">')
    out = list(header_lines)
    if not any(ln.startswith("##FILTER=<ID=PASS,") for ln in out):  # (OracleHeader.filters, like pysam's, always lists PASS)
        # htslib's header always carries the PASS filter (bcf_hdr_parse adds the line first, right after
        # ##fileformat); pysam.VariantFile(out, "w", header=hdr) writes it even when the input lacked it
        at = 1 if out and out[0].startswith("##fileformat") else 0
        out.insert(at, '##FILTER=<ID=PASS,Description="All filters passed">')
    chrom_at = next(i for i, ln in enumerate(out) if ln.startswith("#CHROM"))
    return out[:chrom_at] + add + out[chrom_at:]


def recalibrated_sample_columns(cols: list[str], phreds_row, gq: float, n_alleles: int) -> list[str]:
    """GT / GQ / PL of the first sample, filter_variants_pipeline.py:203-215 (pysam setters:
    existing FORMAT keys keep their place, new ones are appended; phasing is kept)."""
    keys = [] if cols[8] == "." else cols[8].split(":")
    vals = cols[9].split(":")
    vals += ["."] * (len(keys) - len(vals))
    n_pl = (n_alleles + 1) * n_alleles // 2
    pl = [int(x) for x in phreds_row[:n_pl]]
    gt = get_gt_from_pl_idx(int(np.argmin(pl)))
    sep = "|" if "GT" in keys and "|" in vals[keys.index("GT")] else "/"

    def put(key, val):
        if key in keys:
            vals[keys.index(key)] = val
        else:
            keys.append(key)
            vals.append(val)

    put("GQ", str(int(gq)))
    put("PL", ",".join(str(v) for v in pl))
    put("GT", sep.join(str(a) for a in gt))
    return cols[:8] + [":".join(keys), ":".join(vals)] + cols[10:]


def write_record(rec: OracleRecord, qual: float | None, threshold: float, *, overwrite_qual: bool,
                 blacklist_value: str | None, recal: tuple | None = None) -> tuple[str, list[str]]:
    """Apply filter_variants_pipeline.py:188-228 to one record; returns (line, filter keys).

    Untouched columns keep their input bytes (the product splices rather than
    re-serialising through htslib); FILTER, TREE_SCORE, QUAL (optional) and
    BLACKLST follow the reference rules.
    """
    cols = rec.line.rstrip("\n").split("\t")
    keys = list(rec.filter_keys)
    info = [] if cols[7] == "." else [kv for kv in cols[7].split(";") if kv]
    if qual is not None:
        if qual <= threshold:
            if "PASS" in keys:
                keys.remove("PASS")
            if "LOW_SCORE" not in keys:
                keys.append("LOW_SCORE")
        if recal is None:
            score_txt = "TREE_SCORE=" + format_float_g(qual)
            for i, kv in enumerate(info):
                if kv.split("=", 1)[0] == "TREE_SCORE":
                    info[i] = score_txt
                    break
            else:
                info.append(score_txt)
            if overwrite_qual:
                cols[5] = format_float_g(qual)
        else:  # --recalibrate_genotype: GQ / PL / GT instead of TREE_SCORE, QUAL = gq
            phreds_row, gq = recal
            if overwrite_qual:
                cols[5] = format_float_g(gq)
            cols = recalibrated_sample_columns(cols, phreds_row, gq, len(rec.alleles))
    if blacklist_value is not None and blacklist_value != "PASS":
        vals = [v for v in blacklist_value.split(";") if v != "PASS"]
        if vals:
            bl_txt = "BLACKLST=" + ",".join(vals)
            for i, kv in enumerate(info):
                if kv.split("=", 1)[0] == "BLACKLST":
                    info[i] = bl_txt
                    break
            else:
                info.append(bl_txt)
    if len(keys) == 0:
        keys.append("PASS")
    cols[6] = ";".join(keys)
    cols[7] = ";".join(info) if info else "."
    return "\t".join(cols), keys


def filter_variants(vcf: OracleVariantFile, model, transformer, *, custom_annotations=None,
                    decision_threshold: float = 30.0, blacklist_cg: bool = False,
                    position_blacklists: list | None = None, overwrite_qual_tag: bool = False,
                    limit_to_contigs: list[str] | None = None, timings: dict | None = None,
                    recalibrate_genotype: bool = False, treat_multiallelics: bool = False,
                    ref_fasta: dict | None = None) -> dict:
    """The serial contig loop of filter_variants_pipeline.py:116-229.

    ``treat_multiallelics`` takes the :145-166 branch (``ref_fasta``: contig -> sequence; the
    features / probs returned are then those of the *split* frame, in its row order).

    Returns dict(header=[...], lines=[...], filters=[...], quals=ndarray,
    probs=ndarray, features=ndarray).
    """
    if treat_multiallelics and ref_fasta is None:
        raise ValueError("Reference FASTA file is required for multiallelic treatment")
    timings = timings if timings is not None else {}
    out_lines, out_filters, all_quals, all_probs, all_feats = [], [], [], [], []
    with_bl = blacklist_cg or bool(position_blacklists)
    header = edited_header_lines(vcf.header_lines, with_model=model is not None, with_blacklist=with_bl)
    contigs = limit_to_contigs if limit_to_contigs is not None else list(vcf.header.contigs.keys())

    def tick(key, t0):
        timings[key] = timings.get(key, 0.0) + (time.perf_counter() - t0)

    for contig in contigs:
        t0 = time.perf_counter()
        df = get_vcf_df(vcf, chromosome=str(contig), custom_info_fields=custom_annotations)
        tick("parse", t0)
        if df.shape[0] == 0:
            continue
        if position_blacklists:
            applied = [apply_position_blacklist(df, s, name) for (s, name) in position_blacklists]
            blacklist = merge_blacklists(applied)
        else:
            blacklist = pd.Series("PASS", index=df.index, dtype=str)
        if blacklist_cg:
            blacklist = merge_blacklists([blacklist_cg_insertions(df), blacklist])
        quals = phreds = gq = None
        if model is not None:
            df_original = None
            if treat_multiallelics:
                from oracle import multiallelic_ref as MR

                df_original = df.copy()
                df = MR.process_multiallelic_spandel(df, ref_fasta[str(contig)], vcf.header)
            t0 = time.perf_counter()
            x = transform_features(df, transformer)
            tick("transform", t0)
            t0 = time.perf_counter()
            x_in = x.to_numpy() if not hasattr(model, "feature_names_in_") else x
            model.predict(x_in)  # the reference evaluates the model twice (predictions discarded)
            scores = model.predict_proba(x_in)
            tick("predict", t0)
            t0 = time.perf_counter()
            if treat_multiallelics:
                src = [i in df_original.index for i in df.index]
                dst = [i in df.index for i in df_original.index]
                df_original["ml_lik"] = pd.Series([list(r) for r in scores[src, :]], index=df_original.loc[dst].index)
                df_original = MR.combine_multiallelic_spandel(df, df_original, scores)
                phreds, quals, gq = score_math(list(df_original["ml_lik"]))
            else:
                phreds, quals, gq = score_math(scores)
            tick("score", t0)
            all_quals.append(quals)
            all_probs.append(np.asarray(scores, dtype=np.float64))
            all_feats.append(x.to_numpy(dtype=np.float64))
        t0 = time.perf_counter()
        bl_values = list(blacklist) if with_bl else None
        for i, rec in enumerate(vcf.fetch(str(contig))):  # PARSE #2, like the reference's fetch iterator
            line, keys = write_record(
                rec, None if quals is None else float(quals[i]), decision_threshold,
                overwrite_qual=overwrite_qual_tag, blacklist_value=None if bl_values is None else bl_values[i],
                recal=(phreds[i], float(gq[i])) if (recalibrate_genotype and quals is not None) else None)
            out_lines.append(line)
            out_filters.append(";".join(keys))
        tick("write", t0)
    return {
        "header": header,
        "lines": out_lines,
        "filters": out_filters,
        "quals": np.concatenate(all_quals) if all_quals else np.zeros(0),
        "probs": np.concatenate(all_probs) if all_probs else np.zeros((0, 2)),
        "features": np.concatenate(all_feats) if all_feats else np.zeros((0, 0)),
    }
