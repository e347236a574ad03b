"""ORACLE (test infrastructure, not product code): the ``--treat_multiallelics`` branch.

CPU restatement, on pandas objects like the reference, of
  * ``select_overlapping_variants``                ugbio_filtering/multiallelics.py:13-62
  * ``split_multiallelic_variants`` / ``extract_allele_subset_from_multiallelic``   :65-177
  * GT / PL / indel helpers                         :180-383
  * ``classify_hmer_indel_relative``                :385-465  (with the flow-space helpers it calls:
      ``get_reference_from_region``  ugbio_core/concordance/flow_based_concordance.py:519-535,
      ``apply_variants_to_reference`` :263-339, ``compare_haplotypes`` :342-415,
      ``generate_key_from_sequence``  ugbio_core/flow_format/flow_based_read.py:55-112)
  * ``cleanup_multiallelics``                       :503-559
  * ``split_multiallelic_variants_with_spandel``    ugbio_filtering/spandel.py:11-128
  * ``process_multiallelic_spandel``                ugbio_filtering/training_prep.py:226-287
  * ``header_record_number`` / ``subsample_to_alleles``  ugbio_core/vcfbed/vcftools.py:687-778
  * ``combine_multiallelic_spandel`` / ``merge_and_assign_pls``  ugbio_filtering/variant_filtering_utils.py:309-408

PARITY STATUS: pinned.  ``tests/golden/multiallelic_split_*.json`` hold the data frames and merged
likelihoods produced by the reference's own functions (scripts/make_golden_multiallelic.py imports
them from /root/reference with stand-ins for pysam/pyfaidx, which they only use to read header
Numbers and slice a chromosome string); tests/test_multiallelics_cpu.py compares this module with
them cell by cell.

The reference's failure modes are kept: a contig without any multi-allelic site raises ValueError
(``pd.concat`` of an empty list), one without a spanning-deletion cluster raises KeyError
('spanning_deletion'), a genotype without the selected alleles raises AssertionError.
"""
from __future__ import annotations

import itertools
import re

import numpy as np
import pandas as pd

SPAN_DEL = "*"
FLOW_ORDER = "TGCA"


# ---------------------------------------------------------------- small index helpers
def pl_index(pair) -> int:
    """Index of genotype (a, b) in the VCF triangular PL order (multiallelics.py:239-252)."""
    hi, lo = max(pair), min(pair)
    return hi * (hi + 1) // 2 + int(lo)


def pl_subset(pl, pair, *, normed=True) -> tuple:
    """PLs of the three genotypes over ``pair`` (multiallelics.py:280-308)."""
    a, b = tuple(pair)
    vals = [pl[pl_index(g)] for g in ((a, a), (a, b), (b, b))]
    if normed:
        low = min(vals)
        vals = [v - low for v in vals]
    return tuple(vals)


def gt_subset(gt, pair) -> tuple:
    """multiallelics.py:206-236"""
    first, second = pair[0] in gt, pair[1] in gt
    assert first or second, "One of the alleles should be present in the GT"  # noqa: S101
    if first and second:
        return (0, 1)
    return (0, 0) if first else (1, 1)


def _has_star(alleles, pair) -> bool:
    return SPAN_DEL in (alleles[pair[0]], alleles[pair[1]])


def indel_subset(alleles, pair, spandel=None) -> bool:
    """multiallelics.py:311-339"""
    if _has_star(alleles, pair):
        if spandel is None:
            raise RuntimeError("Can't deal with spanning deletion allele without the spandel")
        return True
    return any(len(alleles[x]) != len(alleles[pair[0]]) for x in pair)


def indel_class_subset(alleles, pair, spandel=None):
    """-> (x_ic tuple, x_il tuple), multiallelics.py:342-383"""
    if not indel_subset(alleles, pair, spandel):
        return (("NA",), (None,))
    if _has_star(alleles, pair):
        return (("del",), (spandel["x_il"][0],))
    r, a = alleles[pair[0]], alleles[pair[1]]
    return (("del",), (len(r) - len(a),)) if len(r) > len(a) else (("ins",), (len(a) - len(r),))


# ---------------------------------------------------------------- flow-space hmer comparison
def reference_window(ref, region) -> str:
    """[start, end) 1-based; upper case, anything but . A T C G becomes A  (flow_based_concordance.py:519-535)."""
    return re.sub(r"([^.ATCG])", "A", str(ref[region[0] - 1:region[1] - 1]).upper())


def apply_one_variant(window: str, alleles, gt, ref_allele: str, rel_pos: int) -> list:
    """Haplotypes of ``window`` with each allele of ``gt`` placed at ``rel_pos`` (0-based index in
    the window); the single-variant case of apply_variants_to_reference (flow_based_concordance.py:263-339)."""
    left, right = window[0:rel_pos], window[rel_pos:len(window)]
    right = right[len(ref_allele):]
    picked = [alleles[n] for n in sorted(g for g in gt if g is not None)]
    picked = [a for a in picked if not a.startswith("<") and "*" not in a]
    return [left + a + right for a in picked]


def flow_key(sequence: str, flow_order: str = FLOW_ORDER) -> np.ndarray:
    """Bases -> flow-space key (flow_based_read.py:55-112, truncate=None, non_standard_as_a=False)."""
    sequence = sequence.upper()
    if re.search(r"[^ACGT]", sequence):
        raise ValueError("Input contains non ACGTacgt characters" + (f":\n{sequence}" if len(sequence) <= 100 else ""))  # noqa: PLR2004
    key, at = [], 0
    for base in itertools.cycle(flow_order):
        if not sequence:
            break  # the reference's flow string is empty for an empty sequence -> empty key
        run = 0
        while at + run < len(sequence) and sequence[at + run] == base:
            run += 1
        key.append(run)
        at += run
        if at >= len(sequence):
            break
    return np.array(key)


def compare_pair(k0: np.ndarray, k1) -> int:
    """Number of differing flows for one pair; 100 when the keys differ in length or one side is
    missing (compare_haplotypes on single-element lists, flow_based_concordance.py:342-415)."""
    if k1 is None or len(k0) != len(k1):
        return 100
    return int((k0 - k1 != 0).sum())


def hmer_indel_relative(alleles, pair, ref, pos: int, spandel=None) -> tuple:
    """(nucleotide, length) if allele pair[1] is an hmer indel of pair[0], else ('.', 0)
    (multiallelics.py:385-465)."""
    window = reference_window(ref, (max(0, pos - 20), min(pos + 20, len(ref)))).upper()
    if _has_star(alleles, pair):
        if spandel is None:
            raise RuntimeError("when the alleles contain spanning deletion, the line containing the variant "
                               "that is a deletion is required")
        haps = apply_one_variant(window, alleles, [a for a in pair if alleles[a] != SPAN_DEL], alleles[0], 20)
        sp_alleles = spandel["alleles"][0:2]
        haps = haps + apply_one_variant(window, sp_alleles, (0, 1), sp_alleles[0], spandel["pos"] - (pos - 20))[1:2]
    else:
        haps = apply_one_variant(window, alleles, pair, alleles[0], 20)
    keys = [flow_key(h) for h in haps]
    if not keys or compare_pair(keys[0], keys[1] if len(keys) > 1 else None) != 1:
        return (".", 0)
    where = np.nonzero(keys[0] - keys[1])[0]
    return (FLOW_ORDER[where[0] % 4], max(int(keys[0][where[0]]), int(keys[1][where[0]])))


# ---------------------------------------------------------------- header numbers / subsetting
def header_numbers(header) -> dict:
    """tag (and lower-cased tag) -> VCF Number, with the reference's overrides (vcftools.py:687-742).
    ``header`` is an oracle.vcf_reader.OracleHeader."""
    out = {}
    for table in (header.info, header.formats):
        for tag, (number, _vtype) in table.items():
            n = int(number) if number.isdigit() else number
            out[tag] = n
            out[tag.lower()] = n
    out["alleles"], out["ref"] = "R", 1
    for low, up in (("hapcomp", "HAPCOMP"), ("hapdom", "HAPDOM")):
        if (low in out or up in out) and out[low] == "A":
            out[low] = out[up] = 1
    out["RPA"] = out["rpa"] = "R"
    out["RU"] = out["ru"] = 1
    out["STR"] = out["str"] = 1
    for col in ("gt_vcfeval", "alleles_vcfeval", "label", "sync"):
        out[col] = 1
    return out


def subsample(field: tuple, number, pair) -> tuple:
    """vcftools.py:745-778"""
    if number == "A":
        return tuple(field[i - 1] for i in pair[1:])
    if number == "R":
        return tuple(field[i] for i in pair)
    if number == "G":
        raise RuntimeError("Special treatment is required for 'G' fields, not supported by this function")
    if number == ".":
        return field
    raise RuntimeError(f"Number {number} is not supported")


def allele_pair_row(row: pd.Series, pair, numbers: dict, ref, spandel=None) -> pd.Series:
    """One biallelic row over ``pair`` (multiallelics.py:130-177, spandel.py:66-128)."""
    alleles, pos = row["alleles"], row["pos"]
    out = {}
    for col in row.index:
        val = row.at[col]
        if col == "sb":
            out[col] = val
        elif col == "pl":
            out[col] = pl_subset(val, pair)
        elif col == "gt":
            out[col] = gt_subset(val, pair)
        elif col == "ref":
            out[col] = alleles[pair[0]]
        elif col == "indel":
            out[col] = indel_subset(alleles, pair, spandel)
        elif col == "x_ic":
            out[col] = indel_class_subset(alleles, pair, spandel)[0]
        elif col == "x_il":
            out[col] = indel_class_subset(alleles, pair, spandel)[1]
        elif col == "x_hil":
            out[col] = (hmer_indel_relative(alleles, pair, ref, pos, spandel)[1],)
        elif col == "x_hin":
            out[col] = (hmer_indel_relative(alleles, pair, ref, pos, spandel)[0],)
        elif isinstance(val, tuple) and numbers[col] != 1:
            out[col] = subsample(val, numbers[col], pair)
        else:
            out[col] = val
    return pd.Series(out)


def split_multiallelic(row: pd.Series, numbers: dict, ref) -> pd.DataFrame:
    """multiallelics.py:65-127: rows (REF, strongest ALT) and (strongest ALT, second ALT)."""
    alleles = row["alleles"]
    hom_pl = np.array([pl_subset(row["pl"], (0, i), normed=False)[-1] for i in range(1, len(alleles))])
    not_called = np.array([i not in row["gt"] for i in range(1, len(alleles))])
    order = [x for x in np.argsort(hom_pl + not_called * 1000) + 1 if alleles[x] != SPAN_DEL]
    pairs = [(0, order[0])] if len(order) == 1 else [(0, order[0]), (order[0], order[1])]
    return pd.concat([allele_pair_row(row, p, numbers, ref) for p in pairs], axis=1).T


def split_with_spandel(row: pd.Series, deletion: pd.Series, numbers: dict, ref) -> pd.DataFrame:
    """spandel.py:11-63: as above, with the '*' allele forced to be the weakest."""
    alleles = row["alleles"]
    star = alleles.index(SPAN_DEL)
    order = np.argsort([pl_subset(row["pl"], (0, i), normed=False)[-1] + 100000 * (i == star)
                        for i in range(1, len(alleles))]) + 1
    pairs = ((0, order[0]), (order[0], order[1]))
    return pd.concat([allele_pair_row(row, p, numbers, ref, spandel=deletion) for p in pairs], axis=1).T


# ---------------------------------------------------------------- the two entry points
def overlapping_sets(df: pd.DataFrame) -> list:
    """multiallelics.py:13-62 (require_star_for_spandel=True)."""
    alleles, pos = list(df["alleles"]), list(df["pos"])
    multi = {i for i, a in enumerate(alleles) if len(a) > 2}  # noqa: PLR2004
    del_len = [max(len(a[0]) - len(y) for y in a) for a in alleles]
    span, cluster, found = 0, [], []
    for i in range(len(alleles)):
        if not cluster and del_len[i] == 0:
            continue
        if pos[i] > span:
            if len(cluster) > 1:
                multi -= set(cluster)
                found.append(list(cluster))
            cluster = []
        if not cluster or SPAN_DEL in alleles[i]:
            cluster.append(i)
        span = max(span, del_len[i] + pos[i])
    return sorted(found + [[m] for m in multi])


def cleanup(df: pd.DataFrame) -> pd.DataFrame:
    """multiallelics.py:503-559 (without the optional STR/RU/RPA fix-up when there is no 'str' column)."""
    df = df.copy()
    has_len = df["x_il"].apply(lambda x: x[0] is not None and x[0] != 0)  # on the tuples: None must stay None
    df.loc[(df["variant_type"] == "snp") & has_len, "variant_type"] = "non-h-indel"
    to_h = (df["variant_type"] == "non-h-indel") & df["x_hil"].apply(lambda x: x[0] is not None and x[0] > 0)
    df.loc[to_h, "variant_type"] = "h-indel"
    if "str" in df.columns:
        df.loc[to_h, "str"] = True
        df.loc[to_h, "ru"] = df.loc[to_h, "x_hin"].apply(lambda x: x[0])

        def lengths(v):
            ins_len = v["x_il"][0] or 0
            h = v["x_hil"][0]
            return (h, h + ins_len) if v["x_ic"][0] == "ins" else (h + ins_len, h)

        df.loc[to_h, "rpa"] = df.loc[to_h].apply(lengths, axis=1)
    no_hmer = df["x_hil"].apply(lambda x: x[0] is None or x[0] == 0)
    df.loc[(df["variant_type"] == "h-indel") & no_hmer, "variant_type"] = "non-h-indel"
    ordered = df["pl"].apply(sorted)
    df["gq"] = np.clip(ordered.apply(lambda x: x[1] - x[0]), 0, 99)
    df["qual"] = np.clip(df["pl"].apply(lambda x: min(x[1:])) - df["pl"].apply(lambda x: x[0]), 0, None)
    df["qd"] = df["qual"] / df["dp"]
    return df


def process_multiallelic_spandel(df: pd.DataFrame, ref, header) -> pd.DataFrame:
    """training_prep.py:226-287.  ``ref``: the chromosome sequence (str); ``header``: OracleHeader."""
    df = df.copy()
    dtypes = df.dtypes
    sets = overlapping_sets(df)
    numbers = header_numbers(header)
    groups = []
    for (m,) in (s for s in sets if len(s) == 1):
        g = split_multiallelic(df.iloc[m], numbers, ref)
        g.loc[:, "multiallelic_group"] = [(g.iloc[0]["chrom"], g.iloc[0]["pos"])] * g.shape[0]
        groups.append(g)
    mug = cleanup(pd.concat(groups, ignore_index=True))  # ValueError when there is no multi-allelic site
    clusters = [s for s in sets if len(s) > 1]
    if clusters:
        parts = []
        for c in clusters:
            head = df.iloc[c[0]:c[0] + 1]
            if len(head["alleles"].to_numpy()[0]) != 2:  # noqa: PLR2004
                head = split_multiallelic(head.iloc[0], numbers, ref)
            block = pd.concat([head] + [split_with_spandel(df.iloc[i], df.iloc[c[0]], numbers, ref) for i in c[1:]])
            block.loc[:, "spanning_deletion"] = [(df.iloc[c[0]]["chrom"], df.iloc[c[0]]["pos"])] * block.shape[0]
            parts.append(block)
        mug = pd.concat((mug, cleanup(pd.concat(parts, ignore_index=True))), ignore_index=True)
    else:
        mug = mug.reset_index()
    df = df.drop(df.index[sum(sets, [])], axis=0)
    return pd.concat((df, mug)).astype(dtypes)


def merge_split_scores(unsplit: pd.DataFrame, grouped, scores: pd.Series) -> pd.DataFrame:
    """variant_filtering_utils.py:346-408"""
    merged = []
    for key in grouped.groups:
        rows = grouped.get_group(key)
        labels = grouped.groups[key]
        if rows.shape[0] == 1:
            lik = scores[labels[0]]
        else:
            orig = unsplit.at[key, "alleles"]
            lik = np.zeros(len(orig) * (len(orig) + 1) // 2)
            i1, i2 = orig.index(rows.iloc[1]["alleles"][0]), orig.index(rows.iloc[1]["alleles"][1])
            s0, s1 = scores[labels[0]], np.array(scores[labels[1]])
            vals = np.concatenate((s0[:2], np.insert(s0[2] * s1, 1, 0)))
            lik[[pl_index(g) for g in ((0, 0), (0, i1), (i1, i1), (0, i2), (i1, i2), (i2, i2))]] = vals
        merged.append(lik)
    dest = list(grouped.groups.keys())
    unsplit.loc[dest, "ml_lik"] = pd.Series(merged, index=unsplit.loc[dest].index)
    return unsplit


def combine_multiallelic_spandel(split: pd.DataFrame, unsplit: pd.DataFrame, scores: np.ndarray) -> pd.DataFrame:
    """variant_filtering_utils.py:309-343"""
    for col, by in (("multiallelic_group", "multiallelic_group"), ("spanning_deletion", ["chrom", "pos"])):
        sel = ~pd.isna(split[col])
        part = split[sel]
        per_row = pd.Series([list(x) for x in scores[sel, :]], index=part.index)
        unsplit = merge_split_scores(unsplit, part.groupby(by), per_row)
    return unsplit
