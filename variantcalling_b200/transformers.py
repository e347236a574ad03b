"""Host-side mirror of the reference feature-transformer interface.

This is the *descriptor* of the per-site feature vector: a fitted sklearn
``ColumnTransformer`` whose entries carry the reference's own names.  The
filtering hot path never calls ``.transform`` on it -- the model compiler
(``model_compiler.py``) reads the fitted object (entry names, column slices,
fitted ``OrdinalEncoder`` categories) and lowers it to the flat plan that the
CUDA kernels execute.  ``fit``/``fit_transform`` run on the CPU because model
training stays on the CPU in the reference too (``train_models_pipeline`` is
out of scope, its pickle contract is in scope).

Interface mirrored (same names, same argument meaning, same column order):
  reference ``ugbio_utils/src/filtering/ugbio_filtering/transformers.py``
    scalar encoders            :15-123
    per-VCF-type feature lists :144-347  (``modify_features_based_on_vcf_type``)
    ``get_transformer``        :350-369
    label encode / decode      :372-406

Every helper below is a module-level function (the reference nests them in a
closure and relies on ``dill``); module-level functions pickle with the stock
``pickle`` module, and the model compiler dispatches on ``func.__name__`` so a
genuine reference pickle lowers the same way.
"""
from __future__ import annotations

import itertools
from collections.abc import Iterable
from functools import lru_cache

import numpy as np
import pandas as pd
from sklearn import compose, impute, preprocessing
from sklearn.pipeline import make_pipeline

from variantcalling_b200.tprep_constants import VcfType

_BASE_CODE = {"A": 1, "T": 2, "G": 3, "C": 4}
_MOTIF_CODE = {"A": 1, "T": 2, "G": 3, "C": 4, "N": 5}
INS_DEL_ENCODE = {"ins": -1, "del": 1, "NA": 0}
_REGIONS = ("Telomere_Centromere", "Clusters", "Coverage-Mappability")


# --------------------------------------------------------------------------- scalar encoders
def tuple_break(x):
    """First element of a tuple (``None`` element -> NaN); a missing scalar -> 0.
    Reference: transformers.py:15-19."""
    if isinstance(x, tuple):
        head = x[0]
        return np.nan if head is None else head
    if x is None or np.isnan(x):
        return 0
    return x


def tuple_break_second(x):
    """Reference: transformers.py:22-26."""
    if isinstance(x, tuple) and len(x) > 1:
        return x[1]
    if x is None or (isinstance(x, tuple) and len(x) < 2) or np.isnan(x):  # noqa: PLR2004
        return 0
    return x


def tuple_break_third(x):
    """Reference: transformers.py:29-33 (note: guards on len > 1, indexes [2])."""
    if isinstance(x, tuple) and len(x) > 1:
        return x[2]
    if x is None or (isinstance(x, tuple) and len(x) < 2) or np.isnan(x):  # noqa: PLR2004
        return 0
    return x


def _motif_digits(items) -> int:
    acc = 0
    for item in items:
        acc = acc * 10 + _MOTIF_CODE.get(item, 0)
    return acc


def motif_encode_left(x):
    """Base-10 digit string of the motif read right-to-left (closest base most
    significant).  Reference: transformers.py:36-47."""
    return _motif_digits(reversed(list(x)))


def motif_encode_right(x):
    """Reference: transformers.py:50-60."""
    return _motif_digits(list(x))


def allele_encode(x):
    """Single base -> 1..4, anything else -> 0.  Reference: transformers.py:72-77."""
    try:
        return _BASE_CODE.get(x, 0)
    except TypeError:  # unhashable
        return 0


def svtype_encode(x):
    """Reference: transformers.py:80-83."""
    return {"DEL": 1, "DUP": 2, "NEUTRAL": 0}[x]


def cnv_source_encode(x):
    """Reference: transformers.py:86-91."""
    if len(x) != 1:
        raise ValueError(f"Unexpected cnv_source value: {x}")
    return {"cn.mops": 1, "cnvpytor": 2}[x[0]]


def gt_encode(x):
    """1 for a hom-alt ``(1, 1)`` genotype else 0.  Reference: transformers.py:94-98."""
    return 1 if x == (1, 1) else 0


def ins_del_encode(x, encode_dct=INS_DEL_ENCODE):  # pylint: disable=dangerous-default-value
    """Reference: transformers.py:101-105 (KeyError on an unknown class)."""
    return encode_dct[x]


@lru_cache(maxsize=1)
def _get_region_encoding():
    """Subset -> 1..8 code.  Reference: transformers.py:108-116."""
    ordered = sorted(_REGIONS)
    subsets = []
    for r in range(len(ordered) + 1):
        subsets.extend(tuple(sorted(c)) for c in itertools.combinations(ordered, r))
    return {s: i + 1 for i, s in enumerate(subsets)}


def region_annotation_encode(x):
    """Reference: transformers.py:119-123."""
    if x is None:
        return 0
    return _get_region_encoding()[tuple(sorted(x))]


# --------------------------------------------------------------------------- column encoders
def _column(values, index, width=1):
    return pd.DataFrame(np.array(values).reshape((-1, width)), index=index)


def _elementwise(df: pd.DataFrame, fn) -> pd.DataFrame:
    mapper = getattr(df, "map", None) or df.applymap
    return mapper(fn)


def tuple_encode_df(s):
    """Reference: transformers.py:165-166."""
    return _column([tuple_break(y) for y in s], s.index)


def tuple_encode_doublet_df(s):
    """First two elements of each tuple.  Reference: transformers.py:168-169."""
    return _column([y[:2] for y in s], s.index, 2)


def tuple_uniform_encode(s):
    """All elements, ragged rows padded with 1000.  Reference: transformers.py:171-172."""
    return pd.DataFrame(list(s), index=s.index).fillna(1000)


def motif_encode_left_df(s):
    """Reference: transformers.py:174-175."""
    return _column([motif_encode_left(y) for y in s], s.index)


def motif_encode_right_df(s):
    """Reference: transformers.py:177-178."""
    return _column([motif_encode_right(y) for y in s], s.index)


def allele_encode_df(s):
    """REF and first ALT -> base codes.  Reference: transformers.py:180-181."""
    two = pd.DataFrame(np.array([x[:2] for x in s]), index=s.index)
    return _elementwise(two, allele_encode)


def allele_encode_single(df):
    """Reference: transformers.py:183-184."""
    return _elementwise(df, allele_encode)


def gt_encode_df(s):
    """Reference: transformers.py:186-187."""
    return _column([gt_encode(y) for y in s], s.index)


def ins_del_encode_df(df):
    """Reference: transformers.py:189-190."""
    return _column(np.array(df[0].apply(ins_del_encode)), df.index)


def svtype_encode_df(df):
    """Reference: transformers.py:192-193."""
    return _column(np.array(df.apply(svtype_encode)), df.index)


def region_annotation_encode_df(df):
    """Reference: transformers.py:195-198."""
    return _column(np.array(df["region_annotations"].apply(region_annotation_encode)), df.index)


def copy_number_encode_df(df):
    """Reference: transformers.py:200-201."""
    return pd.DataFrame(df.max(axis=1), index=df.index)


def cnv_source_encode_df(df):
    """Reference: transformers.py:203-204."""
    return _column(np.array(df["cnv_source"].apply(cnv_source_encode)), df.index)


def convert_to_numeric(df):
    """Reference: transformers.py:326-327."""
    return pd.DataFrame(pd.to_numeric(pd.Series(df.iloc[:, 0])))


def _fn(func):
    return preprocessing.FunctionTransformer(func)


def _zero_fill():
    return impute.SimpleImputer(strategy="constant", fill_value=0)


# --------------------------------------------------------------------------- feature lists
def modify_features_based_on_vcf_type(  # noqa: C901
    vtype: VcfType = VcfType.SINGLE_SAMPLE, custom_annotations: list | None = None
) -> tuple[list, list, str]:
    """Feature names, the ColumnTransformer entry list and the qual column name
    for a VCF flavour.  Reference: transformers.py:144-347 (entry names, order
    and column specs are the contract: they fix the feature-matrix layout)."""
    tuple_first = _fn(tuple_encode_df)
    doublet = _fn(tuple_encode_doublet_df)
    entries = [
        ("ad", doublet, "ad"),
        ("gt", _fn(gt_encode_df), "gt"),
        ("gq", _zero_fill(), ["gq"]),
        ("pl", _fn(tuple_uniform_encode), "pl"),
        ("af", tuple_first, "af"),
        ("sor", _zero_fill(), ["sor"]),
        ("dp", _zero_fill(), ["dp"]),
        ("alleles", _fn(allele_encode_df), "alleles"),
        ("x_hin", make_pipeline(_fn(tuple_encode_df), _fn(allele_encode_single)), "x_hin"),
        ("x_hil", make_pipeline(_fn(tuple_encode_df), _zero_fill()), "x_hil"),
        ("x_il", make_pipeline(_fn(tuple_encode_df), _zero_fill()), "x_il"),
        ("indel", "passthrough", ["indel"]),
        ("x_ic", make_pipeline(_fn(tuple_encode_df), _fn(ins_del_encode_df)), "x_ic"),
        ("x_lm", _fn(motif_encode_left_df), "x_lm"),
        ("x_rm", _fn(motif_encode_right_df), "x_rm"),
        ("x_css", make_pipeline(_fn(tuple_encode_df), preprocessing.OrdinalEncoder()), "x_css"),
        ("x_gcc", _zero_fill(), ["x_gcc"]),
    ]
    if vtype == VcfType.DEEP_VARIANT:
        entries.append(("vaf", _fn(tuple_encode_df), "vaf"))
        features = [e[0] for e in entries]
    elif vtype == VcfType.DEEP_VARIANT_WITH_SOFTCLIP_COUNTS:
        for name in (
            "mq0_ref", "mq0_alt", "ls_ref", "ls_alt", "rs_ref", "rs_alt",
            "mean_nm_ref", "median_nm_ref", "mean_nm_alt", "median_nm_alt",
            "mean_mis_ref", "median_mis_ref", "mean_mis_alt", "median_mis_alt",
        ):
            entries.append((name, _zero_fill(), [name]))
        entries.append(("qual", "passthrough", ["qual"]))
        features = [e[0] for e in entries]
    elif vtype == VcfType.SINGLE_SAMPLE:
        entries.append(("qual", "passthrough", ["qual"]))
        for name in ("fs", "qd", "mq", "an", "baseqranksum", "excesshet", "mqranksum", "readposranksum"):
            entries.append((name, _zero_fill(), [name]))
        entries.append(("ac", _fn(tuple_encode_df), "ac"))
        entries.append(("xc", _zero_fill(), ["xc"]))
        for name in ("mleac", "mleaf", "hapcomp"):
            entries.append((name, _fn(tuple_encode_df), name))
        for name in ("mq0c", "scl", "scr"):
            entries.append((name, _fn(tuple_encode_doublet_df), name))
        features = [e[0] for e in entries]
    elif vtype == VcfType.JOINT:
        features = [e[0] for e in entries]
    elif vtype == VcfType.CNV:
        entries = [("svtype", _fn(svtype_encode_df), "svtype")]
        for name in ("pytorq0", "pytorp2", "pytorrd", "pytorp1", "pytorp3"):
            entries.append((name, _zero_fill(), [name]))
        for name in (
            "gap_percentage", "cnv_dup_reads", "cnv_del_reads", "cnv_dup_frac", "cnv_del_frac",
            "jalign_dup_support", "jalign_del_support", "jalign_dup_support_strong", "jalign_del_support_strong",
        ):
            entries.append((name, "passthrough", [name]))
        entries.append(("svlen", _fn(tuple_encode_df), "svlen"))
        entries.append(("copynumber", _fn(copy_number_encode_df), ["cn", "copynumber"]))
        entries.append(("cnv_source", _fn(cnv_source_encode_df), ["cnv_source"]))
        features = []
        for e in entries:
            features.extend([e[2]] if isinstance(e[2], str) else e[2])
    else:
        raise ValueError("Unrecognized VCF type")

    # custom annotations: TRUE / absent by default, a few with registered transforms
    # (reference transformers.py:316-344)
    for an in custom_annotations or []:
        if an == "long_hmer":
            trans = make_pipeline(
                impute.SimpleImputer(strategy="constant", fill_value="0", missing_values=None),
                _fn(convert_to_numeric),
            )
        elif an == "region_annotations":
            trans = make_pipeline(_fn(region_annotation_encode_df))
        else:
            trans = make_pipeline(
                impute.SimpleImputer(strategy="constant", missing_values=None, fill_value="FALSE"),
                preprocessing.OrdinalEncoder(),
            )
        features.append(an)
        entries.append((an, trans, [an]))
    return features, entries, "qual__qual"


def get_needed_features(vtype: VcfType = VcfType.SINGLE_SAMPLE, custom_annotations: list | None = None) -> list:
    """Reference: transformers.py:126-141."""
    return modify_features_based_on_vcf_type(vtype, custom_annotations)[0]


def get_transformer(vtype: VcfType, annots: list | None = None) -> compose.ColumnTransformer:
    """ColumnTransformer (pandas output) for a VCF flavour.  Reference:
    transformers.py:350-369."""
    _, entries, _ = modify_features_based_on_vcf_type(vtype, annots)
    transformer = compose.ColumnTransformer(entries)
    transformer.set_output(transform="pandas")
    return transformer


# --------------------------------------------------------------------------- labels
def encode_label(label: tuple[int, ...]) -> int:
    """Reference: transformers.py:388-397."""
    label = tuple(sorted(label))
    if label == (0, 0):
        return 0
    if label == (0, 1):
        return 1
    if label == (1, 1):
        return 2
    raise ValueError(f"Encoding of gt={label} not supported")


def encode_labels(ll: Iterable[tuple[int, int]]) -> list[int]:
    """Reference: transformers.py:372-385."""
    return [encode_label(x) for x in ll]


def decode_label(label: int) -> tuple[int, int]:
    """Reference: transformers.py:400-403."""
    return {0: (0, 1), 1: (1, 1), 2: (0, 0)}[label]


label_encode = preprocessing.FunctionTransformer(encode_labels)
