#!/usr/bin/env python
"""Generate tests/golden/multiallelic_*.json from the REFERENCE's own multi-allelic / spanning-deletion
code (run in the build container; the fixtures travel, /root/reference does not).

The reference modules import pysam / pyfaidx / xgboost, which are not installed here.  None of the
functions exercised below needs their engines -- they only read a header's Number fields and slice
a chromosome string -- so the three packages are replaced by minimal stand-ins:
  pysam.VariantFile(path).header.info[tag].number / .formats[tag].number   <- oracle header parser
  pyfaidx.Fasta(path, ...)[contig]                                          <- plain str per contig
Everything else (select_overlapping_variants, split_multiallelic_variants[_with_spandel],
cleanup_multiallelics, process_multiallelic_spandel, combine_multiallelic_spandel,
merge_and_assign_pls, classify_hmer_indel_relative and the flow-key helpers) is the unmodified
reference code, imported from /root/reference.

Outputs (per seed):
  multiallelic_split_<seed>.json.gz  input VCF text + FASTA + the split data frame the reference
                                   builds (every column, python values) + the merged ml_lik the
                                   reference assigns for a fixed random score matrix
"""
import gzip
import json
import os
import sys
import types
import warnings

import numpy as np
import pandas as pd

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/ugbio_utils/src"
sys.path.insert(0, os.path.join(REF, "filtering"))
sys.path.insert(0, os.path.join(REF, "core"))
pd.DataFrame.applymap = pd.DataFrame.map  # pandas-3 harness shim

# pandas-3 harness shim no. 2: the reference (pinned to pandas < 3) indexes label-indexed Series with
# integers (``del_length[i]``, multiallelics.py:40,59) and relies on the positional fall-back pandas
# had until 3.0.  Restore exactly that fall-back, for integer keys on non-integer indexes only.
_series_getitem = pd.Series.__getitem__


def _getitem_with_positional_fallback(self, key):
    try:
        return _series_getitem(self, key)
    except KeyError:
        if isinstance(key, (int, np.integer)) and not pd.api.types.is_integer_dtype(self.index.dtype):
            return self.iloc[key]
        raise


pd.Series.__getitem__ = _getitem_with_positional_fallback

from oracle import ref_pipeline as R  # noqa: E402
from oracle.vcf_reader import OracleVariantFile  # noqa: E402
from tests import multiallelic_data as M  # noqa: E402


class _Meta:
    def __init__(self, number):
        self.number = int(number) if number.isdigit() else number


class _Header:
    def __init__(self, oh):
        self.info = {k: _Meta(v[0]) for k, v in oh.info.items()}
        self.formats = {k: _Meta(v[0]) for k, v in oh.formats.items()}


class _VariantFile:
    def __init__(self, path, *a, **k):
        self.header = _Header(OracleVariantFile(path).header)


class _VariantHeader:
    pass


class _Fasta(dict):
    def __init__(self, path, **k):
        super().__init__()
        name = None
        for ln in open(path):
            ln = ln.strip()
            if ln.startswith(">"):
                name = ln[1:].split()[0]
                self[name] = []
            elif name:
                self[name].append(ln)
        for k2 in list(self):
            self[k2] = "".join(self[k2])


class _Anything(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return type(k, (), {})


class _Pysam(_Anything):
    VariantFile = _VariantFile
    VariantHeader = _VariantHeader


class _Pyfaidx(_Anything):
    Fasta = _Fasta
    FastaRecord = str


sys.modules["pysam"] = _Pysam("pysam")
sys.modules["pyfaidx"] = _Pyfaidx("pyfaidx")
for name in ("xgboost", "ugbio_comparison", "ugbio_comparison.sv_comparison_pipeline", "ugbio_comparison.vcf_comparison_utils"):
    sys.modules[name] = _Anything(name)

from ugbio_filtering import training_prep as ref_tp  # noqa: E402
from ugbio_filtering import variant_filtering_utils as ref_vfu  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def py(v):
    """DataFrame cell -> JSON value (tuples -> lists, NaN -> "NaN" marker, numpy scalars -> python)."""
    if isinstance(v, (tuple, list, np.ndarray)):
        return [py(x) for x in v]
    if v is None:
        return None
    if isinstance(v, (bool, np.bool_)):
        return bool(v)
    if isinstance(v, (int, np.integer)):
        return int(v)
    if isinstance(v, (float, np.floating)):
        return "NaN" if np.isnan(v) else float(v)
    return str(v)


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    ds = M.generate(seed)
    tmp = "/tmp/mg"
    os.makedirs(tmp, exist_ok=True)
    vcf_path, fa_path = os.path.join(tmp, "in.vcf"), os.path.join(tmp, "ref.fa")
    open(vcf_path, "w").write(ds["header_text"] + ds["text"].decode())
    open(fa_path, "w").write(M.fasta_text(ds["ref"]))
    vf = OracleVariantFile(vcf_path)
    out = {"seed": seed, "contigs": {}, "vcf_text": ds["header_text"] + ds["text"].decode(), "customs": ds["customs"],
           "fasta": M.fasta_text(ds["ref"])}
    rng = np.random.default_rng(seed)
    for contig in M.CONTIGS:
        df = R.get_vcf_df(vf, contig, ds["customs"])
        split = ref_tp.process_multiallelic_spandel(df, fa_path, contig, vcf_path)
        scores = rng.dirichlet(np.ones(3), size=split.shape[0])
        df_original = df.copy()
        set_source = [x in df_original.index for x in split.index]
        set_dest = [x in split.index for x in df_original.index]
        df_original["ml_lik"] = pd.Series([list(x) for x in scores[set_source, :]], index=df_original.loc[set_dest].index)
        merged = ref_vfu.combine_multiallelic_spandel(split, df_original, scores)
        out["contigs"][contig] = {
            "columns": list(split.columns),
            "index": [py(i) for i in split.index],
            "rows": [[py(v) for v in row] for row in split.itertuples(index=False, name=None)],
            "dtypes": [str(t) for t in split.dtypes],
            "scores": scores.tolist(),
            "ml_lik": [py(v) for v in merged["ml_lik"]],
        }
        print(contig, df.shape, "->", split.shape, "multiallelic groups", split["multiallelic_group"].notna().sum(),
              "spandel rows", split["spanning_deletion"].notna().sum() if "spanning_deletion" in split else 0)
    path = os.path.join(OUT, f"multiallelic_split_{seed}.json.gz")
    with gzip.open(path, "wt", compresslevel=9) as fh:
        json.dump(out, fh)
    print("wrote", path, os.path.getsize(path) >> 10, "KiB")


if __name__ == "__main__":
    main()
