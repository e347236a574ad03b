#!/usr/bin/env python
"""Generate tests/golden fixtures from the REFERENCE's own code (run in the build container,
where /root/reference exists; the fixtures travel, the reference does not).

  transformer_single_sample.npz
      vcf_text      : a small synthetic VCF (header + records, SURVEY.md 8d schema + edge rows)
      customs       : custom annotation tags
      features_ref  : ugbio_filtering.transformers.get_transformer(SINGLE_SAMPLE, annots)
                      .fit_transform(df) of the REFERENCE module (harness shims: pandas-3
                      applymap alias, object string columns, int->float columns for sklearn>=1.6)
      categories    : fitted OrdinalEncoder categories of the reference transformer (x_css + customs)
  kats.json
      known answers computed by calling the reference's scalar encoders / phred directly.
"""
import json
import os
import sys
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

from ugbio_core import math_utils as ref_math  # noqa: E402
from ugbio_filtering import transformers as ref_t  # noqa: E402
from ugbio_filtering.tprep_constants import VcfType as RefVcfType  # noqa: E402

from oracle import ref_pipeline as R  # noqa: E402
from oracle.vcf_reader import OracleVariantFile  # noqa: E402
from variantcalling_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def edge_rows() -> list[str]:
    """Hand-written records exercising the typed-decode rules (VCF 4.2 + htslib float32)."""
    fmt = "GT:AD:DP:GQ:PL"
    base_info = ("AC=1;AF=0.500;AN=2;DP=30;ExcessHet=3.0103;FS=0.000;HAPCOMP=1;MLEAC=1;MLEAF=0.500;MQ=60.00;"
                 "MQ0C=0,0;QD=12.5;SCL=0,1;SCR=2,3;SOR=0.693;VARIANT_TYPE=snp;XC=3;X_CSS=non-skip;X_GCC=0.45;"
                 "X_HIL=.;X_HIN=.;X_IC=NA;X_IL=.;X_LM=ACGTA;X_RM=TTGCA")
    rows = [
        # float32 rounding of QUAL / QD / FS; scientific notation; many digits
        f"chrE\t100\t.\tA\tC\t16777217\tPASS\t{base_info}\t{fmt}\t0/1:10,9:19:50:100,0,200",
        f"chrE\t101\trs1\tA\tG\t0.1\t.\t{base_info.replace('QD=12.5', 'QD=1.0000001e1')}\t{fmt}\t1/1:0,22:22:60:500,60,0",
        f"chrE\t102\t.\tG\tT\t33.333333333\tLowQual\t{base_info.replace('FS=0.000', 'FS=1.23456789e-3')}\t{fmt}\t1|1:1,20:21:61:400,50,0",
        # FORMAT DP '.' overrides INFO DP; PL shorter than 3; GQ missing
        f"chrE\t103\t.\tC\tT\t50.00\t.\t{base_info}\t{fmt}\t0/1:5,6:.:.:10,0",
        # missing rank sums already absent; X_HIL/X_IL present; indel with hmer
        f"chrE\t104\t.\tCA\tC\t77.77\t.\t{base_info.replace('X_HIL=.', 'X_HIL=7').replace('X_HIN=.', 'X_HIN=A').replace('X_IC=NA', 'X_IC=del').replace('X_IL=.', 'X_IL=1')}\t{fmt}\t0/1:12,8:20:99:150,0,300",
        # trailing FORMAT sub-fields dropped (PL absent from the sample column)
        f"chrE\t105\t.\tT\tTGGC\t20.5\tPASS\t{base_info.replace('X_IC=NA', 'X_IC=ins')}\t{fmt}\t0/1:7,7:14:40",
        # CG alleles, negative rank sums, integer written for a Float tag
        f"chrE\t106\t.\tGGC\tG\t1e2\t.\t{base_info};BaseQRankSum=-1.5;MQRankSum=2;ReadPosRankSum=-0.001\t{fmt}\t1/1:0,30:30:90:900,90,0",
        # X_LM / X_RM containing N and lower-case (unknown -> 0 digit), long_hmer-like customs are in the synthetic rows
        f"chrE\t107\t.\tA\tT\t5\t.\t{base_info.replace('X_LM=ACGTA', 'X_LM=NNACG').replace('X_RM=TTGCA', 'X_RM=acgtN')}\t{fmt}\t0/0:20,0:20:45:0,45,600",
    ]
    return rows


def main():
    os.makedirs(OUT, exist_ok=True)
    spec = synth.SynthSpec(n_records=600, n_custom=5, seed=7, p_format_dp_missing=0.05)
    header, lines, _ = synth.generate(spec)
    header = [h for h in header if not h.startswith("#CHROM")] + ["##contig=<ID=chrE,length=1000000>",
                                                                   "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE1"]
    lines = lines + edge_rows()
    customs = synth.custom_annotation_names(5)
    text = synth.vcf_text(header, lines)
    vf = OracleVariantFile(text)
    df = R.harness_float_columns(R.get_vcf_df(vf, None, customs))
    annots = [c.lower() for c in customs]
    rtr = ref_t.get_transformer(RefVcfType.SINGLE_SAMPLE, annots)
    with pd.option_context("future.infer_string", False):
        x_ref = rtr.fit_transform(df)
    feats = x_ref.to_numpy(dtype=np.float64)
    cats = {}
    for name, trans, _cols in rtr.transformers_:
        last = trans.steps[-1][1] if hasattr(trans, "steps") else trans
        if hasattr(last, "categories_"):
            cats[name] = [str(c) for c in last.categories_[0]]
    np.savez_compressed(os.path.join(OUT, "transformer_single_sample.npz"),
                        vcf_text=np.frombuffer(text, dtype=np.uint8), customs=np.array(customs),
                        features_ref=feats, categories=json.dumps(cats))
    print("golden features", feats.shape, "categories", cats)

    # ---- CNV flavour: the reference's own CNV transformer on a synthetic CNV frame
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_cnv import CUSTOM, make_cnv_vcf  # noqa: PLC0415  (fixture generator shared with the GPU test)

    chdr, clines = make_cnv_vcf()
    ctext = ("\n".join(chdr) + "\n" + "\n".join(clines) + "\n").encode()
    cdf = R.harness_float_columns(R.get_vcf_df(OracleVariantFile(ctext), None, CUSTOM))
    ctr = ref_t.get_transformer(RefVcfType.CNV, ["region_annotations"])
    with pd.option_context("future.infer_string", False):
        cx = ctr.fit_transform(cdf).to_numpy(dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "transformer_cnv.npz"), vcf_text=np.frombuffer(ctext, dtype=np.uint8),
                        customs=np.array(CUSTOM), features_ref=cx)
    print("golden CNV features", cx.shape)

    kats = {
        "tuple_break": [[[1, 2, 3], ref_t.tuple_break((1, 2, 3))], [None, ref_t.tuple_break(None)]],
        "motif_encode_left": {m: ref_t.motif_encode_left(m) for m in ["ATGC", "ACGTA", "NNACG", "acgtN", ""]},
        "motif_encode_right": {m: ref_t.motif_encode_right(m) for m in ["ATGC", "ACGTA", "NNACG", "acgtN", ""]},
        "motif_encode_left_tuple": ref_t.motif_encode_left(("ACGTA",)),
        "motif_encode_left_tuple_single": ref_t.motif_encode_left(("A", "T")),
        "allele_encode": {a: ref_t.allele_encode(a) for a in ["A", "T", "G", "C", "N", "AT", "*", ""]},
        "gt_encode": [[list(g), ref_t.gt_encode(g)] for g in [(1, 1), (1, 0), (0, 1), (0, 0), (None, None), (1,), (1, 2)]],
        "ins_del_encode": {k: ref_t.ins_del_encode(k) for k in ["ins", "del", "NA"]},
        "encode_labels": ref_t.encode_labels([(0, 1), (0, 0), (1, 0), (1, 1)]),
        "region_annotation_encode": {",".join(k): ref_t.region_annotation_encode(k) for k in
                                     [(), ("Clusters",), ("Telomere_Centromere", "Clusters"),
                                      ("Clusters", "Coverage-Mappability", "Telomere_Centromere")]},
        "phred": [[p, float(ref_math.phred([p])[0])] for p in [0.1, 0.01, 0.001, 0.5, 1e-10, 0.999]],
    }
    with open(os.path.join(OUT, "kats.json"), "w") as fh:
        json.dump(kats, fh, indent=1, sort_keys=True)
    print("kats written")


if __name__ == "__main__":
    main()
