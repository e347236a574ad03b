"""DeepVariant-style call set (the flavour of the reference's real header fixture,
``ugbio_utils/src/core/tests/resources/header.txt``: INFO AF / SOR / VARIANT_TYPE / X_* / region
annotations, FORMAT GT:GQ:DP:AD:VAF:PL with MED_DP / MIN_DP declared, FILTER RefCall / NoCall / LowQual),
derived from the GATK-style synthetic records by dropping the GATK-only keys."""
from __future__ import annotations

import numpy as np

from variantcalling_b200 import synth

CUSTOMS = ["LCR", "MAP_UNIQUE", "LONG_HMER", "UG_HCR", "EXOME"]
KEEP = {"AF", "SOR", "VARIANT_TYPE", "X_CSS", "X_GCC", "X_HIL", "X_HIN", "X_IC", "X_IL", "X_LM", "X_RM", *CUSTOMS}


def header_lines(contigs: dict) -> list[str]:
    h = ["##fileformat=VCFv4.2",
         '##FILTER=<ID=PASS,Description="All filters passed">',
         '##FILTER=<ID=LowQual,Description="Confidence in this variant being real is below calling threshold.">',
         '##FILTER=<ID=NoCall,Description="Site has depth=0 resulting in no call.">',
         '##FILTER=<ID=RefCall,Description="Genotyping model thinks this site is reference.">',
         '##INFO=<ID=DB,Number=0,Type=Flag,Description="dbSNP Membership">',
         '##INFO=<ID=END,Number=1,Type=Integer,Description="End position">']
    for tag in ("EXOME", "LCR", "LONG_HMER", "MAP_UNIQUE", "UG_HCR", "VARIANT_TYPE"):
        h.append(f'##INFO=<ID={tag},Number=1,Type=String,Description="{tag}">')
    h.append('##INFO=<ID=GNOMAD_AF,Number=A,Type=Float,Description="gnomad">')
    h.append('##INFO=<ID=SOR,Number=1,Type=Float,Description="Symmetric Odds Ratio">')
    h.append('##INFO=<ID=X_GCC,Number=1,Type=Float,Description="Flow: GC">')
    for tag, typ in (("X_CSS", "String"), ("X_HIL", "Integer"), ("X_HIN", "String"), ("X_IC", "String"), ("X_IL", "Integer"),
                     ("X_LM", "String"), ("X_RM", "String")):
        h.append(f'##INFO=<ID={tag},Number=A,Type={typ},Description="Flow: {tag}">')
    h.append('##INFO=<ID=AF,Number=A,Type=Float,Description="Allele frequency">')
    h += ['##FORMAT=<ID=AD,Number=R,Type=Integer,Description="Read depth for each allele">',
          '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Read depth">',
          '##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Conditional genotype quality">',
          '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
          '##FORMAT=<ID=MED_DP,Number=1,Type=Integer,Description="Median DP">',
          '##FORMAT=<ID=MIN_DP,Number=1,Type=Integer,Description="Minimum DP">',
          '##FORMAT=<ID=PL,Number=G,Type=Integer,Description="Phred-scaled genotype likelihoods">',
          '##FORMAT=<ID=VAF,Number=A,Type=Float,Description="Variant allele fractions.">']
    h += [f"##contig=<ID={c},length={n}>" for c, n in contigs.items()]
    h.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE1")
    return h


def generate(n_records: int = 3000, seed: int = 3):
    spec = synth.SynthSpec(n_records=n_records, n_custom=len(CUSTOMS), seed=seed)
    _h, lines, labels = synth.generate(spec)
    rng = np.random.default_rng(seed)
    out = []
    for ln in lines:
        c = ln.split("\t")
        info = [kv for kv in c[7].split(";") if kv.split("=", 1)[0] in KEEP]
        if rng.random() < 0.2:
            info.insert(int(rng.integers(0, len(info) + 1)), "DB")
        if rng.random() < 0.1:
            info.append(f"GNOMAD_AF={rng.random():.4g}")
        gt, ad, dp, gq, pl = c[9].split(":")
        ads = [int(v) for v in ad.split(",")]
        tot = max(1, sum(ads))
        vaf = ",".join(f"{a / tot:.6g}" for a in ads[1:])
        c[6] = {".": "PASS", "PASS": "PASS", "LowQual": "RefCall"}[c[6]] if rng.random() < 0.9 else "NoCall"
        c[7] = ";".join(info)
        c[8] = "GT:GQ:DP:AD:VAF:PL"
        c[9] = ":".join([gt, gq, dp, ad, vaf, pl])
        out.append("\t".join(c))
    header = header_lines(dict(spec.contigs))
    return dict(header=header, lines=out, labels=labels, customs=list(CUSTOMS),
                header_text="\n".join(header) + "\n", text=("\n".join(out) + "\n").encode())
