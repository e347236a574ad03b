"""Synthetic single-sample VCF generator (host side, NumPy).

The reference's bundled test VCFs and models are git-LFS pointer stubs
(SURVEY.md section 1 fact 2), so every configuration of BASELINE.json runs on
synthetic records in the same schema: GATK ``single_sample``
INFO/FORMAT tags, UG flow annotations ``X_*`` declared as in the reference's
real header (``ugbio_utils/src/core/tests/resources/header.txt:3372-3398``) and
optional ``Number=1,Type=String`` custom annotations.

This module makes small/medium files for the CPU tests, the CLI plumbing config
and the golden fixtures.  The 50 M-record bench input is produced directly in
HBM by the CUDA twin of this generator (``csrc/synth_kernel.cu``), which follows
the same schema but not the same random stream.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

DEFAULT_SEED = 20260922

# GRCh38 primary contig lengths
CONTIG_LENGTHS = {
    "chr1": 248956422, "chr2": 242193529, "chr3": 198295559, "chr4": 190214555, "chr5": 181538259,
    "chr6": 170805979, "chr7": 159345973, "chr8": 145138636, "chr9": 138394717, "chr10": 133797422,
    "chr11": 135086622, "chr12": 133275309, "chr13": 114364328, "chr14": 107043718, "chr15": 101991189,
    "chr16": 90338345, "chr17": 83257441, "chr18": 80373285, "chr19": 58617616, "chr20": 64444167,
    "chr21": 46709983, "chr22": 50818468, "chrX": 156040895, "chrY": 57227415,
}

BASE_CUSTOM = ["LCR", "MAP_UNIQUE", "LONG_HMER", "UG_HCR", "EXOME"]


def custom_annotation_names(n: int) -> list[str]:
    """The first ``n`` custom annotation tag names (cfg 3 uses 40)."""
    names = list(BASE_CUSTOM[:n])
    i = 0
    while len(names) < n:
        names.append(f"ANN{i:02d}")
        i += 1
    return names


@dataclass
class SynthSpec:
    n_records: int = 10000
    seed: int = DEFAULT_SEED
    contigs: dict = field(default_factory=lambda: dict(CONTIG_LENGTHS))
    n_custom: int = 0                 # number of custom String annotations (cfg 3: 40)
    custom_present_p: float = 0.15
    sample_name: str = "SAMPLE1"
    p_missing_ranksum: float = 0.3
    p_multiallelic: float = 0.0       # only for --treat_multiallelics tests
    p_cg: float = 0.015
    p_format_dp_missing: float = 0.01  # exercises FORMAT-overrides-INFO with a "." value
    region: tuple | None = None       # (contig, start, end) -> all records in one window (cfg 1)


def header_lines(spec: SynthSpec) -> list[str]:
    h = ["##fileformat=VCFv4.2",
         '##FILTER=<ID=LowQual,Description="Low quality">']
    info = [
        ("AC", "A", "Integer"), ("AF", "A", "Float"), ("AN", "1", "Integer"), ("BaseQRankSum", "1", "Float"),
        ("DP", "1", "Integer"), ("ExcessHet", "1", "Float"), ("FS", "1", "Float"), ("HAPCOMP", "A", "Integer"),
        ("MLEAC", "A", "Integer"), ("MLEAF", "A", "Float"), ("MQ", "1", "Float"), ("MQ0C", "R", "Integer"),
        ("MQRankSum", "1", "Float"), ("QD", "1", "Float"), ("ReadPosRankSum", "1", "Float"),
        ("SCL", "R", "Integer"), ("SCR", "R", "Integer"), ("SOR", "1", "Float"),
        ("VARIANT_TYPE", "1", "String"), ("XC", "1", "Integer"),
        ("X_CSS", "A", "String"), ("X_GCC", "1", "Float"), ("X_HIL", "A", "Integer"), ("X_HIN", "A", "String"),
        ("X_IC", "A", "String"), ("X_IL", "A", "Integer"), ("X_LM", "A", "String"), ("X_RM", "A", "String"),
    ]
    for tag, number, vtype in info:
        h.append(f'##INFO=<ID={tag},Number={number},Type={vtype},Description="synthetic {tag}">')
    for tag in custom_annotation_names(spec.n_custom):
        h.append(f'##INFO=<ID={tag},Number=1,Type=String,Description="synthetic annotation, {tag}">')
    h += [
        '##FORMAT=<ID=AD,Number=R,Type=Integer,Description="Allelic depths">',
        '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Read depth">',
        '##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Genotype quality">',
        '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
        '##FORMAT=<ID=PL,Number=G,Type=Integer,Description="Phred-scaled likelihoods">',
    ]
    for name, length in spec.contigs.items():
        h.append(f"##contig=<ID={name},length={length}>")
    h.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + spec.sample_name)
    return h


def _contig_counts(n: int, contigs: dict) -> list[tuple[str, int]]:
    total = float(sum(contigs.values()))
    names = list(contigs)
    counts = [int(n * contigs[c] / total) for c in names]
    counts[0] += n - sum(counts)
    return list(zip(names, counts))


_BASES = np.array(list("ACGT"))


def _rand_seq(rng, length: int) -> str:
    return "".join(_BASES[rng.integers(0, 4, size=length)])


def generate(spec: SynthSpec) -> tuple[list[str], list[str], np.ndarray]:
    """Returns (header lines, record lines, truth labels 0/1 per record).

    Labels are a noisy function of the annotations so that a trained model has
    real structure (used to fit the synthetic models and for cfg 5 truth).
    """
    rng = np.random.default_rng(spec.seed)
    n = spec.n_records
    customs = custom_annotation_names(spec.n_custom)
    lines: list[str] = []
    labels = np.zeros(n, dtype=np.int64)
    if spec.region is not None:
        plan = [(spec.region[0], n)]
    else:
        plan = _contig_counts(n, spec.contigs)
    k = 0
    for contig, cnt in plan:
        if cnt == 0:
            continue
        if spec.region is not None:
            lo, hi = spec.region[1], spec.region[2]
        else:
            lo, hi = 1, spec.contigs[contig]
        gaps = rng.exponential(1.0, size=cnt)
        pos = lo + np.floor(np.cumsum(gaps) / (gaps.sum() + 1.0) * (hi - lo)).astype(np.int64)
        kind_u = rng.random(cnt)
        qual = np.exp(rng.normal(5.0, 1.2, size=cnt))
        dp = rng.poisson(35, size=cnt)
        fs = rng.exponential(2.0, size=cnt)
        mq = np.clip(60.0 - rng.exponential(1.5, size=cnt), 20, 60)
        qd = rng.uniform(1.0, 35.0, size=cnt)
        sor = rng.gamma(2.0, 0.6, size=cnt)
        ranks = rng.normal(0, 1, size=(cnt, 3))
        rank_missing = rng.random((cnt, 3)) < spec.p_missing_ranksum
        gcc = rng.uniform(0, 1, size=cnt)
        gt_u = rng.random(cnt)
        gq = rng.integers(0, 100, size=cnt)
        id_u = rng.random(cnt)
        filt_u = rng.random(cnt)
        xc = rng.integers(0, 12, size=cnt)
        hapcomp = rng.integers(0, 7, size=cnt)
        dbl = rng.integers(0, 4, size=(cnt, 6))
        hil = rng.integers(0, 21, size=cnt)
        css = rng.integers(0, 3, size=cnt)
        cust_u = rng.random((cnt, max(1, len(customs))))
        noise = rng.normal(0, 1.0, size=cnt)
        for i in range(cnt):
            u = kind_u[i]
            multi = spec.p_multiallelic > 0 and u > 1.0 - spec.p_multiallelic
            ref = _BASES[rng.integers(0, 4)]
            x_ic, x_il, x_hil, x_hin = "NA", ".", ".", "."
            vtype = "snp"
            if u < 0.80 or multi:
                alt = _BASES[(np.where(_BASES == ref)[0][0] + rng.integers(1, 4)) % 4]
                if multi:
                    alt2 = ref + _rand_seq(rng, int(rng.integers(1, 4)))
                    alt = f"{alt},{alt2}"
            elif u < 0.90:  # 1-bp indel
                extra = _rand_seq(rng, 1)
                if rng.random() < 0.5:
                    alt, x_ic = ref + extra, "ins"
                else:
                    ref, alt, x_ic = ref + extra, ref, "del"
                x_il, vtype = "1", "h-indel"
                x_hil, x_hin = str(hil[i]), extra
            elif u < 0.98:  # 2-10 bp indel
                ln = int(rng.integers(2, 11))
                extra = _rand_seq(rng, ln)
                if rng.random() < 0.5:
                    alt, x_ic = ref + extra, "ins"
                else:
                    ref, alt, x_ic = ref + extra, ref, "del"
                x_il, vtype = str(ln), "non-h-indel"
                if rng.random() < 0.3:
                    x_hil, x_hin = str(hil[i]), extra[0]
            else:  # CG-type insertion/deletion alleles (blacklist_cg_insertions target)
                trip = "GGC" if rng.random() < 0.5 else "CCG"
                if rng.random() < 0.5:
                    ref, alt, x_ic = trip[0], trip, "ins"
                else:
                    ref, alt, x_ic = trip, trip[0], "del"
                x_il, vtype = "2", "non-h-indel"
            n_alt = alt.count(",") + 1
            g = gt_u[i]
            if multi:
                gt, ac = "1/2", "1,1"
            elif g < 0.60:
                gt, ac = "0/1", "1"
            elif g < 0.98:
                gt, ac = "1/1", "2"
            else:
                gt, ac = "0/0", "0"
            af = ",".join(["0.500"] * n_alt) if gt != "1/1" else "1.00"
            d = int(dp[i])
            alt_reads = rng.binomial(d, 0.5 if gt != "1/1" else 0.97) if d > 0 else 0
            ad = [d - alt_reads, alt_reads] + [int(rng.integers(0, 5)) for _ in range(n_alt - 1)]
            n_pl = (n_alt + 1) * (n_alt + 2) // 2
            pl = rng.integers(20, 2000, size=n_pl)
            pl[rng.integers(0, n_pl)] = 0
            info = [f"AC={ac}", f"AF={af}", "AN=2"]
            if not rank_missing[i, 0]:
                info.append(f"BaseQRankSum={ranks[i, 0]:.3f}")
            info.append(f"DP={d + int(rng.integers(0, 4))}")
            info.append("ExcessHet=3.0103")
            info.append(f"FS={fs[i]:.3f}")
            info.append("HAPCOMP=" + ",".join(str(int(hapcomp[i])) for _ in range(n_alt)))
            info.append(f"MLEAC={ac}")
            info.append(f"MLEAF={af}")
            info.append(f"MQ={mq[i]:.2f}")
            info.append("MQ0C=" + ",".join(str(int(v)) for v in ([dbl[i, 0], dbl[i, 1]] + [0] * (n_alt - 1))))
            if not rank_missing[i, 1]:
                info.append(f"MQRankSum={ranks[i, 1]:.3f}")
            info.append(f"QD={qd[i]:.2f}")
            if not rank_missing[i, 2]:
                info.append(f"ReadPosRankSum={ranks[i, 2]:.3f}")
            info.append("SCL=" + ",".join(str(int(v)) for v in ([dbl[i, 2], dbl[i, 3]] + [0] * (n_alt - 1))))
            info.append("SCR=" + ",".join(str(int(v)) for v in ([dbl[i, 4], dbl[i, 5]] + [0] * (n_alt - 1))))
            info.append(f"SOR={sor[i]:.3f}")
            info.append(f"VARIANT_TYPE={vtype}")
            info.append(f"XC={int(xc[i])}")
            rep = lambda v: ",".join([v] * n_alt)  # noqa: E731
            info.append("X_CSS=" + rep(("non-skip", "possible-cycle-skip", "cycle-skip")[css[i]]))
            info.append(f"X_GCC={gcc[i]:.2f}")
            info.append("X_HIL=" + rep(x_hil))
            info.append("X_HIN=" + rep(x_hin))
            info.append("X_IC=" + rep(x_ic))
            info.append("X_IL=" + rep(x_il))
            info.append("X_LM=" + rep(_rand_seq(rng, 5)))
            info.append("X_RM=" + rep(_rand_seq(rng, 5)))
            lcr = False
            for j, tag in enumerate(customs):
                if cust_u[i, j] < spec.custom_present_p:
                    if tag == "LONG_HMER":
                        info.append(f"{tag}={int(rng.integers(7, 20))}")
                    else:
                        info.append(f"{tag}=TRUE")
                    lcr = lcr or tag == "LCR"
            f_u = filt_u[i]
            filt = "." if f_u < 0.9 else ("PASS" if f_u < 0.97 else "LowQual")
            vid = "." if id_u[i] < 0.85 else f"rs{int(rng.integers(1, 10**8))}"
            fmt_dp = "." if rng.random() < spec.p_format_dp_missing else str(d)
            sample = f"{gt}:{','.join(str(v) for v in ad)}:{fmt_dp}:{int(gq[i])}:{','.join(str(int(v)) for v in pl)}"
            lines.append("\t".join([contig, str(int(pos[i])), vid, ref, alt, f"{qual[i]:.2f}", filt,
                                    ";".join(info), "GT:AD:DP:GQ:PL", sample]))
            # truth: a noisy score over the annotations (higher = more likely a true variant)
            z = (0.08 * (qd[i] - 15) - 0.25 * (fs[i] - 2) - 0.9 * (sor[i] - 1.2) + 0.04 * (d - 30)
                 + 0.15 * (mq[i] - 58) + (0.0 if vtype == "snp" else -0.8) - (0.7 if lcr else 0.0)
                 + 0.004 * (min(qual[i], 600) - 150) + 0.8 * noise[i])
            labels[k] = 1 if z > 0 else 0
            k += 1
    return header_lines(spec), lines, labels


def vcf_text(header: list[str], lines: list[str]) -> bytes:
    return ("\n".join(header) + "\n" + "\n".join(lines) + ("\n" if lines else "")).encode()
