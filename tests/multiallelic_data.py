"""Synthetic call set for the ``--treat_multiallelics`` branch: records that are consistent with a
(synthetic) reference FASTA, with multi-allelic sites, deletions and the ``*`` rows they span.

Only shapes the reference itself can process are generated (it raises on a hom-ref genotype at a
multi-allelic site, on a spanned row whose only ALT is ``*``, and near contig starts its window
slice goes negative -- see DESIGN.md section 4): every multi-allelic genotype carries an ALT
allele and all positions are > 40.
"""
from __future__ import annotations

import numpy as np

from variantcalling_b200 import synth

BASES = "ACGT"
CONTIGS = {"chrM1": 9000, "chrM2": 5000}


def make_reference(seed: int) -> dict:
    """Random sequence with planted homopolymer runs (so hmer indels exist) and a few N / lower case."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, length in CONTIGS.items():
        seq = [BASES[i] for i in rng.integers(0, 4, size=length)]
        p = 0
        while p < length - 12:  # explicit runs
            if rng.random() < 0.12:
                ln = int(rng.integers(2, 10))
                seq[p:p + ln] = [BASES[rng.integers(0, 4)]] * ln
                p += ln
            p += 1
        seq = seq[:length]
        for _ in range(6):
            p = int(rng.integers(100, length - 100))
            seq[p] = "N" if rng.random() < 0.5 else seq[p].lower()
        out[name] = "".join(seq)
    return out


def fasta_text(ref: dict, width: int = 60) -> str:
    parts = []
    for name, seq in ref.items():
        parts.append(">" + name)
        parts.extend(seq[i:i + width] for i in range(0, len(seq), width))
    return "\n".join(parts) + "\n"


def _other_base(rng, b: str) -> str:
    b = b.upper() if b.upper() in BASES else "A"
    return BASES[(BASES.index(b) + int(rng.integers(1, 4))) % 4]


def _pl_index(a: int, b: int) -> int:
    lo, hi = min(a, b), max(a, b)
    return hi * (hi + 1) // 2 + lo


def _record(rng, contig, pos, alleles, gt, customs, fs_bias=0.0) -> str:
    n_all = len(alleles)
    n_alt = n_all - 1
    n_pl = n_all * (n_all + 1) // 2
    # genotypes made of called alleles are likelier than the rest (as a caller's PLs are); the
    # reference's allele ordering relies on it (it asserts that the strongest allele is in GT)
    pl = np.zeros(n_pl, dtype=np.int64)
    for b in range(n_all):
        for a in range(b + 1):
            pl[_pl_index(a, b)] = int(rng.integers(15, 900)) + 1000 * sum(1 for g in (a, b) if g not in gt)
    if rng.random() < 0.15:  # ties in the PL vector exercise the ordering rules
        pl[int(rng.integers(0, n_pl))] = pl[int(rng.integers(0, n_pl))]
    pl[_pl_index(*gt)] = 0
    dp = int(rng.poisson(34)) + 1
    ad = rng.multinomial(dp, np.ones(n_all) / n_all)
    rep = lambda f: ",".join(f() for _ in range(n_alt))  # noqa: E731
    ref, alts = alleles[0], alleles[1:]
    ics, ils = [], []
    for a in alts:
        if a == "*" or len(a) == len(ref):
            ics.append("NA"), ils.append(".")
        elif len(a) > len(ref):
            ics.append("ins"), ils.append(str(len(a) - len(ref)))
        else:
            ics.append("del"), ils.append(str(len(ref) - len(a)))
    hil = [str(int(rng.integers(0, 12))) if ic != "NA" and rng.random() < 0.6 else "." for ic in ics]
    hin = [BASES[rng.integers(0, 4)] if h != "." else "." for h in hil]
    vtype = "snp" if all(ic == "NA" for ic in ics) else ("h-indel" if any(h not in (".", "0") for h in hil) else "non-h-indel")
    ac = [sum(1 for g in gt if g == k) for k in range(1, n_all)]
    info = ["AC=" + ",".join(map(str, ac)), "AF=" + ",".join(f"{c / 2:.3f}" for c in ac), "AN=2"]
    if rng.random() < 0.7:
        info.append(f"BaseQRankSum={rng.normal():.3f}")
    info += [f"DP={dp + int(rng.integers(0, 3))}", "ExcessHet=3.0103", f"FS={rng.exponential(2.0) + fs_bias:.3f}",
             "HAPCOMP=" + rep(lambda: str(int(rng.integers(0, 6)))), "MLEAC=" + ",".join(map(str, ac)),
             "MLEAF=" + ",".join(f"{c / 2:.3f}" for c in ac), f"MQ={60 - rng.exponential(1.5):.2f}",
             "MQ0C=" + ",".join(str(int(rng.integers(0, 4))) for _ in range(n_all))]
    if rng.random() < 0.7:
        info.append(f"MQRankSum={rng.normal():.3f}")
    info.append(f"QD={rng.uniform(1, 35):.2f}")
    if rng.random() < 0.7:
        info.append(f"ReadPosRankSum={rng.normal():.3f}")
    info += ["SCL=" + ",".join(str(int(rng.integers(0, 4))) for _ in range(n_all)),
             "SCR=" + ",".join(str(int(rng.integers(0, 4))) for _ in range(n_all)),
             f"SOR={rng.gamma(2.0, 0.6):.3f}", f"VARIANT_TYPE={vtype}", f"XC={int(rng.integers(0, 12))}",
             "X_CSS=" + rep(lambda: ("non-skip", "possible-cycle-skip", "cycle-skip")[rng.integers(0, 3)]),
             f"X_GCC={rng.uniform():.2f}", "X_HIL=" + ",".join(hil), "X_HIN=" + ",".join(hin),
             "X_IC=" + ",".join(ics), "X_IL=" + ",".join(ils),
             "X_LM=" + rep(lambda: "".join(BASES[i] for i in rng.integers(0, 4, size=5))),
             "X_RM=" + rep(lambda: "".join(BASES[i] for i in rng.integers(0, 4, size=5)))]
    for tag in customs:
        if rng.random() < 0.2:
            info.append(f"{tag}={int(rng.integers(7, 20))}" if tag == "LONG_HMER" else f"{tag}=TRUE")
    filt = "." if rng.random() < 0.85 else ("PASS" if rng.random() < 0.7 else "LowQual")
    qual = f"{np.exp(rng.normal(5.0, 1.2)):.2f}"
    gq = int(rng.integers(0, 100))
    sample = (f"{gt[0]}/{gt[1]}:{','.join(map(str, ad))}:{dp if rng.random() > 0.03 else '.'}:{gq}:"
              f"{','.join(str(int(v)) for v in pl)}")
    return "\t".join([contig, str(pos), ".", ref, ",".join(alts), qual, filt, ";".join(info), "GT:AD:DP:GQ:PL", sample])


def _ins_allele(rng, seq, p):
    """anchor + inserted bases; half of the time the insertion extends the homopolymer after p."""
    anchor = seq[p]
    nxt = seq[p + 1].upper()
    if rng.random() < 0.5 and nxt in BASES:
        return anchor + nxt * int(rng.integers(1, 4))
    return anchor + "".join(BASES[i] for i in rng.integers(0, 4, size=int(rng.integers(1, 5))))


def _gt_with_alt(rng, n_all, forbid=()):
    while True:
        a, b = sorted(int(v) for v in rng.integers(0, n_all, size=2))
        if (a, b) != (0, 0) and (a, b) not in forbid:
            return (a, b)


def generate(seed: int = 11, n_custom: int = 3):
    """-> dict(header, lines, customs, ref (contig -> sequence), labels (0/1/2 genotype class))."""
    rng = np.random.default_rng(seed)
    ref = make_reference(seed)
    customs = synth.custom_annotation_names(n_custom)
    spec = synth.SynthSpec(n_records=0, n_custom=n_custom, contigs=dict(CONTIGS))
    header = synth.header_lines(spec)
    lines = []
    for contig, seq in ref.items():
        p = 45
        while p < len(seq) - 80:
            if "N" in seq[p:p + 12].upper():  # the reference raises on a non-ACGT allele (flow key)
                p += 12
                continue
            u = rng.random()
            s = seq.upper()
            if u < 0.50:  # biallelic SNP
                al = (seq[p], _other_base(rng, seq[p]))
                gt = _gt_with_alt(rng, 2) if rng.random() < 0.9 else (0, 0)
                lines.append(_record(rng, contig, p + 1, al, gt, customs))
            elif u < 0.60:  # biallelic insertion
                al = (seq[p], _ins_allele(rng, s, p))
                lines.append(_record(rng, contig, p + 1, al, _gt_with_alt(rng, 2), customs))
            elif u < 0.80:  # deletion (possibly multi-allelic), maybe with spanned rows
                dl = int(rng.integers(1, 9))
                r = seq[p:p + 1 + dl]
                if rng.random() < 0.25:
                    second = r[:1 + int(rng.integers(0, dl))] if rng.random() < 0.5 else _other_base(rng, r[0]) + r[1:]
                    if second == r or second == r[0]:
                        second = r + "T"
                    al = (r, r[0], second)
                else:
                    al = (r, r[0])
                gt = _gt_with_alt(rng, len(al))
                lines.append(_record(rng, contig, p + 1, al, gt, customs))
                if rng.random() < 0.7:
                    inner = sorted({int(v) for v in rng.integers(p + 1, p + 1 + dl, size=int(rng.integers(1, 3)))})
                    for q in inner:
                        kind = rng.random()
                        if kind < 0.65:  # spanned SNP / insertion row with '*'
                            other = _other_base(rng, seq[q]) if rng.random() < 0.7 else _ins_allele(rng, s, q)
                            al2 = (seq[q], other, "*") if rng.random() < 0.6 else (seq[q], "*", other)
                            star = al2.index("*")
                            gt2 = _gt_with_alt(rng, 3, forbid={(star, star)})
                        elif kind < 0.8:  # four alleles with '*'
                            al2 = (seq[q], _other_base(rng, seq[q]), "*", seq[q] + "GG")
                            gt2 = _gt_with_alt(rng, 4, forbid={(2, 2), (0, 2)})
                            if len(set(al2)) < 4:
                                continue
                        else:  # plain SNP inside the deleted span (no '*': stays a normal row)
                            al2 = (seq[q], _other_base(rng, seq[q]))
                            gt2 = _gt_with_alt(rng, 2)
                        lines.append(_record(rng, contig, q + 1, al2, gt2, customs))
                p += dl
            elif u < 0.97:  # multi-allelic without deletion
                k = rng.random()
                a1 = _other_base(rng, seq[p])
                if k < 0.35:
                    a2 = next(b for b in BASES if b not in (seq[p].upper(), a1))
                elif k < 0.7:
                    a2 = _ins_allele(rng, s, p)
                else:
                    a1, a2 = _ins_allele(rng, s, p), _ins_allele(rng, s, p) + "A"
                al = (seq[p], a1, a2)
                if len(set(al)) < 3:
                    p += 1
                    continue
                gt = _gt_with_alt(rng, 3)
                lines.append(_record(rng, contig, p + 1, al, gt, customs))
            else:  # four alleles
                al = (seq[p], _other_base(rng, seq[p]), seq[p] + "T", seq[p] + "TT")
                if len(set(al)) < 4:
                    p += 1
                    continue
                gt = _gt_with_alt(rng, 4)
                lines.append(_record(rng, contig, p + 1, al, gt, customs))
            p += int(rng.integers(3, 30))
    # labels: genotype class from the written GT (0: hom-ref, 1: het, 2: hom-alt)
    labels = []
    for ln in lines:
        g = ln.split("\t")[9].split(":")[0].split("/")
        a, b = int(g[0]), int(g[1])
        labels.append(0 if a == b == 0 else (2 if a == b else 1))
    return dict(header=header, lines=lines, customs=customs, ref=ref, labels=np.array(labels),
                header_text="\n".join(header) + "\n", text=("\n".join(lines) + "\n").encode())
