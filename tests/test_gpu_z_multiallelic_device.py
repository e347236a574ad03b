"""GPU: the --treat_multiallelics kernels (csrc/multiallelic.cu, C ABI ugvc_ma_*) against the oracle and against the
Python model they were written from (variantcalling_b200.multiallelics.SplitPlan).

 * split rows read back through the oracle loader + the fitted transformer give the features of the oracle's split
   frame (which is pinned cell by cell to frames produced by the reference's own functions,
   tests/test_multiallelics_cpu.py), the overlap sets equal the reference row loop's, the merged likelihoods equal
   combine_multiallelic_spandel's;
 * row for row the device text equals the model's (QD is spelled differently: 19 exact digits of the same double);
 * mutated records (missing / malformed PL, GT, DP, X_IL, VARIANT_TYPE, symbolic and '*' alleles, records next to the
   contig start, INFO keys that are FORMAT keys too, extra samples ...) end the same way on both sides: same text,
   or the same exception type -- the reference's error contract.
The file also runs on the host emulation of the kernel sources (tests/test_host_emu_cpu.py)."""
import random
import re

import numpy as np
import pandas as pd
import pytest

from oracle import multiallelic_ref as MR
from oracle import ref_pipeline as R
from oracle.vcf_reader import OracleVariantFile
from tests import util
from tests.test_multiallelics_cpu import contig_text, cpu_index
from variantcalling_b200 import multiallelics as PM
from variantcalling_b200.vcf_header import VcfHeader

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case():
    ds, tr, _model, _ = util.make_multiallelic_case(12, "lr")
    hdr = VcfHeader(ds["header_text"])
    return ds, tr, hdr, hdr.loader_columns(ds["customs"])


def norm_qd(line: bytes) -> bytes:
    def f(m):
        v = m.group(1)
        try:
            return b"QD=" + (v if v in (b".", b"") else repr(float(v)).encode())
        except ValueError:
            return b"QD=" + v
    return re.sub(rb"QD=([^;\t]*)", f, line)


def test_split_rows_and_merge_match_the_oracle(case, gpu_ctx):
    ds, tr, hdr, cols = case
    rng = np.random.default_rng(12)
    plan = None
    for contig in ("chrM1", "chrM2"):
        df = R.get_vcf_df(ds["vf"], contig, ds["customs"])
        want_split = MR.process_multiallelic_spandel(df, ds["ref"][contig], ds["vf"].header)
        text = contig_text(ds, contig)
        ls, ri = cpu_index(text)
        plan = PM.make_split_plan(hdr, cols, ds["ref"][contig], 0, reuse=plan)  # the handle carries over, as in the tool
        assert isinstance(plan, PM.DeviceSplitPlan)
        new_text = plan.build(np.frombuffer(text, dtype=np.uint8), ls, ri).tobytes()
        assert plan.launch_count() > 0
        assert sorted(int(o) for o in plan.origins) == sorted(sum(MR.overlapping_sets(df), []))
        got_df = R.get_vcf_df(OracleVariantFile(ds["header_text"].encode() + new_text), contig, ds["customs"])
        assert got_df.shape[0] == want_split.shape[0]
        with pd.option_context("future.infer_string", False):
            want_x = tr.transform(R.harness_float_columns(want_split)).to_numpy(dtype=np.float64)
            got_x = tr.transform(R.harness_float_columns(got_df)).to_numpy(dtype=np.float64)
        bad = np.argwhere(want_x.astype(np.float32) != got_x.astype(np.float32))
        assert bad.size == 0, (bad[:5], want_x[tuple(bad[0])], got_x[tuple(bad[0])], new_text.split(b"\n")[bad[0][0]])
        # QD as the reference's frame holds it: the text of a split row is the same double, not only the same float32
        n_kept = int(plan.kept.sum())
        want_qd = want_split["qd"].to_numpy(dtype=np.float64)[n_kept:]
        texts = [re.search(rb"QD=([^;\t]*)", ln).group(1) for ln in new_text.split(b"\n")[n_kept:-1]]
        got_qd = np.array([np.nan if t == b"." else float(t) for t in texts])
        assert len(texts) == int(plan.n_rows.sum()) and np.array_equal(want_qd, got_qd, equal_nan=True)
        scores = rng.dirichlet(np.ones(3), size=want_split.shape[0])
        original = df.copy()
        src = [x in original.index for x in want_split.index]
        dst = [x in want_split.index for x in original.index]
        original["ml_lik"] = pd.Series([list(x) for x in scores[src, :]], index=original.loc[dst].index)
        merged = MR.combine_multiallelic_spandel(want_split, original, scores)
        lik = plan.merge(scores)
        for i, want in enumerate(merged["ml_lik"]):
            want = np.asarray(want, dtype=np.float64)
            assert np.array_equal(lik[i, :want.size], want) and not lik[i, want.size:].any(), (i, lik[i], want)
    plan.close()


def test_rows_equal_the_python_model_line_for_line(case, gpu_ctx):
    ds, _tr, hdr, cols = case
    for contig in ("chrM1", "chrM2"):
        text = np.frombuffer(contig_text(ds, contig), dtype=np.uint8)
        ls, ri = cpu_index(text.tobytes())
        host = PM.SplitPlan(hdr, cols, ds["ref"][contig])
        want = host.build(text, ls, ri).tobytes().split(b"\n")
        dev = PM.DeviceSplitPlan(hdr, cols, ds["ref"][contig])
        got = dev.build(text, ls, ri).tobytes().split(b"\n")
        assert [g.origin for g in host.groups] == [int(o) for o in dev.origins]
        assert [len(g.rows) for g in host.groups] == [int(r) for r in dev.n_rows]
        assert len(want) == len(got)
        for i, (a, b) in enumerate(zip(want, got)):
            assert norm_qd(a) == norm_qd(b), (contig, i, a, b)
        lik = np.random.default_rng(3).dirichlet(np.ones(3), size=len(want) - 1)
        assert np.array_equal(host.merge(lik), dev.merge(lik))
        with pytest.raises(IndexError):
            dev.merge(lik[:, :2].copy())  # a 2-class model has no hom-alt likelihood to spread
        with pytest.raises(ValueError, match="unexpected number of rows"):
            dev.merge(lik[:-1])
        dev.close()


def _mutate(rng, line: str) -> str:  # noqa: PLR0912, PLR0915
    c = line.split("\t")
    info, fmt, smp, alts = c[7].split(";"), c[8].split(":"), c[9].split(":"), c[4].split(",")

    def drop(tag):
        return [kv for kv in info if kv.split("=")[0] != tag]

    def swap(tag, choices):
        return [kv if not kv.startswith(tag + "=") else tag + "=" + rng.choice(choices) for kv in info]

    m = rng.randrange(26)
    if m == 0 and "PL" in fmt:
        pl = smp[fmt.index("PL")].split(",")
        pl[rng.randrange(len(pl))] = "."
        smp[fmt.index("PL")] = ",".join(pl)
    elif m == 1 and "PL" in fmt:
        i = fmt.index("PL")
        del fmt[i], smp[i]
    elif m == 2:
        smp[0] = rng.choice(["0|1", "./.", ".", "1", "1|2", "0/1/2", "2/1", "x/1", "1/1", ""])
    elif m == 3:
        c[3] = c[3].lower()
        alts = [a.lower() if rng.random() < 0.5 else a for a in alts]  # noqa: PLR2004
    elif m == 4 and "DP" in fmt:
        smp[fmt.index("DP")] = rng.choice([".", "0", "", "7", "x"])
    elif m == 5:
        info = drop("X_IL")
    elif m == 6:
        info = drop("VARIANT_TYPE") + rng.choice([[], ["VARIANT_TYPE"], ["VARIANT_TYPE=snp"], ["VARIANT_TYPE=h-indel"],
                                                   ["VARIANT_TYPE=non-h-indel"], ["VARIANT_TYPE=other"], ["VARIANT_TYPE="]])
    elif m == 7:
        info = drop("QD")
    elif m == 8 and "GQ" in fmt:
        i = fmt.index("GQ")
        del fmt[i], smp[i]
    elif m == 9:
        c += [":".join(smp), "./."]
    elif m == 10 and "AD" in fmt:
        ad = smp[fmt.index("AD")].split(",")
        smp[fmt.index("AD")] = ",".join(ad[:rng.randrange(1, len(ad) + 1)])
    elif m == 11:
        alts[rng.randrange(len(alts))] = rng.choice(["<DEL>", "<NON_REF>", "*", "N", "ANT", "a*"])
    elif m == 12:
        info = []
    elif m == 13 and len(alts) > 1:
        alts[1] = alts[0]
    elif m == 14:
        info.insert(rng.randrange(len(info) + 1), rng.choice(["AD=1,2,3", "GQ=5", "PL=1,2,3", "gt=0/1", "DP"]))
    elif m == 15:
        smp = smp[:rng.randrange(0, len(smp))] or [smp[0]]
    elif m == 16:
        alts = [rng.choice([".", "*"])]
    elif m == 17:
        extra = ["C", "G", "TT", "GA", "*"]
        rng.shuffle(extra)
        alts = alts + extra[:rng.randrange(1, 4)]
        if "PL" in fmt and rng.random() < 0.8:  # noqa: PLR2004
            n = len(alts) + 1
            smp[fmt.index("PL")] = ",".join(str(rng.randrange(0, 500)) for _ in range(n * (n + 1) // 2))
    elif m == 18:
        info = drop("X_HIL")
    elif m == 19:
        info = swap("X_IL", [".", "", "3", "x", ".,2"])
    elif m == 20:
        info = swap("DP", [".", "0", "-3", "12"])
        if "DP" in fmt and rng.random() < 0.7:  # noqa: PLR2004
            i = fmt.index("DP")
            del fmt[i], smp[i]
    elif m == 21:
        info = swap("AC", ["1"])
    elif m == 22:
        info = swap("X_HIL", [".", "0", "5", ""])
    elif m == 23:
        info = [""] + info + [""]
    elif m == 24 and "PL" in fmt:
        pl = smp[fmt.index("PL")].split(",")
        smp[fmt.index("PL")] = ",".join(pl[:rng.randrange(1, len(pl))])
    elif m == 25:
        c[3] = c[3] + rng.choice(["N", "n", "R"])
    c[4], c[7], c[8], c[9] = ",".join(alts), (";".join(info) if info else "."), ":".join(fmt), ":".join(smp)
    return "\t".join(c)


def _outcome(plan, text, ls, ri, seed):
    try:
        out = plan.build(np.frombuffer(text, dtype=np.uint8), ls, ri).tobytes()
        lik = np.random.default_rng(seed).dirichlet(np.ones(3), size=out.count(b"\n"))
        return "ok", [norm_qd(ln) for ln in out.split(b"\n")], plan.merge(lik)
    except Exception as e:  # noqa: BLE001
        return type(e).__name__, None, None


def test_mutated_records_end_the_same_way_as_in_the_model(case, gpu_ctx):
    ds, _tr, hdr, cols = case
    ends = {}
    dev = None
    for it in range(120):
        rng = random.Random(7000 + it)
        contig = rng.choice(["chrM1", "chrM2"])
        lines = [ln for ln in ds["lines"] if ln.split("\t", 1)[0] == contig]
        a = rng.randrange(0, max(1, len(lines) - 60))
        lines = lines[a:a + rng.randrange(20, 60)]
        if rng.random() < 0.3:  # noqa: PLR2004  (records next to the contig start: the window slice goes negative)
            shift = int(lines[0].split("\t")[1]) - rng.randrange(1, 25)
            lines = ["\t".join([c[0], str(int(c[1]) - shift)] + c[2:]) for c in (ln.split("\t") for ln in lines)]
        multi = [i for i, ln in enumerate(lines) if "," in ln.split("\t")[4] or len(ln.split("\t")[3]) > 1]
        for _ in range(rng.randrange(0, 4)):
            i = rng.choice(multi) if multi and rng.random() < 0.85 else rng.randrange(len(lines))  # noqa: PLR2004
            lines[i] = _mutate(rng, lines[i])
        text = ("\n".join(lines) + "\n").encode()
        try:
            ls, ri = cpu_index(text)
        except ValueError:
            continue
        want = _outcome(PM.SplitPlan(hdr, cols, ds["ref"][contig]), text, ls, ri, it)
        dev = PM.make_split_plan(hdr, cols, ds["ref"][contig], 0, reuse=dev)
        got = _outcome(dev, text, ls, ri, it)
        assert want[0] == got[0], (it, want[0], got[0], text[:300])
        ends[want[0]] = ends.get(want[0], 0) + 1
        if want[0] == "ok":
            assert want[1] == got[1], (it, [(x, y) for x, y in zip(want[1], got[1]) if x != y][:1])
            assert np.array_equal(want[2], got[2]), it
    assert ends.get("ok", 0) > 40 and len(ends) >= 5, ends  # noqa: PLR2004  (the reference's failure types all occur)


def test_lines_longer_than_the_staging_buffer(case, gpu_ctx):
    """The row kernels stage a group's line in 4 KiB of shared memory; a longer line (a long ID, a padded annotation)
    is read where it lies.  Same rows either way."""
    ds, _tr, hdr, cols = case
    contig = "chrM1"
    lines = [ln for ln in ds["lines"] if ln.split("\t", 1)[0] == contig][:400]
    rng = random.Random(5)
    grown = 0
    for i, ln in enumerate(lines):
        c = ln.split("\t")
        if ("," in c[4] or len(c[3]) > 1) and rng.random() < 0.5:  # noqa: PLR2004  (records that end up in groups)
            c[2] = "rs" + "7" * rng.choice([3000, 4090, 4096, 5000, 9000])
            if rng.random() < 0.5:  # noqa: PLR2004
                c[7] += ";PAD=" + "x" * 4200
            lines[i] = "\t".join(c)
            grown += 1
    assert grown > 20  # noqa: PLR2004
    text = np.frombuffer(("\n".join(lines) + "\n").encode(), dtype=np.uint8)
    ls, ri = cpu_index(text.tobytes())
    host = PM.SplitPlan(hdr, cols, ds["ref"][contig])
    want = [norm_qd(x) for x in host.build(text, ls, ri).tobytes().split(b"\n")]
    dev = PM.DeviceSplitPlan(hdr, cols, ds["ref"][contig])
    got = [norm_qd(x) for x in dev.build(text, ls, ri).tobytes().split(b"\n")]
    assert want == got
    assert max(len(x) for x in got) > 8000  # noqa: PLR2004
    dev.close()


def test_overlap_scan_on_random_layouts(case, gpu_ctx):
    """The chunked scans of the overlap detection against the model's find_overlaps (itself checked against the
    reference's row loop in tests/test_multiallelics_cpu.py) on random layouts: deletions spanning deletions, '*' rows
    before any deletion, clusters left open at the end, contigs longer than one scan chunk."""
    ds, _tr, hdr, cols = case
    rng = np.random.default_rng(17)
    dev = None
    seen = set()
    for it in range(160):
        n = int(rng.integers(1, 40)) if it % 8 else int(rng.integers(600, 1500))  # some layouts span several chunks of 256
        pos = np.sort(rng.integers(30, 60 + 3 * n, size=n))
        lines = []
        for i in range(n):
            kind = rng.random()
            ref = "A" * int(rng.integers(1, 6)) if kind < 0.4 else "A"
            alts = ["C"]
            if kind < 0.4 and rng.random() < 0.8:
                alts = ["A" * int(rng.integers(1, len(ref) + 1))]
            if rng.random() < 0.3:
                alts.append("*")
            if rng.random() < 0.3:
                alts.append("AT")
            n_all = len(alts) + 1
            pl = ",".join(str(int(v)) for v in rng.integers(0, 200, size=n_all * (n_all + 1) // 2))
            gt = f"{int(rng.integers(0, n_all))}/{int(rng.integers(1, n_all))}"
            lines.append(f"chrM1\t{pos[i]}\t.\t{ref}\t{','.join(alts)}\t50\t.\tDP=20;VARIANT_TYPE=snp;X_HIL=0;X_IL=1\tGT:PL\t{gt}:{pl}")
        text = ("\n".join(lines) + "\n").encode()
        ls, ri = cpu_index(text)
        ref_seq = "ACGT" * 2000
        host = PM.SplitPlan(hdr, cols, ref_seq)
        try:
            host.build(np.frombuffer(text, dtype=np.uint8), ls, ri)
            want = ("ok", [g.origin for g in host.groups], [len(g.rows) for g in host.groups])
        except Exception as e:  # noqa: BLE001
            want = (type(e).__name__, None, None)
        dev = PM.make_split_plan(hdr, cols, ref_seq, 0, reuse=dev)
        try:
            dev.build(np.frombuffer(text, dtype=np.uint8), ls, ri)
            got = ("ok", [int(o) for o in dev.origins], [int(r) for r in dev.n_rows])
        except Exception as e:  # noqa: BLE001
            got = (type(e).__name__, None, None)
        assert want == got, (it, want[0], got[0])
        seen.add(want[0])
    assert {"ok", "ValueError", "KeyError"} <= seen, seen


# ---------------------------------------------------------------- the reference's own unit-test tables, through the kernels
def _kat_line(pos, alleles, gt, hom, extra_info=""):
    """One record: `hom` = hom-alt PL per ALT (sets the allele order), every other PL entry 500 (300 for 0/0)."""
    n = len(alleles)
    pl = [500] * (n * (n + 1) // 2)
    pl[0] = 300
    for i, v in enumerate(hom, start=1):
        pl[i * (i + 1) // 2 + i] = v
    info = "DP=20;QD=1.00;VARIANT_TYPE=snp;X_HIL=0;X_HIN=.;X_IC=NA;X_IL=." + extra_info
    return f"chrM1\t{pos}\t.\t{alleles[0]}\t{','.join(alleles[1:])}\t50\t.\t{info}\tGT:DP:GQ:PL\t{gt}:20:10:{','.join(map(str, pl))}"


_KAT_FILLERS = [_kat_line(100, ("A", "C", "G"), "1/2", (10, 20)),                      # a plain multi-allelic single
                _kat_line(200, ("GAT", "G"), "0/1", (10,)).replace("X_IL=.", "X_IL=2"),  # a deletion ...
                _kat_line(201, ("A", "C", "*"), "1/2", (10, 20)),                       # ... and the row it spans
                _kat_line(300, ("A", "C"), "0/1", (10,))]  # a later record closes the cluster (the reference never flushes the last one)


def _kat_rows(dev_plan, lines, ref_seq, pos):
    text = ("\n".join(sorted(lines, key=lambda ln: int(ln.split("\t")[1]))) + "\n").encode()
    ls, ri = cpu_index(text)
    dev_plan.set_reference(ref_seq)
    out = dev_plan.build(np.frombuffer(text, dtype=np.uint8), ls, ri).tobytes().decode().split("\n")[:-1]
    rows = [ln.split("\t") for ln in out[int(dev_plan.kept.sum()):]]
    rows = [r for r in rows if int(r[1]) == pos]
    return [dict(alleles=(r[3], r[4]), info=dict(kv.split("=", 1) for kv in r[7].split(";")), gt=r[9].split(":")[0],
                 pl=tuple(int(v) for v in r[9].split(":")[3].split(","))) for r in rows]


def test_reference_unit_test_tables_through_the_kernels(case, gpu_ctx):
    """ugbio_filtering tests/unit/test_multiallelics.py:14-120: the in-code tables of select_overlapping_variants,
    encode_gt_for_allele_subset, select_pl_for_allele_subset, indel_classify_subset (with and without the spanning
    deletion) and classify_hmer_indel_relative, each reached through a record whose PLs select the table's allele pair."""
    from tests.test_multiallelics_cpu import (REF_KAT_HMER, REF_KAT_HMER_REF, REF_KAT_INDEL_CLASS, REF_KAT_INDEL_CLASS_SPANDEL,
                                              REF_KAT_OVERLAP)

    _ds, _tr, hdr, cols = case
    ref_seq = REF_KAT_HMER_REF + "ACGT" * 100
    dev = PM.DeviceSplitPlan(hdr, cols, ref_seq)

    def order_for(pair, n):
        """(gt, hom-alt PLs) that make `pair` the first row (0, k) or the second row (a, b) of the split."""
        hom = [400] * (n - 1)
        if pair[0] == 0:
            hom[pair[1] - 1] = 5
            return f"0/{pair[1]}", hom
        hom[pair[0] - 1], hom[pair[1] - 1] = 5, 10
        return f"{pair[0]}/{pair[1]}", hom

    # select_overlapping_variants: the table's records, one '*' cluster and one multi-allelic single
    lines = []
    for al, p in zip(REF_KAT_OVERLAP["alleles"], REF_KAT_OVERLAP["positions"]):
        n = len(al)
        x_il = f"X_IL={max(1, len(al[0]) - 1)}"
        lines.append(_kat_line(p, tuple(al), "0/1", [10 * (i + 1) for i in range(n - 1)]).replace("X_IL=.", x_il))
    text = ("\n".join(lines) + "\n").encode()
    ls, ri = cpu_index(text)
    dev.build(np.frombuffer(text, dtype=np.uint8), ls, ri)
    n_singles = int(dev.stats[0])
    assert [[int(o)] for o in dev.origins[:n_singles]] + [[int(o) for o in dev.origins[n_singles:]]] == REF_KAT_OVERLAP["expected"]

    # encode_gt_for_allele_subset / select_pl_for_allele_subset (the reachable rows of the tables)
    rows = _kat_rows(dev, _KAT_FILLERS + [_kat_line(22, ("A", "C", "G"), "0/1", (20, 60))], ref_seq, 22)
    assert [r["gt"] for r in rows] == ["0/1", "0/0"]                      # (0, 1) over (1, 2) -> (0, 0)
    rows = _kat_rows(dev, _KAT_FILLERS + [_kat_line(22, ("A", "C", "G"), "1/1", (20, 60))], ref_seq, 22)
    assert rows[1]["gt"] == "0/0"                                          # (1, 1) over (1, 2) -> (0, 0)
    rows = _kat_rows(dev, _KAT_FILLERS + [_kat_line(22, ("A", "C", "G"), "1/2", (20, 60))], ref_seq, 22)
    assert rows[1]["gt"] == "0/1"                                          # (1, 2) over (1, 2) -> (0, 1)
    line = _kat_line(22, ("A", "C", "G"), "1/2", (20, 60)).rsplit(":", 1)[0] + ":0,10,20,40,50,60"
    rows = _kat_rows(dev, _KAT_FILLERS + [line], ref_seq, 22)
    assert rows[1]["pl"] == (0, 30, 40)                                    # (0,10,20,40,50,60) over (1, 2)
    line = _kat_line(22, ("A", "C", "G", "T"), "0/3", (20, 60, 140)).rsplit(":", 1)[0] + ":0,10,20,40,50,60,100,120,130,140"
    rows = _kat_rows(dev, _KAT_FILLERS + [line], ref_seq, 22)
    assert rows[0]["alleles"] == ("A", "T") and rows[0]["pl"] == (0, 100, 140)   # ... over (0, 3)

    # indel_classify_subset
    for alleles, pair, (ic, il) in REF_KAT_INDEL_CLASS:
        gt, hom = order_for(pair, len(alleles))
        rows = _kat_rows(dev, _KAT_FILLERS + [_kat_line(22, alleles, gt, hom)], ref_seq, 22)
        r = rows[0 if pair[0] == 0 else 1]
        assert r["alleles"] == (alleles[pair[0]], alleles[pair[1]])
        assert (r["info"]["X_IC"], r["info"]["X_IL"]) == (ic[0], "." if il[0] is None else str(il[0])), (alleles, pair, r["info"])
    # ... with the spanning deletion: x_il = (4, 5) on the deletion's line
    head = _kat_line(21, ("GA", "G"), "0/1", (10,)).replace("X_IL=.", "X_IL=4,5")
    for alleles, pair, (ic, il) in REF_KAT_INDEL_CLASS_SPANDEL:
        if pair == (0, 2):
            continue  # (REF, '*') is never a row: '*' is forced to be the weakest allele (spandel.py:11-63)
        gt, hom = order_for((0, 1) if pair[0] == 0 else (1, 2), len(alleles))
        rows = _kat_rows(dev, [_KAT_FILLERS[0], head, _kat_line(22, alleles, "0/1" if pair[0] == 0 else "1/1", hom)], ref_seq, 22)
        r = rows[0 if pair[0] == 0 else 1]
        want_alt = alleles[pair[1]] if alleles[pair[1]] != "*" else "*" * (len(alleles[pair[0]]) + 1)
        assert r["alleles"] == (alleles[pair[0]], want_alt)
        assert (r["info"]["X_IC"], r["info"]["X_IL"]) == (ic[0], "." if il[0] is None else str(il[0])), (alleles, pair, r["info"])

    # classify_hmer_indel_relative at position 22 of the table's reference (the biallelic first row of the table is no split)
    for alleles, pair, (nuc, length) in REF_KAT_HMER[1:]:
        star = "*" in alleles
        gt, hom = order_for(pair, len(alleles))
        others = [_KAT_FILLERS[0], head] if star else _KAT_FILLERS
        rows = _kat_rows(dev, others + [_kat_line(22, alleles, "1/1" if star else gt, hom)], ref_seq, 22)
        r = rows[1]
        assert (r["info"]["X_HIN"], r["info"]["X_HIL"]) == (nuc, str(length)), (alleles, pair, r["info"])
    dev.close()
