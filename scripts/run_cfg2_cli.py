#!/usr/bin/env python
"""BASELINE.json configs[1] at full size, file to file, through the drop-in CLI:

  5 M-record synthetic single-sample VCF (bgzip + tabix, written here from the device generator)
  + logistic-regression model pickle  ->  `filter_variants_pipeline.run`  ->  filtered .vcf.gz + .tbi

and the size-independent checks of the parity plan: record count, LOW_SCORE count equal to an
independent device-resident pass over the same text, the head of chr1 and a mid-file contig (found
through both .tbi files) equal to the oracle line for line.  Prints one JSON line.
"""
import argparse
import json
import os
import pickle
import sys
import tempfile
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEED = 20260922


def build_lr_model(n_custom: int, n_train: int = 20000):
    import pandas as pd
    from sklearn.linear_model import LogisticRegression

    from oracle import ref_pipeline as R  # training frame only (CPU, like train_models_pipeline)
    from oracle.vcf_reader import OracleVariantFile
    from variantcalling_b200 import synth
    from variantcalling_b200 import transformers as T
    from variantcalling_b200.tprep_constants import VcfType

    header, lines, labels = synth.generate(synth.SynthSpec(n_records=n_train, n_custom=n_custom, seed=1984))
    customs = synth.custom_annotation_names(n_custom)
    df = R.get_vcf_df(OracleVariantFile(synth.vcf_text(header, lines)), None, customs)
    tr = T.get_transformer(VcfType.SINGLE_SAMPLE, [c.lower() for c in customs])
    with pd.option_context("future.infer_string", False):
        x = tr.fit_transform(R.harness_float_columns(df)).to_numpy(dtype=np.float64)
    model = LogisticRegression(max_iter=400)
    model.fit(x, labels)
    return model, tr, customs


def contig_of_records(total: int):
    """record index -> contig index, as the device generator lays the contigs out (bench.py)."""
    from variantcalling_b200.synth import CONTIG_LENGTHS

    lens = list(CONTIG_LENGTHS.values())
    genome, cum, first = sum(lens), 0, []
    for ln in lens:
        first.append(total * cum // genome)
        cum += ln
    return list(CONTIG_LENGTHS), np.array(first, dtype=np.int64)


def oracle_lines(header_text: str, record_text: bytes, model, tr, customs):
    from oracle import ref_pipeline as R
    from oracle.vcf_reader import OracleVariantFile

    return R.filter_variants(OracleVariantFile(header_text.encode() + record_text), model, tr, custom_annotations=customs)["lines"]


def main():  # noqa: PLR0915
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=5_000_000)
    ap.add_argument("--n-custom", type=int, default=5)
    ap.add_argument("--batch-records", type=int, default=1_000_000)
    ap.add_argument("--check-records", type=int, default=15000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--workdir", default=None)
    ap.add_argument("--host-io", action="store_true", help="time the host inflate / writer / deflate path (--host_io) too")
    ap.add_argument("--runs", type=int, default=2, help="timed CLI runs (the first pays page-cache and allocation warm-up)")
    args = ap.parse_args()
    import torch

    from variantcalling_b200 import bgzf_io, lib
    from variantcalling_b200 import filter_variants_pipeline as fvp
    from variantcalling_b200 import model_compiler as MC

    out = {"config": "BASELINE configs[1]: 5M-record synthetic single-sample VCF, logistic regression, CLI file to file",
           "records": args.records}
    work = args.workdir or tempfile.mkdtemp(prefix="cfg2_")
    t0 = time.perf_counter()
    model, tr, customs = build_lr_model(args.n_custom)
    header_text = lib.synth_header(args.n_custom)
    out["model_fit_s"] = time.perf_counter() - t0

    # ---- input file: device generator -> host -> BGZF + .tbi (index pass gives POS / REF length per record)
    t0 = time.perf_counter()
    ctx = lib.Context(0)
    plan = MC.compile_plan(header_text, tr, model, customs)
    ctx.load_plan(plan.blob)
    idx = lib.Context(0)
    idx.load_plan(MC.compile_plan_no_model(header_text).blob)
    B = args.batch_records
    cap = int(B * 470 * 1.3) + (8 << 20)
    idx.reserve(cap, B + 128, 1)
    ctx.reserve(cap, B + 128, 1)
    d_text = torch.empty(cap, dtype=torch.uint8, device="cuda")
    vcf = os.path.join(work, "in.vcf.gz")
    w = bgzf_io.BgzfWriter(vcf, level=1, n_threads=args.threads)
    hdr_bytes = header_text.encode()
    w.write(hdr_bytes)
    names, first = contig_of_records(args.records)
    pos_all, end_all, ustart_all, uend_all = [], [], [], []
    n_low_device, n_text, head_text = 0, 0, None
    for b0 in range(0, args.records, B):
        nb = min(B, args.records - b0)
        nbytes = ctx.synth_device(SEED, b0, nb, args.records, args.n_custom, d_text.data_ptr(), cap - 64)
        torch.cuda.synchronize()
        host = d_text[:nbytes].cpu().numpy()
        res = idx.filter_batch(host)
        assert res["n_records"] == nb, (res["n_records"], nb)
        scored = ctx.filter_batch(host, 30.0, want_recinfo=False)  # independent of the CLI's lanes / batching
        n_low_device += int(scored["low_score"].sum())
        if head_text is None:
            head_text = host[: int(res["line_start"][min(nb, args.check_records)])].tobytes()
        base = w.uoffset
        w.write(host)
        ri, ls = res["recinfo"], res["line_start"]
        pos_all.append(ri["pos"].astype(np.int64) - 1)
        end_all.append(ri["pos"].astype(np.int64) - 1 + np.maximum(1, (ri["flags"] >> 8).astype(np.int64)))
        ustart_all.append(base + ls[:-1])
        uend_all.append(base + ls[1:])
        n_text += nbytes
    w.close()
    cat = np.concatenate
    contig_idx = (np.searchsorted(first, np.arange(args.records), side="right") - 1).astype(np.int32)
    present = np.unique(contig_idx)
    remap = np.full(len(names), -1, dtype=np.int32)
    remap[present] = np.arange(present.size)
    vs = w.virtual_offsets(cat(ustart_all))
    ve = w.virtual_offsets(cat(uend_all) - 1) + np.uint64(1)
    bgzf_io.write_tbi(vcf + ".tbi", bgzf_io.build_tbi([names[i] for i in present], remap[contig_idx], cat(pos_all),
                                                     cat(end_all), vs, ve))
    idx.close()
    ctx.close()
    del d_text
    out.update(input_text_bytes=n_text, input_file_bytes=os.path.getsize(vcf), input_build_s=time.perf_counter() - t0,
               n_low_score_device_pass=n_low_device)
    mpath = os.path.join(work, "model.pkl")
    with open(mpath, "wb") as fh:
        pickle.dump({"xgb": model, "transformer": tr}, fh)

    # ---- the CLI, timed wall clock file to file
    dst = os.path.join(work, "out.vcf.gz")
    argv = ["--input_file", vcf, "--model_file", mpath, "--output_file", dst]
    for c in customs:
        argv += ["--custom_annotations", c]
    if args.threads:
        argv += ["--io_threads", str(args.threads)]
    walls = []
    for _ in range(max(1, args.runs)):
        t0 = time.perf_counter()
        totals = fvp.run(argv)
        walls.append(time.perf_counter() - t0)
    wall = min(walls)
    out.update(cli_wall_s=wall, cli_wall_s_runs=walls, variants_per_s_file_to_file=args.records / wall,
               output_file_bytes=os.path.getsize(dst), cli_totals={k: int(v) for k, v in totals.items()},
               host_cores=os.cpu_count(), io="device (ugvc_filter_bgzf) where it applies")
    # the same job as its own process (`python ugvc filter_variants_pipeline ...`): interpreter start-up, imports and the
    # CUDA context included
    import subprocess
    dst_p = os.path.join(work, "out_proc.vcf.gz")
    cmd = [sys.executable, os.path.join(ROOT, "ugvc"), "filter_variants_pipeline"] + [a if a != dst else dst_p for a in argv]
    t0 = time.perf_counter()
    rc = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False).returncode
    out["cli_process_wall_s"] = time.perf_counter() - t0
    out["cli_process_rc"] = rc
    out["variants_per_s_process"] = args.records / out["cli_process_wall_s"]
    out["timing_note"] = ("cli_wall_s: filter_variants_pipeline.run(argv) called in this process (best of the runs; imports and the "
                          "CUDA context are already warm); cli_process_wall_s: one fresh `python ugvc ...` process, everything included")
    if args.host_io:
        dst_h = os.path.join(work, "out_host.vcf.gz")
        t0 = time.perf_counter()
        fvp.run([a if a != dst else dst_h for a in argv] + ["--host_io"])
        out["cli_wall_s_host_io"] = time.perf_counter() - t0
        out["variants_per_s_host_io"] = args.records / out["cli_wall_s_host_io"]
        import gzip
        with gzip.open(dst) as a, gzip.open(dst_h) as b:  # the two paths write the same text
            same = True
            while same:
                x, y = a.read(1 << 24), b.read(1 << 24)
                same = x == y
                if not x:
                    break
        out["device_io_text_equals_host_io_text"] = bool(same)

    # ---- checks
    checks = {}
    try:
        checks["record_count"] = totals["n_records"] == args.records
        checks["low_score_equals_device_pass"] = totals["n_low_score"] == n_low_device
        in_idx, out_idx = bgzf_io.read_tbi(vcf + ".tbi"), bgzf_io.read_tbi(dst + ".tbi")
        checks["same_contigs_indexed"] = list(in_idx) == list(out_idx)
        # head of the file
        first_contig = list(out_idx)[0]
        got = bgzf_io.inflate(dst, *out_idx[first_contig]).tobytes().decode().split("\n")
        want = oracle_lines(header_text, head_text, model, tr, customs)
        checks["head_lines_equal_oracle"] = got[: len(want)] == want
        checks["head_lines_checked"] = len(want)
        # a contig in the middle of the file, reached through both indexes
        mid = list(out_idx)[len(out_idx) // 2]
        src_mid = bgzf_io.inflate(vcf, *in_idx[mid]).tobytes()
        cut = 0
        for _ in range(min(args.check_records // 3, src_mid.count(b"\n"))):
            cut = src_mid.index(b"\n", cut) + 1
        want_mid = oracle_lines(header_text, src_mid[:cut], model, tr, customs)
        got_mid = bgzf_io.inflate(dst, *out_idx[mid]).tobytes().decode().split("\n")
        checks["mid_contig"] = mid
        checks["mid_contig_lines_equal_oracle"] = got_mid[: len(want_mid)] == want_mid and want_mid[0].startswith(mid + "\t")
        checks["mid_contig_lines_checked"] = len(want_mid)
        n_out = sum(bgzf_io.inflate(dst, *out_idx[c]).tobytes().count(b"\n") for c in out_idx)
        checks["records_in_output_file"] = n_out
    except Exception:  # noqa: BLE001
        checks["error"] = traceback.format_exc()[-1500:]
    out["checks"] = checks
    print(json.dumps(out))


if __name__ == "__main__":
    main()
