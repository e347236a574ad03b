"""--treat_multiallelics split / merge: device kernels (csrc/multiallelic.cu) against the Python model of the same
algorithm (UGVC_MA_HOST=1), on one synthetic contig dense with multi-allelic sites and deletion clusters
(tests/multiallelic_data.py).  Run on a GPU box:

    python scripts/bench_multiallelic.py [contig_bp] > gpurun_out/ma_bench.json

Prints one JSON line: records, groups, split rows, seconds of build and merge on both sides (the device numbers
include the copies of text / index / reference to the device and of the scored text back), identical output
(QD spelled differently, same double)."""
import json
import os
import re
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import multiallelic_data as MD  # noqa: E402
from variantcalling_b200 import lib, multiallelics as PM  # noqa: E402
from variantcalling_b200 import model_compiler as MC  # noqa: E402
from variantcalling_b200.vcf_header import VcfHeader  # noqa: E402


def norm_qd(line: bytes) -> bytes:
    def f(m):
        v = m.group(1)
        return b"QD=" + (v if v in (b".", b"", b"inf") else repr(float(v)).encode())
    return re.sub(rb"QD=([^;\t]*)", f, line)


def main():
    contig_bp = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    MD.CONTIGS = {"chrM1": contig_bp}
    t0 = time.perf_counter()
    ds = MD.generate(21, n_custom=3)
    gen_s = time.perf_counter() - t0
    hdr = VcfHeader(ds["header_text"])
    cols = hdr.loader_columns(ds["customs"])
    text = np.frombuffer(ds["text"], dtype=np.uint8)
    # the index pass of the tool: K1 with the model-less plan
    idx_ctx = lib.Context(0)
    idx_ctx.load_plan(MC.compile_plan_no_model(hdr).blob)
    idx_ctx.reserve(text.size + 4096, len(ds["lines"]) + 128, 1)
    t0 = time.perf_counter()
    idx = idx_ctx.filter_batch(text, 30.0)
    index_s = time.perf_counter() - t0
    ls, ri = idx["line_start"], idx["recinfo"]
    ref = ds["ref"]["chrM1"]
    out = {"records": int(idx["n_records"]), "text_bytes": int(text.size), "generate_s": round(gen_s, 2), "index_pass_s": round(index_s, 4)}

    dev = PM.DeviceSplitPlan(hdr, cols, ref)
    dev.build(text, ls, ri)  # warm-up: allocations
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        got = dev.build(text, ls, ri)
        best = min(best, time.perf_counter() - t0)
    out["device_build_s"] = round(best, 4)
    out["groups"], out["split_rows"] = int(dev.origins.size), int(dev.n_rows.sum())
    lik = np.random.default_rng(1).dirichlet(np.ones(3), size=int(dev.stats[2] + dev.stats[3]))
    n0 = dev.launch_count()
    dev.build(text, ls, ri)
    dev.merge(lik)
    out["device_launches_per_contig"] = dev.launch_count() - n0
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        m_dev = dev.merge(lik)
        best = min(best, time.perf_counter() - t0)
    out["device_merge_s"] = round(best, 4)

    host = PM.SplitPlan(hdr, cols, ref)
    t0 = time.perf_counter()
    want = host.build(text, ls, ri)
    out["python_build_s"] = round(time.perf_counter() - t0, 3)
    t0 = time.perf_counter()
    m_host = host.merge(lik)
    out["python_merge_s"] = round(time.perf_counter() - t0, 3)
    a, b = want.tobytes().split(b"\n"), got.tobytes().split(b"\n")
    out["same_rows"] = len(a) == len(b) and all(norm_qd(x) == norm_qd(y) for x, y in zip(a, b))
    out["same_merge"] = bool(np.array_equal(m_host, m_dev))
    out["build_speedup"] = round(out["python_build_s"] / max(out["device_build_s"], 1e-9), 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
