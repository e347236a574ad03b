#!/usr/bin/env python
"""Turn a round-2 GPU session directory (gpurun_out/<tag>/, scripts/gpu_session_r2.sh) into the tracked files
under profiles/: the bench line, the ncu launch list and its shares, the per-kernel summary of the `--set full`
captures, and profiles/ncu_traffic.json (DRAM bytes of the captured launch + the records it parsed) that
bench.py reads for roofline.traffic.
usage: save_profiles_r2.py <tag> [prefix]      e.g.  save_profiles_r2.py r2b r2"""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tag = sys.argv[1]
prefix = sys.argv[2] if len(sys.argv) > 2 else "r2"
src = os.path.join(ROOT, "gpurun_out", tag)
dst = os.path.join(ROOT, "profiles")
NCU_RECORDS_TOTAL, NCU_LAUNCH_INDEX = 4_000_000, 2  # gpu_session_r2.sh: bench.py --records 4000000, ncu -s 2 -c 1

if os.path.exists(f"{src}/bench_n1.json"):
    shutil.copy(f"{src}/bench_n1.json", f"{dst}/{prefix}_bench_n1_50M.json")
if os.path.exists(f"{src}/box.txt"):
    shutil.copy(f"{src}/box.txt", f"{dst}/{prefix}_box.txt")
# ---- launch list
if os.path.exists(f"{src}/launches.csv"):
    shutil.copy(f"{src}/launches.csv", f"{dst}/{prefix}_launches.csv")
    rows = [r for r in csv.reader(open(f"{src}/launches.csv")) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    hot = [k for k in agg if k.startswith(("k0_", "k1_", "k2_", "k3_"))]
    tot = sum(agg[k][1] for k in hot)
    with open(f"{dst}/{prefix}_launch_shares.txt", "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --records 8000000 (first 300 launches)\n"
                "# cold-cache serialised launch times: compare SHARES of the hot path, not absolutes\n")
        for k in sorted(hot, key=lambda k: -agg[k][1]):
            f.write(f"{k:22s} launches {agg[k][0]:4d}  total {agg[k][1] / 1e3:9.3f} ms  share of hot path {100 * agg[k][1] / tot:5.1f}%\n")
        f.write("# other launches in the capture (input generation): " +
                ", ".join(f"{k.split('<')[0]} x{v[0]}" for k, v in agg.items() if k not in hot) + "\n")
# ---- full captures
import bench  # noqa: E402  (contig layout of the captured launch)

rng = bench.contig_record_ranges(NCU_RECORDS_TOTAL)[NCU_LAUNCH_INDEX]
records = rng[2] - rng[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = {}
out_rows = []
for fn in sorted(os.listdir(src)):
    if not (fn.startswith("prof_") and fn.endswith("_raw.csv")):
        continue
    rows = list(csv.reader(open(os.path.join(src, fn))))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        st = sorted(((float(r[idx[h]] or 0), h) for h in stalls), reverse=True)[:5]
        out_rows.append([r[idx[x]] + ("" if x == "Kernel Name" else " " + units[idx[x]]) for x in want] + ["; ".join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for v, h in st)]
            + [str(records)])
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
        b = sum(float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        traffic[name] = {"dram_bytes": b, "records": records, "bytes_per_record": b / records,
                         "duration_us_under_ncu": float(r[idx["gpu__time_duration.sum"]].replace(",", "")),
                         "source": f"{prefix}_ncu_full_summary.csv ({fn})"}
if out_rows:
    with open(f"{dst}/{prefix}_ncu_full_summary.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(want + ["top stalls (warps per issue)", "records parsed by the captured launch"])
        w.writerows(out_rows)
    old = json.load(open(f"{dst}/ncu_traffic.json")) if os.path.exists(f"{dst}/ncu_traffic.json") else {}
    old.update(traffic)
    json.dump(old, open(f"{dst}/ncu_traffic.json", "w"), indent=1)
    for k, v in traffic.items():
        print(f"{k}: {v['bytes_per_record']:.0f} DRAM bytes per record ({v['duration_us_under_ncu']:.0f} us under ncu)")
if os.path.exists(f"{dst}/{prefix}_bench_n1_50M.json"):
    d = json.load(open(f"{dst}/{prefix}_bench_n1_50M.json"))
    print("value %.1fM e2e %.1fM" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6), d["roofline"]["kernel"],
          round(d["roofline"]["frac"], 4), d["roofline"]["stage_ms_per_launch"], d["roofline"]["path"])
