#!/usr/bin/env python
"""Turn the scratch outputs of a GPU run (gpurun_out/) into the tracked summaries under profiles/.
usage: save_profiles.py <tag>   (expects gpurun_out/{bench_<tag>_n1.json, launches_<tag>.csv, prof_<tag>.ncu-rep})"""
import collections
import csv
import json
import shutil
import subprocess
import sys

tag = sys.argv[1]
shutil.copy(f"gpurun_out/bench_{tag}_n1.json", f"profiles/{tag}_bench_n1_50M.json")
shutil.copy(f"gpurun_out/launches_{tag}.csv", f"profiles/{tag}_launches.csv")
rows = [r for r in csv.reader(open(f"gpurun_out/launches_{tag}.csv")) if len(r) > 5]
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
with open(f"profiles/{tag}_launch_shares.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --records 16000000 (first 300 launches)\n"
            "# cold-cache serialised launch times: compare SHARES of the hot path (K0..K3), not absolutes\n")
    for k in sorted(hot, key=lambda k: -agg[k][1]):
        f.write(f"{k:22s} launches {agg[k][0]:4d}  total {agg[k][1] / 1e3:9.3f} ms  share of hot path {100 * agg[k][1] / tot:5.1f}%\n")
    f.write("# other launches in the capture (input generation, fills): " +
            ", ".join(f"{k.split('<')[0]} x{v[0]}" for k, v in agg.items() if k not in hot) + "\n")
raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
with open(f"profiles/{tag}_ncu_full_summary.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(want + ["top stalls (warps per issue)"])
    w.writerow([units[idx[x]] for x in want] + [""])
    for r in rows[2:]:
        st = sorted(((float(r[idx[h]] or 0), h) for h in stalls), reverse=True)[:5]
        w.writerow([r[idx[x]] for x in want] + ["; ".join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for v, h in st)])
d = json.load(open(f"profiles/{tag}_bench_n1_50M.json"))
print("value %.1fM e2e %.1fM" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6), d["roofline"]["kernel"],
      round(d["roofline"]["frac"], 4), d["roofline"]["stage_ms_per_launch"], d["roofline"]["path"])
