#!/usr/bin/env python
"""Print the handful of ncu raw metrics we read for each captured kernel. usage: ncu_summary.py <raw csv>"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_alu.sum', 'smsp__inst_executed_pipe_fma.sum',
        'smsp__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_xu.sum', 'smsp__inst_executed_pipe_cbu.sum',
        'smsp__inst_executed_pipe_adu.sum', 'smsp__inst_executed_pipe_uniform.sum']
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('---')
    for w in want:
        if w in idx:
            print(f"  {w} = {r[idx[w]]} {units[idx[w]]}")
    st = sorted(((float(r[idx[h]] or 0), h) for h in stalls), reverse=True)[:7]
    print("  top stalls (warps per issue): " + ", ".join(
        f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for v, h in st))
