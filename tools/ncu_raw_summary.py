#!/usr/bin/env python
"""Summary of `ncu --page raw --csv` output in the format of profiles/*_ncu_full.txt: the last profiled
launch of every kernel name (or all with --all), a fixed list of metrics.
usage: python tools/ncu_raw_summary.py raw.csv [--all] > profiles/xxx.txt"""
import csv
import sys

METRICS = """launch__grid_size launch__block_size launch__registers_per_thread launch__shared_mem_per_block_dynamic
gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct sm__warps_active.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active smsp__inst_executed.sum sm__icc_request_hit_rate.pct
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum launch__occupancy_limit_shared_mem
launch__occupancy_limit_registers""".split()

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
if "--all" not in sys.argv:
    last = {}
    for r in body:
        last[r[ix["Kernel Name"]]] = r
    body = list(last.values())
for r in body:
    print("Kernel Name  %s" % r[ix["Kernel Name"]])
    for m in METRICS:
        if m in ix:
            print("%-92s %s %s" % (m, r[ix[m]], units[ix[m]]))
    print()
