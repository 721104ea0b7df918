#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): ncu -i X --page raw --csv | tools/ncu_summary.py [pattern...]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
pats = sys.argv[1:] or ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct", "gpu__dram_throughput",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct",
    "issue_stalled", "sm__pipe_tensor", "sm__inst_executed_pipe", "local_ld", "local_st", "bank_conflict", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__cycles_active.avg", "sm__inst_executed.avg.per_cycle_elapsed"]
for d in data:
    print("=== ", d[hdr.index("Kernel Name")], "grid", d[hdr.index("Grid Size")], "block", d[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats):
            v = d[i]
            if v in ("", "0", "n/a"): continue
            print(f"  {h:95s} {v} {units[i]}")
