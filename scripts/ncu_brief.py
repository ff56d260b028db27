#!/usr/bin/env python
"""Short summary of an `ncu --page raw --csv` dump (first kernel row): time, DRAM traffic, pipe utilisation, stalls.

    python scripts/ncu_brief.py gpurun_out/tc_topk_d10_raw.csv
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for r in rows[2:2 + int(sys.argv[2]) if len(sys.argv) > 2 else 3]:
    print("-" * 60)
    for key in KEYS:
        if key in hdr:
            i = hdr.index(key)
            print("%-70s %s %s" % (key, r[i], units[i]))
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(r[i] or 0) >= 0.15:
            print("%-70s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall "), r[i]))
