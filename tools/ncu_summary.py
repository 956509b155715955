#!/usr/bin/env python3
"""Prints the metrics we care about from an .ncu-rep (development tool): python tools/ncu_summary.py file.ncu-rep [kernel-index]"""
import csv, subprocess, sys
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[2 + idx]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
d = dict(zip(h, zip(u, v)))
for k in want:
    if k in d:
        print("%-75s %-12s %s" % (k, d[k][0], d[k][1]))
for k in h:
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
        try:
            if float(d[k][1]) > 0.05:
                print("%-75s %-12s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall:"), d[k][0], d[k][1]))
        except ValueError:
            pass
for k in h:
    if "pipe" in k and "pct_of_peak_sustained_active" in k and ".avg." in k and k not in want:
        try:
            if float(d[k][1]) > 1.0:
                print("%-75s %-12s %s" % (k, d[k][0], d[k][1]))
        except ValueError:
            pass
