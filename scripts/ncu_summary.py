#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion uses.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [substring filters...]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
for k in keys:
    for i, h in enumerate(hdr):
        if h == k or h.endswith("." + k):
            print(f"{h[-80:]:80s} {units[i]:14s} {[d[i] for d in data]}")
            break
for i, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "average_warps" in h:
        vals = [d[i] for d in data]
        try:
            if max(float(v.replace(",", "")) for v in vals) >= 0.3:
                print(f"{h[-80:]:80s} {units[i]:14s} {vals}")
        except ValueError:
            pass
    for e in extra:
        if e in h:
            print(f"{h[-80:]:80s} {units[i]:14s} {[d[i] for d in data]}")
