#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters the roofline analysis uses.
Usage: python tools/ncu_summary.py report.ncu-rep [--stalls]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    print("----")
    for w in WANT:
        if w in idx:
            print(f"{w:75s} {r[idx[w]]} {units[idx[w]]}")
    if "--stalls" in sys.argv:
        vals = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") or \
               (h.startswith("smsp__average_warp_latency_issue_stalled") and h.endswith(".ratio")):
                try:
                    vals.append((float(r[idx[h]].replace(",", "")), h))
                except ValueError:
                    pass
        for v, h in sorted(vals, reverse=True)[:10]:
            print(f"   {v:10.3f} {h}")
