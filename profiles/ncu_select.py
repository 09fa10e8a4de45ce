#!/usr/bin/env python
"""Selected columns of an `ncu -i X.ncu-rep --page raw --csv` export (one row per captured launch).

usage: ncu -i full.ncu-rep --page raw --csv | python profiles/ncu_select.py > profiles/rNN_ncu_full.csv"""
import csv
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
cols = [h for h in KEEP if h in hdr] + stall
idx = [hdr.index(h) for h in cols]
w = csv.writer(sys.stdout)
w.writerow([c.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "") for c in cols])
for r in rows[1:]:
    w.writerow([r[i] for i in idx])
