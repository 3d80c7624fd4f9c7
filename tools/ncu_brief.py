#!/usr/bin/env python
"""Condensed per-kernel summary of one or more .ncu-rep files (read here, on the CPU box), in the format of profiles/r2_jpeg.txt and
profiles/r2_bow.txt:  python tools/ncu_brief.py rep [rep ...] > profiles/<name>.txt"""
import csv
import subprocess
import sys

ROWS = [("duration", "gpu__time_duration.sum"), ("warp instructions", "smsp__inst_executed.sum"),
        ("active threads per instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
        ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warp slots occupied %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("registers/thread", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
        ("cluster", "launch__cluster_size"),
        ("DRAM read", "dram__bytes_read.sum"), ("DRAM written", "dram__bytes_write.sum"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
        ("L1 hit %", "l1tex__t_sector_hit_rate.pct")] + \
       [("stall %s / issue" % k, "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k)
        for k in ("long_scoreboard", "short_scoreboard", "wait", "barrier", "branch_resolving", "lg_throttle", "not_selected", "membar")]

for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    for r in rows[2:]:
        print("== " + r[H.index("Kernel Name")][:120])
        for label, key in ROWS:
            if key in H:
                print("  %-36s %s %s" % (label, r[H.index(key)], U[H.index(key)]))
        print()
