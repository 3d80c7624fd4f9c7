#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` capture of k_fm_ransac into profiles/<tag>_fmat.txt:
headline metrics of the launch and the share of warp-stall samples per phase of the kernel (the SASS between two CTA barriers).
Usage: python tools/make_fmat_profile.py gpurun_out/fmat_<tag>.ncu-rep <tag> "<workload>" """
import csv
import os
import subprocess
import sys

rep, tag, workload = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
H, U, R = rr[0], rr[1], rr[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.max.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_dispatch_stall"]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sr = list(csv.reader(src.splitlines()))
hdr = sr[1]
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(r[1], int(r[si]), int(r[ie])) for r in sr[2:] if len(r) > si]
tot = sum(d[1] for d in data)
out = os.path.join(ROOT, "profiles", "%s_fmat.txt" % tag)
with open(out, "w") as f:
    f.write("ncu --set full --clock-control none --import-source on, one launch of k_fm_ransac: %s\n" % workload)
    f.write("(duration under ncu replay is not a bench value)\n\n")
    for w in want:
        if w in H:
            i = H.index(w)
            f.write("  %-82s %16s %s\n" % (w, R[i], U[i]))
    f.write("\nwarp-stall samples and executed warp instructions per phase (SASS between consecutive CTA barriers; the out-of-line\n"
            "IEEE division / sqrt / pow subroutines the compiler places after the last barrier are listed as 'subroutines')\n")
    seg = acc = ex = start = 0
    for i, (s, n, e) in enumerate(data):
        acc += n
        ex += e
        if "BAR.SYNC" in s:
            if acc > tot * 0.003:
                f.write("  phase %2d  sass[%5d..%5d]  samples %7d (%5.1f %%)  warp instructions %11d\n" % (seg, start, i, acc, 100.0 * acc / tot, ex))
            seg += 1
            acc = ex = 0
            start = i + 1
    f.write("  subroutines + Jacobi (after the last barrier) sass[%5d..]  samples %7d (%5.1f %%)  warp instructions %11d\n"
            % (start, acc, 100.0 * acc / tot, ex))
print(open(out).read())
