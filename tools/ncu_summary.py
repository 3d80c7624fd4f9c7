#!/usr/bin/env python
"""Print the metrics that matter from an .ncu-rep (read here, on the CPU box): python tools/ncu_summary.py rep [--source N]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_drain_per_warp_active.pct"]
for r in rows[2:]:
    print("==", r[H.index("Kernel Name")][:80])
    for w in want:
        if w in H:
            i = H.index(w)
            print("  %-75s %12s %s" % (w, r[i], U[i]))
if len(sys.argv) > 3 and sys.argv[2] == "--source":
    n = int(sys.argv[3])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = [i for i, r in enumerate(rows) if "Source" in r and any("Sampl" in c for c in r)]
    if hdr:
        H = rows[hdr[0]]
        si = [i for i, c in enumerate(H) if c.startswith("Warp Stall Sampling (All")]
        si = si[0] if si else [i for i, c in enumerate(H) if "Sampl" in c][0]
        body = [r for r in rows[hdr[0] + 1:] if len(r) > si and r[si].replace('.', '').isdigit()]
        tot = sum(float(r[si]) for r in body)
        print("total samples", tot)
        top = sorted(enumerate(body), key=lambda t: -float(t[1][si]))[:n]
        for idx, r in sorted(top):
            print("%5d %6.2f%%  %s" % (idx, 100 * float(r[si]) / tot, r[H.index("Source")][:110]))
