#!/usr/bin/env python
"""Turn the ncu artefacts of one `tools/profile_step.sh <tag>` run (gpurun_out/) into the tracked summaries under profiles/.

    python tools/make_profiles.py <tag>        # reads gpurun_out/launches_<tag>.csv, gpurun_out/step_<tag>.ncu-rep

Writes profiles/<tag>_launches.csv (every launch of the captured steps with its device time and share of the step),
profiles/<tag>_kernels.txt (per-kernel --set full metrics) and profiles/fast_traffic.json (DRAM bytes per k_fast launch,
read by bench.py for roofline.traffic)."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list
txt = open(os.path.join(G, "launches_%s.csv" % tag)).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
def short(n):
    n = n.split("(")[0]
    return n.split("::")[-1]
launches = [(short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3) for r in rows]
# one step = from one k_ingest to the next
starts = [i for i, l in enumerate(launches) if l[0] == "k_ingest"]
step = launches[starts[0]:starts[1]] if len(starts) > 1 else launches
total = sum(l[3] for l in step)
with open(os.path.join(P, "%s_launches.csv" % tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none; one step (64 frames 1920x1080) of `python bench.py --no-cpu --no-hamming`\n")
    f.write("# per-launch times are serialised and cold-cache: compare SHARES, not absolutes.  step total %.1f us\n" % total)
    f.write("kernel,grid,block,us,share_of_step\n")
    for k, g, b, us in step:
        f.write('%s,"%s","%s",%.2f,%.4f\n' % (k, g, b, us, us / total))
    agg = {}
    for k, g, b, us in step:
        agg[k] = agg.get(k, 0.0) + us
    f.write("# per kernel: " + "; ".join("%s %.1f us (%.1f%%)" % (k, v, 100 * v / total) for k, v in sorted(agg.items(), key=lambda t: -t[1])) + "\n")

# ---- full metrics
rep = os.path.join(G, "step_%s.ncu-rep" % tag)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
H, U = rr[0], rr[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle"]
fast = None
with open(os.path.join(P, "%s_kernels.txt" % tag), "w") as f:
    f.write("ncu --set full --clock-control none, one step (64 frames 1920x1080, 2000 kp) of `python bench.py --no-cpu --no-hamming`\n")
    f.write("(times under ncu replay are not bench values; see %s_launches.csv for shares)\n\n" % tag)
    for r in rr[2:]:
        name = short(r[H.index("Kernel Name")])
        f.write("== %s  grid %s block %s\n" % (name, r[H.index("Grid Size")], r[H.index("Block Size")]))
        for w in want:
            if w in H:
                i = H.index(w)
                f.write("  %-82s %16s %s\n" % (w, r[i], U[i]))
        f.write("\n")
        if name == "k_fast":
            def val(m):
                v = float(r[H.index(m)].replace(",", "")); u = U[H.index(m)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            fast = {"kernel": "k_fast", "frames_per_launch": 64, "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                    "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                    "warp_instructions_per_launch": float(r[H.index("smsp__inst_executed.sum")].replace(",", "")),
                    "source": "profiles/%s_kernels.txt (ncu --set full, gpurun_out/step_%s.ncu-rep)" % (tag, tag)}
if fast:
    json.dump(fast, open(os.path.join(P, "fast_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, "%s_launches.csv" % tag)).read())
