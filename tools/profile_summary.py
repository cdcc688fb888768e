#!/usr/bin/env python3
"""Compact per-kernel table from an ncu report (`--set full`): duration, DRAM bytes, issue activity, stalls.
usage: profile_summary.py <file.ncu-rep> > profiles/<name>.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
M = [("time_us", "gpu__time_duration.sum"), ("warp_inst", "smsp__inst_executed.sum"),
     ("issue_active_%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
     ("warps_active_%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
     ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
     ("dram_%peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     ("l1tex_%", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
     ("lts_%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
     ("st_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
     ("st_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
     ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
     ("st_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
     ("st_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
     ("st_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
     ("st_not_sel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio")]
ki = h.index("Kernel Name")
def val(r, name):
    if name not in h: return "n/a"
    i = h.index(name); v = r[i]; u = units[i]
    try:
        f = float(v.replace(",", ""))
    except ValueError:
        return v
    if name.startswith("dram__bytes"):
        f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
    if name == "gpu__time_duration.sum":
        f *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    return "%.2f" % f if f < 1e6 else "%.0f" % f
print("# source: %s (ncu --set full --clock-control none; cold-cache, serialised: compare shares, not absolutes)" % rep)
print("%-28s " % "kernel" + " ".join("%13s" % m[0] for m in M))
for r in data:
    print("%-28s " % r[ki].split("(")[0].replace("void ", "").replace("orbx::", "")[:28] + " ".join("%13s" % val(r, m[1]) for m in M))
