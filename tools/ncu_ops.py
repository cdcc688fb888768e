#!/usr/bin/env python3
"""Dynamic opcode histogram + pipe utilisation of one kernel from an ncu report (--set full --import-source on).
usage: ncu_ops.py <file.ncu-rep> <kernel-regex> [top]"""
import csv, re, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ii, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
c = collections.Counter(); tot = 0
for r in rows[2:]:
    try: n = int(r[ii])
    except (ValueError, IndexError): continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
    op = m.group(1) if m else "?"
    c[op] += n; tot += n
print("# %s: %d warp instructions" % (rows[0][1][:60], tot))
for op, n in c.most_common(top): print("%-12s %11d %5.1f%%" % (op, n, 100.0 * n / tot))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
for r in rr[2:3]:
    o = []
    for i, n in enumerate(h):
        m = re.match(r"sm__(inst_executed_pipe_\w+|pipe_\w+_cycles_active)\.avg\.pct_of_peak_sustained_active$", n)
        if m and float(r[i] or 0) > 1: o.append("%s=%s" % (m.group(1).replace("inst_executed_pipe_", "i_").replace("_cycles_active", "_cyc"), r[i][:5]))
    for key in ("gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size"):
        if key in h: o.append("%s=%s" % (key.split(".")[0].split("__")[1], r[h.index(key)]))
    print("# pipes:", " ".join(o))
