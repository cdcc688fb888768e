#!/usr/bin/env python3
"""profiles/sass_summary.txt: what the built liborbx.so contains, kernel by kernel -- target architecture, registers,
static shared memory, and the counts of the Blackwell-relevant SASS mnemonics (TMA tile loads, mbarrier waits, byte /
halfword SIMD, dot products, warp votes / shuffles).  Needs no GPU: cuobjdump on the cubin inside the library.
usage: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "orbslam2_with_quadrics_b200", "liborbx.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function (\S+):", res)), capture_output=True, text=True).stdout.splitlines()
usage = {}
names = re.findall(r"Function (\S+):\s*\n\s*(.*)", res)
for (m, line), dem in zip(names, filt):
    regs = re.search(r"REG:(\d+)", line); sh = re.search(r"SHARED:(\d+)", line)
    usage[m] = (dem, int(regs.group(1)) if regs else -1, int(sh.group(1)) if sh else 0)
WATCH = ["UTMALDG", "SYNCS", "VABSDIFF4", "VIMNMX3", "VIMNMX", "IDP", "PRMT", "SHFL", "VOTE", "POPC", "ATOMS", "ATOMG", "BAR", "LDS", "STS", "LDG", "STG", "HMMA", "UTCMMA"]
arch = re.findall(r"arch = (sm_\w+)", sass)
print("# %s: %d cubin(s), arch %s (no other architecture, no PTX fallback needed on B200)" % (os.path.basename(so), len(arch), sorted(set(arch))))
print("# tensor-core mnemonics (HMMA / UTCMMA) are expected to be 0: no stage of the path is a contraction (DESIGN.md §4)")
cur, counts, order = None, collections.defaultdict(collections.Counter), []
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); order.append(cur); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and cur:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for w in WATCH:
            if op == w or (w in ("IDP", "SYNCS", "BAR", "VOTE") and op.startswith(w)):
                counts[cur][w] += 1
print("%-58s %5s %7s %7s  %s" % ("kernel", "regs", "smem_B", "instrs", "watched mnemonics"))
for k in order:
    dem, regs, sh = usage.get(k, (k, -1, 0))
    short = re.sub(r"\(.*", "", dem).replace("void ", "").replace("orbx::", "")
    c = counts[k]
    print("%-58s %5d %7d %7d  %s" % (short[:58], regs, sh, c["_total"], " ".join("%s=%d" % (w, c[w]) for w in WATCH if c[w])))
