#!/usr/bin/env python3
"""Aggregate an ncu SASS source-page CSV by CUDA source line.
usage: ncu_lines.py <ncu-rep> <kernel-regex> [top]   (needs the same liborbx.so that was profiled)"""
import csv, re, subprocess, sys, os, collections, tempfile
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "orbslam2_with_quadrics_b200", "liborbx.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.startswith("orbx_kernels.") and f.endswith(".cubin")][0]
# with inlining, each instruction is preceded by a chain of "//## File ..., line N inlined at ..." comments; the LAST
# one of the chain is the line in the kernel's own body (outermost call site)
sass = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kname = rows[0][1]
mangled_hint = re.sub(r"[^A-Za-z0-9_]", "", kname.split("(")[0].split("::")[-1].split("<")[0])
hdr = rows[1]
ia, ii, isrc, ist = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
# address -> line from nvdisasm
sec = None; line = 0; amap = {}; inl = ""
for ln in sass.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m: sec = m.group(1); continue
    m = re.search(r'//## File ".*?", line (\d+)(.*)', ln)
    if m:
        line = int(m.group(1)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and sec and mangled_hint in sec:
        amap.setdefault(sec, {})[int(m.group(1), 16)] = line
# choose the section with most addresses matching
best = max(amap.items(), key=lambda kv: len(kv[1]))[1] if amap else {}
if len(amap) > 1:
    # template instantiations: pick by first-row count equality is hard; take the one whose size matches
    n = len(rows) - 2
    best = min(amap.values(), key=lambda d: abs(len(d) - n))
base = None
agg = collections.Counter(); samp = collections.Counter(); tot = 0
for r in rows[2:]:
    try:
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    except ValueError:
        continue
    if base is None: base = a
    off = a - base
    n = int(float(r[ii] or 0)); s = int(float(r[ist] or 0))
    l = best.get(off, -1)
    agg[l] += n; samp[l] += s; tot += n
src = open(os.path.join(ROOT, "orbslam2_with_quadrics_b200", "csrc", "orbx_kernels.cu")).read().splitlines()
print("kernel:", kname[:90]); print("total warp instructions:", tot)
for l, n in agg.most_common(top):
    txt = src[l - 1].strip()[:100] if 0 < l <= len(src) else "?"
    print("%6.2f%% %10d samples=%6d  L%-4d %s" % (100.0 * n / max(tot, 1), n, samp[l], l, txt))
