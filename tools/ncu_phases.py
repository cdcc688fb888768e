#!/usr/bin/env python3
"""Executed warp instructions of one kernel by PHASE, from an ncu report (--set full --import-source on) and the liborbx.so
that was profiled (same sources).  An instruction is attributed to its innermost CUDA source line (nvdisasm -g); a line
inside the kernel body belongs to the latest preceding `// ---- <phase>` comment, a line inside an inlined helper
(fast_score_packed, fast_pretest_word, mbar_*, intrinsics headers) to that helper.
usage: ncu_phases.py <file.ncu-rep> <kernel-regex> <kernel name in the source, e.g. fast_strips_kernel>"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kre, kname = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "orbslam2_with_quadrics_b200", "liborbx.so")
src = open(os.path.join(ROOT, "orbslam2_with_quadrics_b200", "csrc", "orbx_kernels.cu")).read().splitlines()
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, check=True, capture_output=True)
    cub = max((os.path.join(d, f) for f in os.listdir(d)), key=os.path.getsize)
    sass = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first kernel of the page only
hdr = rows[1]
ia, ii = hdr.index("Address"), hdr.index("Instructions Executed")
body = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ia].startswith("0x"):
        break
    body.append((int(r[ia], 16), int(r[ii] or 0)))
base = body[0][0]
# the section of the same kernel in the local build: the one whose name contains kname and whose size matches
secs, sec, line, fname = {}, None, None, None
for ln in sass.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        sec = m.group(1) if kname in m.group(1) else None
        if sec: secs[sec] = []
        continue
    if ln.startswith("\t.section") or ln.startswith(".section"):
        sec = None
    if sec is None:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        fname, line = os.path.basename(m.group(1)), int(m.group(2)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        secs[sec].append((int(m.group(1), 16), fname, line))
best = min(secs.values(), key=lambda v: abs(len(v) - len(body)))
amap = {a: (f, l) for a, f, l in best}
# kernel body range and phase markers
k0 = next(i for i, t in enumerate(src) if re.search(r"\b%s\(" % kname, t) and "__global__" in "".join(src[max(0, i - 3):i + 1])) + 1
k1 = next(i for i in range(k0, len(src)) if src[i].startswith("}")) + 1
def helper_of(l):
    for i in range(l - 1, -1, -1):
        m = re.match(r"(?:template.*)?\s*__device__ __forceinline__ \S+ (\w+)\(", src[i])
        if m: return m.group(1)
        if src[i].startswith("}"): break
    return "other helper"
agg = collections.Counter(); tot = 0
for a, n in body:
    f, l = amap.get(a - base, (None, None))
    tot += n
    if f != "orbx_kernels.cu" or l is None:
        agg["intrinsics headers (shuffles, ballots, atomics, min/max)"] += n
    elif k0 <= l <= k1:
        lab = "prologue / per-strip bookkeeping"
        for i in range(l - 1, k0 - 1, -1):
            m = re.search(r"// ---- ([^:;.]+)", src[i])
            if m: lab = m.group(1).strip(); break
        agg[lab] += n
    else:
        agg["helper " + helper_of(l)] += n
print("# %s: %d warp instructions (first launch matching '%s')" % (kname, tot, kre))
for lab, n in agg.most_common():
    print("%6.2f%%  %12d  %s" % (100.0 * n / tot, n, lab))
