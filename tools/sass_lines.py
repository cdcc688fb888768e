#!/usr/bin/env python3
"""Static SASS instruction count by CUDA source line for one kernel of liborbx.so (no GPU needed).

    python tools/sass_lines.py <kernel substring> [lib.so] [--top N]

Extracts the cubin (cuobjdump -xelf), disassembles it with line info (nvdisasm -g -c) and prints, per source line of
orbx_kernels.cu, how many SASS instructions were generated for it (inlined callees are attributed to the innermost
line).  Used to budget a kernel's phases before spending GPU time (profiles/README.md).
"""
import collections, os, re, subprocess, sys, tempfile

def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    top = 40
    if "--top" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1]); args = [a for a in args if a != str(top)]
    pat = args[0]
    lib = args[1] if len(args) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                      "orbslam2_with_quadrics_b200", "liborbx.so")
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
        cub = max((os.path.join(d, f) for f in os.listdir(d)), key=os.path.getsize)
        out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    infn = False
    line = None
    counts = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    total = 0
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            infn = pat in m.group(1)
            continue
        if ln.startswith("\t.section") or ln.startswith(".section"):
            infn = False
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = int(m.group(2)) if m.group(1).endswith("orbx_kernels.cu") else (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m:
            counts[line] += 1
            ops[line][m.group(1)] += 1
            total += 1
    print("kernel ~", pat, " static SASS instructions:", total)
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orbslam2_with_quadrics_b200", "csrc",
                            "orbx_kernels.cu")).read().splitlines()
    for line, c in counts.most_common(top):
        text = src[line - 1].strip()[:70] if isinstance(line, int) and 0 < line <= len(src) else str(line)
        o = " ".join("%s=%d" % kv for kv in ops[line].most_common(6))
        print("%5d  L%-5s %-70s | %s" % (c, line if isinstance(line, int) else "-", text, o))

if __name__ == "__main__":
    main()
