import re, sys
rows=[]
for ln in open(sys.argv[1]):
    m=re.match(r"\s*([\d.]+)%\s+(\d+) samples=\s*(\d+)\s+L(-?\d+)\s+(.*)",ln)
    if m: rows.append((float(m.group(1)),int(m.group(2)),int(m.group(3)),int(m.group(4)),m.group(5)))
tot=sum(r[1] for r in rows); ts=sum(r[2] for r in rows)
src=open('/root/repo/orbslam2_with_quadrics_b200/csrc/orbx_kernels.cu').read().splitlines()
def find(s, start=0):
    for i,l in enumerate(src):
        if i >= start and s in l: return i+1
    raise KeyError(s)
k0 = find("fast_strips_kernel(const __grid_constant__")
marks=[("fast_cell_path (inlined)", find("__device__ __forceinline__ void fast_cell_path")), ("legacy kernel", find("fast_cells_kernel(const __grid_constant__")),
("pretest helper", find("__device__ __forceinline__ uint32_t fast_pretest_word")),
("kernel prologue", k0), ("loop top/geometry", find("while (cur.x != ORBX_FS_NONE)", k0)), ("phase1 pretest", find("// ---- phase 1: aligned SIMD pre-test", k0)),
("phase1b scan", find("// ---- phase 1b: scan, then every lane appends", k0)), ("write loop", find("uint16_t* wq = queue + wbase + (incl - cnt);", k0)), ("phase2 score", find("// ---- phase 2: exact score; corners to the strip map", k0)),
("phase3 nms", find("// ---- phase 3: strict 3x3 NMS; kept corners -> bitmap", k0)), ("phase4 emit", find("// ---- phase 4: warp = cell, lane = scoring row", k0)), ("zeroing", find("// leave the score map and the bitmap all-zero", k0)), ("tail/cell_path call", find("// B6: map / bitmap zero", k0)), ("end", find("// DistributeOctTree (:539-763)"))]
agg={}; aggs={}
for pct,n,s,l,txt in rows:
    name="other"
    for (nm,st),(nm2,en) in zip(marks,marks[1:]):
        if st<=l<en: name=nm
    agg[name]=agg.get(name,0)+n; aggs[name]=aggs.get(name,0)+s
print("total warp instr (both launches)", tot)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1]): print("%-28s %6.2f%% instr  %6.2f%% samples"%(k,100*v/tot,100*aggs[k]/ts))
