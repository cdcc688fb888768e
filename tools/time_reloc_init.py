#!/usr/bin/env python3
"""Timing of the two remaining Frame-side matchers through the C ABI beside the reference's own lines on one host thread
(oracle/_ref/libstereoref.so, when built):  orbx_search_by_projection_kf  (src/ORBmatcher.cc:1472-1599)  and
orbx_search_for_initialization  (:405-520).  One JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
import numpy as np
from oracle import match_oracle
from orbslam2_with_quadrics_b200 import ORBextractor
from orbslam2_with_quadrics_b200 import frames as fr
from orbslam2_with_quadrics_b200 import match_cases as mc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("mmg", os.path.join(ROOT, "tests", "golden", "make_match_golden.py"))
mmg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mmg)
name = sys.argv[1] if len(sys.argv) > 1 else "rgbd_1080p"
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 16
w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS[name]
K4, D = (1050.0, 1050.0, w / 2.0, h / 2.0), (0.0, 0.0, 0.0, 0.0)
gx = ORBextractor(nf, sf, nl, it, mt, max_batch=NQ)
res = gx.extract_batch([fr.cluttered_scene(w, h, 40 + (i % 4)) for i in range(NQ)])
grids = gx.undistort_grid(K4, D)
rng = np.random.default_rng(3)
out = {"config": name, "queries": NQ}

def view(f):
    xy, start, items, bounds = grids[f]
    k = res[f][0]
    return dict(xy_un=xy, cur_octave=k["octave"].astype(np.int32), cur_angle=k["angle"].astype(np.float32), desc=res[f][1],
                cell_start=start, cell_items=items, bounds=bounds, sf=np.asarray(gx.GetScaleFactors(), np.float32), nlevels=nl)

# ---- relocalisation matcher
qs, scs = [], []
for f in range(NQ):
    cf = view(f)
    Tc = mc.pose(rng)
    kf = mc.make_reloc_keyframe(rng, cf["xy_un"], cf["cur_octave"], cf["cur_angle"], cf["desc"], K4, Tc, len(cf["desc"]), cf["sf"])
    sc = dict(**kf, Tcw_cur=Tc, xy_un=cf["xy_un"], cur_octave=cf["cur_octave"], cur_angle=cf["cur_angle"], desc=cf["desc"],
              cell_start=cf["cell_start"], cell_items=cf["cell_items"], bounds=cf["bounds"], K4=K4, sf=cf["sf"])
    inr, pred = match_oracle.kf_prepare(sc["valid"], sc["world"], sc["min_dist"], sc["max_dist"], Tc, mmg.kf_log_scale(cf), nl)
    qs.append(dict(cur_frame=f, search=((sc["valid"] == 1) & (inr > 0)).astype(np.uint8), world=sc["world"], pred_level=pred,
                   mp_desc=sc["mp_desc"], kf_angle=sc["kf_angle"], Tcw_cur=Tc, cur_held=sc["cur_held"].astype(np.int32)))
    scs.append(sc)
for _ in range(3): r = gx.search_by_projection_kf(qs, K4, 10.0, 100, True)
t0 = time.perf_counter(); K = 20
for _ in range(K): r = gx.search_by_projection_kf(qs, K4, 10.0, 100, True)
out["kf_projection"] = {"ms_per_batch_e2e": (time.perf_counter() - t0) / K * 1e3, "points_per_query": int(len(scs[0]["valid"])),
                        "nmatches_q0": int(r[0][0]), "rounds_q0": int(r[0][2])}
if match_oracle.ref_has("matchref_search_by_projection_kf"):
    t0 = time.perf_counter()
    n, m, _, _ = match_oracle.ref_search_by_projection_kf(log_scale_factor=mmg.kf_log_scale(view(0)), th=10.0, orb_dist=100, **scs[0])
    out["kf_projection"]["reference_lines_ms_per_query_1_thread"] = (time.perf_counter() - t0) * 1e3
    out["kf_projection"]["identical_q0"] = bool(n == r[0][0] and np.array_equal(m, r[0][1]))

# ---- initialisation matcher
qi, sci = [], []
for f in range(NQ):
    cf = view(f)
    f1 = mc.make_initial_frame(rng, cf["xy_un"], cf["cur_octave"], cf["cur_angle"], cf["desc"], len(cf["desc"]))
    sc = dict(**f1, xy_un2=cf["xy_un"], octave2=cf["cur_octave"], angle2=cf["cur_angle"], desc2=cf["desc"], cell_start=cf["cell_start"],
              cell_items=cf["cell_items"], bounds=cf["bounds"])
    qi.append(dict(cur_frame=f, octave1=sc["octave1"], angle1=sc["angle1"], desc1=sc["desc1"], prev_matched=sc["prev_matched"]))
    sci.append(sc)
for _ in range(3): r = gx.search_for_initialization(qi, 0.9, True, 100)
t0 = time.perf_counter()
for _ in range(K): r = gx.search_for_initialization(qi, 0.9, True, 100)
out["initialization"] = {"ms_per_batch_e2e": (time.perf_counter() - t0) / K * 1e3, "f1_keypoints": int(len(sci[0]["octave1"])),
                         "nmatches_q0": int(r[0][0])}
t0 = time.perf_counter()
r1 = gx.search_for_initialization(qi[:1], 0.9, True, 100)
out["initialization"]["ms_single_query_e2e"] = (time.perf_counter() - t0) * 1e3
if match_oracle.ref_has("matchref_search_for_initialization"):
    t0 = time.perf_counter()
    n, m, p = match_oracle.ref_search_for_initialization(sf=view(0)["sf"], nnratio=0.9, check_orientation=True, window=100, **sci[0])
    out["initialization"]["reference_lines_ms_per_query_1_thread"] = (time.perf_counter() - t0) * 1e3
    out["initialization"]["identical_q0"] = bool(n == r[0][0] and np.array_equal(m, r[0][1]) and np.array_equal(p, r[0][2]))
print(json.dumps(out))
