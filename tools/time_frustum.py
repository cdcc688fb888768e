#!/usr/bin/env python3
"""Timing of orbx_is_in_frustum (Frame::isInFrustum + MapPoint::PredictScale, src/Frame.cc:269-325) through the C ABI beside the
reference's own lines on one host thread (oracle/_ref/libstereoref.so, when built).  One JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import match_oracle
from orbslam2_with_quadrics_b200 import ORBextractor
from orbslam2_with_quadrics_b200 import match_cases as mc

NQ = int(sys.argv[1]) if len(sys.argv) > 1 else 16
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 4000          # local map points per frame
nl, sf = 8, 1.2
K4, bounds = (1050.0, 1050.0, 960.0, 540.0), (0.0, 1920.0, 0.0, 1080.0)
gx = ORBextractor(2000, sf, nl, 20, 7, max_batch=NQ)
sfs = np.asarray(gx.GetScaleFactors(), np.float32)
lsf = float(np.float32(np.log(np.float32(sf))))
rng = np.random.default_rng(5)
qs = []
for _ in range(NQ):
    Tc = mc.pose(rng, scale_r=0.3)
    sc = mc.make_frustum_points(rng, K4, bounds, Tc, NP, sfs)
    sc["Tcw"] = Tc
    qs.append(sc)
for _ in range(3):
    res = gx.is_in_frustum(qs, K4, 40.0, bounds, lsf, 0.5)
t0 = time.perf_counter()
R = 20
for _ in range(R):
    res = gx.is_in_frustum(qs, K4, 40.0, bounds, lsf, 0.5)
ms = (time.perf_counter() - t0) / R * 1e3
t0 = time.perf_counter()
for _ in range(R):
    one = gx.is_in_frustum(qs[:1], K4, 40.0, bounds, lsf, 0.5)
ms1 = (time.perf_counter() - t0) / R * 1e3
out = {"queries": NQ, "points_per_query": NP, "in_view_per_query": int(np.mean([r[0].sum() for r in res])),
       "orbx_ms_per_call": ms, "orbx_us_per_query": ms / NQ * 1e3, "orbx_ms_single_query": ms1,
       "path": "Python wrapper over the C ABI: staging into pinned memory, H2D, frustum_kernel, D2H, host copies of the result arrays"}
if match_oracle.ref_has("matchref_is_in_frustum"):
    t0 = time.perf_counter()
    for sc in qs[:4]:
        match_oracle.ref_is_in_frustum(sc["consider"], sc["world"], sc["normal"], sc["min_dist"], sc["max_dist"], sc["Tcw"], K4, 40.0, bounds, lsf, nl, 0.5)
    out["reference_lines_ms_per_query_one_thread"] = (time.perf_counter() - t0) / 4 * 1e3
print(json.dumps(out))
