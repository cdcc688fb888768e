#!/usr/bin/env python3
"""Runs orbx_search_by_projection_device on a batch of (LastFrame, CurrentFrame) queries (1080p frames); meant to be run
under `ncu --metrics gpu__time_duration.sum` to separate the two kernels' durations from the host staging."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orbslam2_with_quadrics_b200 import ORBextractor
from orbslam2_with_quadrics_b200 import frames as fr
from orbslam2_with_quadrics_b200 import match_cases as mc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["rgbd_1080p"]
gx = ORBextractor(nf, sf, nl, it, mt, max_batch=B, download_pyramid=False)
imgs = [fr.cluttered_scene(w, h, 3000 + i) for i in range(4)]
res = gx.extract_batch([imgs[i % 4] for i in range(B)])
K4, D = (1050.0, 1050.0, 959.5, 539.5), (0.05, -0.11, 0.0004, -0.0003, 0.02)
grids = gx.undistort_grid(K4, D)
rng = np.random.default_rng(77)
qs = []
for f in range(B):
    kps, desc = res[f]
    Tc = mc.pose(rng)
    last = mc.make_last_frame(rng, grids[f][0], kps["octave"].astype(np.int32), kps["angle"].astype(np.float32), desc, K4, Tc,
                              len(kps), nl)
    qs.append(dict(cur_frame=f, Tcw_cur=Tc, Tcw_last=mc.pose(rng), **last))
prepared = gx._projection_queries(qs)
for _ in range(5):
    gx.search_by_projection_device(prepared, K4, 0.0, 0.0, 15.0, True)
gx.synchronize()
out = gx.search_by_projection(qs, K4, 0.0, 0.0, 15.0, True)
print("nmatches", [n for n, _, _ in out][:4], "rounds", [r for _, _, r in out][:4])
