#!/usr/bin/env python3
"""Runs every device-side consumer once on a batch of 1080p frames (after a warm-up round); the target of the
`ncu --set full` capture summarised in profiles/r01_consumer_kernels_ncu_full.txt."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orbslam2_with_quadrics_b200 import ORBextractor, Vocabulary
from orbslam2_with_quadrics_b200 import frames as fr
from orbslam2_with_quadrics_b200 import match_cases as mc
from orbslam2_with_quadrics_b200 import vocabulary as vc

B = 16
w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["rgbd_1080p"]
K4, D = (1050.0, 1050.0, 959.5, 539.5), (0.05, -0.11, 0.0004, -0.0003, 0.02)
gx = ORBextractor(nf, sf, nl, it, mt, max_batch=B, download_pyramid=False)
gv = Vocabulary(vc.random_vocabulary(10, 6, seed=1))
imgs = [fr.cluttered_scene(w, h, 3000 + i) for i in range(4)]
res = gx.extract_batch([imgs[i % 4] for i in range(B)])
grids = gx.undistort_grid(K4, D)
rng = np.random.default_rng(77)
ql, qp, qb = [], [], []
for f in range(B):
    kps, desc = res[f]
    octv, ang = kps["octave"].astype(np.int32), kps["angle"].astype(np.float32)
    Tc = mc.pose(rng)
    ql.append(dict(cur_frame=f, Tcw_cur=Tc, Tcw_last=mc.pose(rng), **mc.make_last_frame(rng, grids[f][0], octv, ang, desc, K4, Tc, len(kps), nl)))
    qp.append(dict(cur_frame=f, **mc.make_local_points(rng, grids[f][0], octv, desc, 2 * len(kps), nl)))
for rnd in range(2):
    gx.search_by_projection(ql, K4, 0.0, 0.0, 15.0, True)
    gx.search_local_points(qp, 3.0)
    bows = gx.compute_bow(gv)
    if not qb:
        for f in range(B):
            kps, desc = res[f]
            kf = mc.make_keyframe(rng, desc, kps["angle"].astype(np.float32), len(kps))
            # the KeyFrame's FeatureVector: reuse the frame's own node assignment for the features it was copied from is
            # not possible (descriptors differ), so run the KeyFrame descriptors through the same vocabulary on the GPU path's
            # sibling: a second extractor is overkill here -- take the frame's vector (same nodes, valid indices)
            qb.append(dict(cur_frame=f, kf_fv_nodes=bows[f][2], kf_fv_features=np.minimum(bows[f][3], len(kps) - 1), **kf))
    out = gx.search_by_bow(qb, 0.7, True)
print("ok", [n for n, _ in out[:4]])
