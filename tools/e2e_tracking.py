#!/usr/bin/env python3
"""One tracked frame end to end, everything behind the C ABI: extraction of a batch of host frames, then the consumers that
run on the device-resident results -- UndistortKeyPoints + AssignFeaturesToGrid, SearchByProjection against the last frame
(TrackWithMotionModel), SearchByProjection against the local map (SearchLocalPoints), ComputeBoW -- with their results
copied back.  Prints frames/s of the extraction alone and of the whole chain (wall clock over K steps, one host thread).

    python tools/e2e_tracking.py [config] [frames_per_step] [steps]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C

import numpy as np

from orbslam2_with_quadrics_b200 import ORBextractor, Vocabulary, _capi
from orbslam2_with_quadrics_b200 import frames as fr
from orbslam2_with_quadrics_b200 import match_cases as mc
from orbslam2_with_quadrics_b200 import vocabulary as vc

name = sys.argv[1] if len(sys.argv) > 1 else "rgbd_1080p"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS[name]
K4 = (0.55 * w, 0.55 * w, 0.5 * w - 0.5, 0.5 * h - 0.5)
D = (0.05, -0.11, 0.0004, -0.0003, 0.02)
gx = ORBextractor(nf, sf, nl, it, mt, max_batch=B, download_pyramid=False)
gv = Vocabulary(vc.random_vocabulary(10, 6, seed=1))
imgs = [fr.cluttered_scene(w, h, 4000 + i) for i in range(4)]
import torch
pinned = torch.empty((B, h, w), dtype=torch.uint8).pin_memory()          # host frames in pinned memory, as bench.py's e2e leg
for i in range(B):
    pinned[i].copy_(torch.from_numpy(imgs[i % 4]))
batch = [pinned[i].numpy() for i in range(B)]
res = gx.extract_batch(batch)
grids = gx.undistort_grid(K4, D)
rng = np.random.default_rng(5)
q_last, q_local = [], []
for f in range(B):
    kps, desc = res[f]
    Tc = mc.pose(rng)
    last = mc.make_last_frame(rng, grids[f][0], kps["octave"].astype(np.int32), kps["angle"].astype(np.float32), desc, K4, Tc, len(kps), nl)
    q_last.append(dict(cur_frame=f, Tcw_cur=Tc, Tcw_last=mc.pose(rng), **last))
    q_local.append(dict(cur_frame=f, **mc.make_local_points(rng, grids[f][0], kps["octave"].astype(np.int32), desc, 2 * len(kps), nl)))


# everything marshalled once: the timed loop is C-ABI calls only (what a C++ caller pays)
L, H = gx._L, gx._h
n = len(batch)
ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in batch])
strides = (C.c_size_t * n)(*[im.strides[0] for im in batch])
eres = (_capi.OrbxResult * n)()
k4 = (C.c_float * 4)(*K4)
dd = (C.c_float * len(D))(*D)
gres = (_capi.OrbxGridResult * n)()
pq, keep1 = gx._projection_queries(q_last)
lq, keep2 = gx._local_queries(q_local)
pres = (_capi.OrbxProjectionResult * n)()
bres = (_capi.OrbxBowResult * n)()


def extraction_only():
    _capi.check(L.orbx_extract_batch(H, n, ptrs, w, h, strides, eres), H)


def chain():
    _capi.check(L.orbx_extract_batch(H, n, ptrs, w, h, strides, eres), H)
    _capi.check(L.orbx_undistort_grid(H, n, None, k4, dd, len(D), gres), H)
    _capi.check(L.orbx_search_by_projection(H, n, pq, k4, 0.0, 0.0, 15.0, 1, 1, 0, pres), H)
    _capi.check(L.orbx_search_local_points(H, n, lq, 3.0, 0.8, 0, pres), H)
    _capi.check(L.orbx_compute_bow(H, gv._v, n, None, 4, bres), H)


out = {}
for label, fn in (("extraction_only", extraction_only), ("extraction_plus_consumers", chain)):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(K):
        fn()
    dt = (time.perf_counter() - t0) / K
    out[label] = {"ms_per_step": dt * 1e3, "frames_per_s": B / dt}
# each consumer alone (after one extraction), same calls
extraction_only()
parts = {"orbx_undistort_grid": lambda: _capi.check(L.orbx_undistort_grid(H, n, None, k4, dd, len(D), gres), H),
         "orbx_search_by_projection": lambda: _capi.check(L.orbx_search_by_projection(H, n, pq, k4, 0.0, 0.0, 15.0, 1, 1, 0, pres), H),
         "orbx_search_local_points": lambda: _capi.check(L.orbx_search_local_points(H, n, lq, 3.0, 0.8, 0, pres), H),
         "orbx_compute_bow": lambda: _capi.check(L.orbx_compute_bow(H, gv._v, n, None, 4, bres), H)}
out["per_consumer_ms_per_step"] = {}
for label, fn in parts.items():
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(K):
        fn()
    out["per_consumer_ms_per_step"][label] = (time.perf_counter() - t0) / K * 1e3
out["config"] = {"workload": name, "frames_per_step": B, "steps": K, "host_threads": 1,
                 "consumers": ["orbx_undistort_grid", "orbx_search_by_projection (N map points per frame)",
                               "orbx_search_local_points (2N map points per frame)", "orbx_compute_bow (10^6-word tree)"],
                 "note": "C-ABI calls only inside the timed loop (queries marshalled once); every call returns its results in pinned host memory"}
print(json.dumps(out))
if os.path.isdir("gpurun_out"):
    json.dump(out, open("gpurun_out/r01_e2e_tracking_%s.json" % name, "w"), indent=1)
