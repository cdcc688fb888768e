#!/usr/bin/env python3
"""Regenerates tests/golden/stereo_golden.json: SHA-256 of the mvuRight / mvDepth float32 bit patterns that the
stereo oracle (oracle/stereo_oracle.py, pinned against the reference's own Frame::ComputeStereoMatches lines) produces
for the oracle extractor's output on frames.stereo_pair(); plus match counts as human-readable KATs.

    python tests/golden/make_stereo_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orb_oracle, stereo_oracle                   # noqa: E402
from orbslam2_with_quadrics_b200 import frames as fr           # noqa: E402

# (config, seed, mbf = baseline * fx, fx): Examples/Stereo/EuRoC.yaml:13,33 and KITTI00-02.yaml:12,24; the VGA case
# reuses the EuRoC camera
CASES = [("stereo_euroc", 1234, 47.90639384423901, 435.2046959714599),
         ("stereo_euroc", 2234, 47.90639384423901, 435.2046959714599),
         ("stereo_kitti", 1234, 386.1448, 718.856),
         ("mono_tum", 1234, 47.90639384423901, 435.2046959714599)]


def run_case(name, seed, mbf, fx):
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS[name]
    left, right = fr.stereo_pair(w, h, seed)
    ex = orb_oracle.ORBextractor(nf, sf, nl, it, mt)
    rl, rr = ex(left), ex(right)
    mb = float(np.float32(mbf) / np.float32(fx))
    u, d, sad = stereo_oracle.compute_stereo_matches(rl.keypoints, rl.descriptors, rl.pyramid, rr.keypoints, rr.descriptors,
                                                     rr.pyramid, ex.GetScaleFactors(), ex.GetInverseScaleFactors(), mbf, mb)
    return rl, rr, mb, u, d, sad


def digest(u, d):
    return {"n": int(len(u)), "matched": int((u >= 0).sum()),
            "u_right_sha256": hashlib.sha256(np.ascontiguousarray(u, np.float32).tobytes()).hexdigest(),
            "depth_sha256": hashlib.sha256(np.ascontiguousarray(d, np.float32).tobytes()).hexdigest()}


if __name__ == "__main__":
    out = {}
    for name, seed, mbf, fx in CASES:
        *_x, u, d, sad = run_case(name, seed, mbf, fx)
        out["%s/%d" % (name, seed)] = digest(u, d)
        print(name, seed, out["%s/%d" % (name, seed)]["matched"], "of", len(u))
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "stereo_golden.json"), "w"), indent=1, sort_keys=True)
