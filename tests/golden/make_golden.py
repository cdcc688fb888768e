#!/usr/bin/env python3
"""Regenerates tests/golden/orb_golden.json from the oracle (cv2 4.13.0 primitives, canonical
rules B-1/B-2).  Digests are SHA-256 over: the padded pyramid bytes (levels concatenated), the
ordered (x, y, size, response, octave) int32 tuples, the angle float32 bit patterns and the
descriptor bytes; plus the first keypoint and the descriptor byte sum as human-readable KATs.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orb_oracle                                  # noqa: E402
from orbslam2_with_quadrics_b200 import frames as fr           # noqa: E402


def digest(kps, desc, pyramid=None):
    ints = np.stack([kps["x"].view(np.uint32), kps["y"].view(np.uint32), kps["size"].view(np.uint32),
                     kps["response"].view(np.uint32), kps["octave"].view(np.uint32),
                     kps["class_id"].view(np.uint32)], axis=1) if len(kps) else np.zeros((0, 6), np.uint32)
    d = {"n": int(len(kps)),
         "keypoints_sha256": hashlib.sha256(np.ascontiguousarray(ints).tobytes()).hexdigest(),
         "angles_sha256": hashlib.sha256(np.ascontiguousarray(kps["angle"]).tobytes()).hexdigest(),
         "descriptors_sha256": hashlib.sha256(np.ascontiguousarray(desc).tobytes()).hexdigest(),
         "descriptor_byte_sum": int(desc.astype(np.int64).sum())}
    if len(kps):
        k = kps[0]
        d["first_keypoint"] = [float(k["x"]), float(k["y"]), float(k["size"]), float(k["angle"]), float(k["response"]), int(k["octave"])]
    if pyramid is not None:
        hsh = hashlib.sha256()
        for p in pyramid:
            hsh.update(np.ascontiguousarray(p).tobytes())
        d["pyramid_sha256"] = hsh.hexdigest()
    return d


def cases():
    for name, (w, h, nf, sf, nl, it, mt, nimg) in fr.CONFIGS.items():
        for seed in (1234, 2234):
            yield "%s/seed%d" % (name, seed), fr.cluttered_scene(w, h, seed), (nf, sf, nl, it, mt)
    yield "noise640x480/seed5", fr.noise_frame(640, 480, 5), (1000, 1.2, 8, 20, 7)
    yield "checker640x480/seed5", fr.checker_frame(640, 480, 5), (1000, 1.2, 8, 20, 7)
    yield "flat640x480", fr.flat_frame(640, 480), (1000, 1.2, 8, 20, 7)
    yield "odd517x333/seed77", np.ascontiguousarray(fr.cluttered_scene(700, 500, 77)[10:343, 20:537]), (500, 1.2, 6, 20, 7)


def main():
    out = {"_about": "oracle digests; regenerate with tests/golden/make_golden.py (cv2 %s)" % orb_oracle.cv2.__version__}
    for key, img, args in cases():
        r = orb_oracle.ORBextractor(*args)(img)
        out[key] = digest(r.keypoints, r.descriptors, r.pyramid)
        out[key]["image_sha256"] = hashlib.sha256(img.tobytes()).hexdigest()
        out[key]["args"] = list(args)
        print(key, out[key]["n"])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "orb_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
