#!/usr/bin/env python3
"""Regenerates tests/golden/bow_golden.json: SHA-256 digests of the BowVector (word ids, float64 value bits) and the
FeatureVector pairs that the REFERENCE'S OWN DBoW2 lines (oracle/_ref/libbowref.so, built by oracle/build_bow_ref.sh)
produce for seeded synthetic vocabularies and descriptor sets.

    python tests/golden/make_bow_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bow_oracle                                   # noqa: E402
from orbslam2_with_quadrics_b200 import vocabulary as vc        # noqa: E402

# (k, L, irregular, levelsup, seed, descriptors)
CASES = [(10, 3, False, 1, 1, 1500), (10, 4, True, 2, 2, 1200), (5, 5, True, 4, 3, 2000), (10, 4, False, 4, 4, 2000),
         (10, 4, False, 6, 5, 300), (3, 6, True, 4, 6, 900)]


def descriptors(voc, seed, n):
    """Perturbed leaf descriptors (so that the descent is not pure noise), with repeats (words hit several times) and a
    tail of random ones."""
    rng = np.random.default_rng(seed)
    leaves = np.where(voc["node_word"] >= 0)[0]
    pick = rng.choice(leaves[:max(8, len(leaves) // 3)], n)
    d = voc["node_desc"][pick].copy()
    d ^= ((rng.random(d.shape) < 0.02).astype(np.uint8) << rng.integers(0, 8, d.shape).astype(np.uint8)).astype(np.uint8)
    tail = n // 10
    d[-tail:] = rng.integers(0, 256, (tail, 32), dtype=np.uint8)
    return d


def digest(res):
    ids, vals, fn, ff = res
    h = lambda a, t: hashlib.sha256(np.ascontiguousarray(a, t).tobytes()).hexdigest()
    return {"n_words": int(len(ids)), "n_features": int(len(fn)), "word_ids": h(ids, np.uint32), "word_values_bits": h(vals, np.float64),
            "fv_nodes": h(fn, np.uint32), "fv_features": h(ff, np.uint32)}


def key(case):
    return "k%d/L%d/%s/up%d/seed%d/n%d" % (case[0], case[1], "irregular" if case[2] else "regular", case[3], case[4], case[5])


if __name__ == "__main__":
    bow_oracle.ref_build()
    assert bow_oracle.ref_available()
    out = {}
    for case in CASES:
        voc = vc.random_vocabulary(case[0], case[1], seed=case[4], irregular=case[2])
        d = descriptors(voc, case[4], case[5])
        out[key(case)] = digest(bow_oracle.ref_transform(voc, d, case[3]))
        print(key(case), out[key(case)]["n_words"], out[key(case)]["n_features"])
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bow_golden.json"), "w"), indent=1, sort_keys=True)
