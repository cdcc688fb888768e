"""CPU restatement of Frame::ComputeBoW (reference src/Frame.cc:395-402) = DBoW2's
TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup = 4)
(Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1194, the per-feature descent :1217-1259, FORB::distance FORB.cpp:81-101,
BowVector::addWeight / normalize BowVector.cpp:34-46, :62-84, FeatureVector::addFeature FeatureVector.cpp:31-45) for the
ORB vocabulary's settings: TF-IDF weighting, L1 scoring (so the vector is L1-normalised).

TEST INFRASTRUCTURE: the checker of ``orbx_compute_bow``; only tests/ may import it.  Pinned against the reference's own
lines compiled against a stub (oracle/_ref/libbowref.so, ``ref_transform`` below) by tests/test_bow_oracle.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def descend(voc, d, levelsup):
    """One feature down the tree (:1217-1259): returns (word id, weight, node id at level L - levelsup)."""
    cs, ci, nd = voc["child_start"], voc["child_items"], voc["node_desc"]
    nid_level = voc["L"] - levelsup
    nid = 0                                   # root when nid_level <= 0 (:1227)
    final, level = 0, 0
    while True:
        level += 1
        kids = ci[cs[final]:cs[final + 1]]
        dist = _POP[nd[kids] ^ d[None, :]].sum(axis=1)
        final = int(kids[int(np.argmin(dist))])          # first strictly smaller wins = first minimum (:1244-1248)
        if level == nid_level:
            nid = final
        if cs[final + 1] == cs[final]:
            break
    return int(voc["node_word"][final]), float(voc["node_weight"][final]), nid


def transform(voc, desc, levelsup=4):
    """Returns (bow_ids uint32[], bow_values float64[], fv_nodes uint32[], fv_features uint32[]): the BowVector in map
    order and the FeatureVector as (node, feature) pairs in map / push_back order."""
    bow, fv = {}, {}
    for i in range(len(desc)):
        wid, w, nid = descend(voc, desc[i], levelsup)
        if w > 0:                                          # not stopped (:1157)
            bow[wid] = bow[wid] + np.float64(w) if wid in bow else np.float64(w)     # addWeight: += in feature order
            fv.setdefault(nid, []).append(i)
    ids = np.array(sorted(bow), np.uint32)
    vals = np.array([bow[k] for k in sorted(bow)], np.float64)
    norm = np.float64(0.0)
    for v in vals:                                         # BowVector::normalize(L1): sequential sum in map order
        norm = norm + abs(v)
    if norm > 0.0:
        vals = vals / norm
    fn = np.array([k for k in sorted(fv) for _ in fv[k]], np.uint32)
    ff = np.array([j for k in sorted(fv) for j in fv[k]], np.uint32)
    return ids, vals, fn, ff


# ----------------------------------------------------------------------------- the reference's own lines
_ref = None


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libbowref.so"))


def ref_build() -> bool:
    if not os.path.isdir("/root/reference/Thirdparty/DBoW2/DBoW2"):
        return ref_available()
    subprocess.run(["sh", os.path.join(_HERE, "build_bow_ref.sh")], check=True, stdout=subprocess.DEVNULL)
    return True


def ref_transform(voc, desc, levelsup=4):
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libbowref.so"))
        _ref.bowref_transform.restype = C.c_int
    n = len(desc)
    d = np.ascontiguousarray(desc, np.uint8)
    a = [np.ascontiguousarray(voc["child_start"], np.int32), np.ascontiguousarray(voc["child_items"], np.int32),
         np.ascontiguousarray(voc["node_desc"], np.uint8), np.ascontiguousarray(voc["node_weight"], np.float64),
         np.ascontiguousarray(voc["node_word"], np.int32)]
    ids, vals = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.float64)
    fn, ff = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.uint32)
    nfv = C.c_int(0)
    p = lambda x: C.c_void_p(x.ctypes.data)
    k = _ref.bowref_transform(C.c_int(voc["n_nodes"]), p(a[0]), p(a[1]), p(a[2]), p(a[3]), p(a[4]), C.c_int(voc["L"]), C.c_int(n),
                              p(d), C.c_int(levelsup), p(ids), p(vals), p(fn), p(ff), C.byref(nfv))
    return ids[:k].copy(), vals[:k].copy(), fn[:nfv.value].copy(), ff[:nfv.value].copy()


def ref_last_transform_seconds() -> float:
    """Wall time of the transform() call inside the last ref_transform (without the stub's tree set-up)."""
    _ref.bowref_last_transform_seconds.restype = C.c_double
    return float(_ref.bowref_last_transform_seconds())
