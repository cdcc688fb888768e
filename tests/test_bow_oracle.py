"""CPU tests of SURVEY.md §8(f) row 4, Frame::ComputeBoW (reference src/Frame.cc:395-402 -> DBoW2
TemplatedVocabulary::transform, Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1194, :1217-1259): the restatement against
the golden digests made by the reference's own DBoW2 lines, and against those lines directly where they can be built."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle import bow_oracle
from orbslam2_with_quadrics_b200 import vocabulary as vc

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_bow_golden", os.path.join(HERE, "golden", "make_bow_golden.py"))
mbg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mbg)
GOLD = json.load(open(os.path.join(HERE, "golden", "bow_golden.json")))


@pytest.mark.parametrize("case", mbg.CASES)
def test_restatement_matches_golden_and_reference_lines(case):
    voc = vc.random_vocabulary(case[0], case[1], seed=case[4], irregular=case[2])
    d = mbg.descriptors(voc, case[4], case[5])
    res = bow_oracle.transform(voc, d, case[3])
    assert mbg.digest(res) == GOLD[mbg.key(case)]
    try:
        bow_oracle.ref_build()
    except Exception:
        pass
    if bow_oracle.ref_available():
        ref = bow_oracle.ref_transform(voc, d, case[3])
        assert all(np.array_equal(a, b) for a, b in zip(res, ref))
        assert np.array_equal(res[1].view(np.uint64), ref[1].view(np.uint64))          # float64 bits
    ids, vals, fn, ff = res
    assert np.all(np.diff(ids.astype(np.int64)) > 0) and abs(vals.sum() - 1.0) < 1e-12    # map order, L1-normalised
    assert len(fn) <= case[5] and np.all(np.diff(fn.astype(np.int64)) >= 0)
    assert len(ids) < len(fn)                                                           # words are hit more than once


def test_vocabulary_shape_of_orbvoc():
    voc = vc.random_vocabulary(10, 6, seed=1)
    assert voc["n_nodes"] == 1111111 and int((voc["node_word"] >= 0).sum()) == 10 ** 6
    assert voc["child_start"][-1] == voc["n_nodes"] - 1
    irr = vc.random_vocabulary(10, 4, seed=2, irregular=True)
    counts = np.diff(irr["child_start"])
    assert counts.max() <= 10 and set(np.unique(counts)) != {0, 10}                      # fewer than k children somewhere
    assert np.any((irr["node_word"] >= 0) & (irr["node_weight"] == 0.0))                 # stopped words exist
