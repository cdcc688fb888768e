"""Committed golden digests (tests/golden/orb_golden.json): the synthetic frames are stable, the
oracle reproduces its own digests, and -- where oracle/_ref is available -- so does the
reference's own translation unit."""
import hashlib
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle import orb_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)
GOLD = json.load(open(os.path.join(HERE, "golden", "orb_golden.json")))
CASES = {k: (img, args) for k, img, args in mg.cases()}
SMALL = [k for k in CASES if not k.startswith(("mono_4k", "rgbd_1080p/seed2234"))]


@pytest.mark.parametrize("key", sorted(CASES))
def test_synthetic_frames_are_stable(key):
    assert hashlib.sha256(CASES[key][0].tobytes()).hexdigest() == GOLD[key]["image_sha256"]


@pytest.mark.parametrize("key", sorted(SMALL))
def test_oracle_reproduces_golden(key):
    img, args = CASES[key]
    r = orb_oracle.ORBextractor(*args)(img)
    d = mg.digest(r.keypoints, r.descriptors, r.pyramid)
    for f, v in d.items():
        assert GOLD[key][f] == v, f


@pytest.mark.parametrize("key", sorted(SMALL))
def test_reference_tu_reproduces_golden(ref_available, key):
    img, args = CASES[key]
    ref = ref_available.RefORBextractor(*args)
    kp, ds = ref(img)
    d = mg.digest(kp, ds, ref.pyramid())
    for f, v in d.items():
        assert GOLD[key][f] == v, f


def test_known_answers_are_human_checkable():
    g = GOLD["mono_tum/seed1234"]
    assert g["n"] == 1007 and g["first_keypoint"][:3] == [364.0, 113.0, 31.0] and g["first_keypoint"][4:] == [39.0, 0]
