"""GPU parity of orbx_stereo_match (Frame::ComputeStereoMatches, reference src/Frame.cc:466-640) against the stereo
oracle on the same inputs: the CUDA extractor's own keypoints / descriptors / pyramids of a synthetic rectified pair are
fed to the oracle, so this test isolates the matcher.  Bar: mvuRight and mvDepth bit-exact (float32 bit patterns)."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle import stereo_oracle
from orbslam2_with_quadrics_b200 import ORBextractor, OrbxError
from orbslam2_with_quadrics_b200 import frames as fr

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_stereo_golden", os.path.join(HERE, "golden", "make_stereo_golden.py"))
msg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(msg)
GOLD = json.load(open(os.path.join(HERE, "golden", "stereo_golden.json")))


def oracle_on(exl, fl, kl, dl, exr, frr, kr, dr, mbf, mb):
    return stereo_oracle.compute_stereo_matches(kl, dl, exl.pyramid(fl), kr, dr, exr.pyramid(frr), exl.GetScaleFactors(),
                                                exl.GetInverseScaleFactors(), mbf, mb)


@pytest.mark.parametrize("case", msg.CASES, ids=lambda c: "%s-%d" % (c[0], c[1]))
def test_two_handles_match_oracle_and_golden(case):
    """The reference's pattern (src/Frame.cc:78-81): one extractor per eye, then ComputeStereoMatches."""
    name, seed, mbf, fx = case
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS[name]
    left, right = fr.stereo_pair(w, h, seed)
    mb = float(np.float32(mbf) / np.float32(fx))
    exl, exr = ORBextractor(nf, sf, nl, it, mt), ORBextractor(nf, sf, nl, it, mt)
    kl, dl = exl(left)
    kr, dr = exr(right)
    (u, d), = exl.stereo_match(exr, mbf, mb)
    uo, do, sad = oracle_on(exl, 0, kl, dl, exr, 0, kr, dr, mbf, mb)
    assert len(u) == len(kl)
    assert (u >= 0).sum() > 100
    assert np.array_equal(u.view(np.uint32), uo.view(np.uint32))
    assert np.array_equal(d.view(np.uint32), do.view(np.uint32))
    assert msg.digest(u, d) == GOLD["%s/%d" % (name, seed)]      # extraction + matching end to end vs the committed digests
    exl.close(); exr.close()


def test_one_handle_batch_of_pairs():
    """A batch that holds both eyes of several pairs (even frames left, odd frames right) on ONE handle."""
    name, mbf, fx = "stereo_euroc", 47.90639384423901, 435.2046959714599
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS[name]
    mb = float(np.float32(mbf) / np.float32(fx))
    pairs = [fr.stereo_pair(w, h, 500 + i) for i in range(3)]
    imgs = [im for p in pairs for im in p]
    ex = ORBextractor(nf, sf, nl, it, mt, max_batch=len(imgs))
    res = ex.extract_batch(imgs)
    out = ex.stereo_match(ex, mbf, mb, left_frames=[0, 2, 4], right_frames=[1, 3, 5])
    for i, (u, d) in enumerate(out):
        (kl, dl), (kr, dr) = res[2 * i], res[2 * i + 1]
        uo, do, _ = oracle_on(ex, 2 * i, kl, dl, ex, 2 * i + 1, kr, dr, mbf, mb)
        assert np.array_equal(u.view(np.uint32), uo.view(np.uint32)), i
        assert np.array_equal(d.view(np.uint32), do.view(np.uint32)), i
    # swapped eyes: disparities are negative, the reference refuses them (:615)
    (u, d), = ex.stereo_match(ex, mbf, mb, left_frames=[1], right_frames=[0])
    (kl, dl), (kr, dr) = res[1], res[0]
    uo, do, _ = oracle_on(ex, 1, kl, dl, ex, 0, kr, dr, mbf, mb)
    assert np.array_equal(u.view(np.uint32), uo.view(np.uint32))
    assert np.array_equal(d.view(np.uint32), do.view(np.uint32))
    ex.close()


def test_unrelated_images_and_no_keypoints():
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    mbf, mb = 47.9, 0.11
    ex = ORBextractor(nf, sf, nl, it, mt, max_batch=4)
    res = ex.extract_batch([fr.cluttered_scene(w, h, 11), fr.cluttered_scene(w, h, 12), fr.flat_frame(w, h), fr.flat_frame(w, h)])
    out = ex.stereo_match(ex, mbf, mb, left_frames=[0, 2], right_frames=[1, 3])
    out += ex.stereo_match(ex, mbf, mb, left_frames=[0], right_frames=[3])
    (kl, dl), (kr, dr) = res[0], res[1]
    uo, do, _ = oracle_on(ex, 0, kl, dl, ex, 1, kr, dr, mbf, mb)
    assert np.array_equal(out[0][0].view(np.uint32), uo.view(np.uint32))
    assert np.array_equal(out[0][1].view(np.uint32), do.view(np.uint32))
    assert len(out[1][0]) == 0                                   # no left keypoints
    assert len(out[2][0]) == len(kl) and (out[2][0] == -1).all()  # no right keypoints: nothing matches
    ex.close()


def test_bad_arguments():
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    a, b = ORBextractor(nf, sf, nl, it, mt), ORBextractor(nf, sf, 6, it, mt)
    with pytest.raises(OrbxError):
        a.stereo_match(a, 47.9, 0.11, left_frames=[0], right_frames=[0])      # nothing extracted yet
    a(fr.cluttered_scene(w, h, 1)); b(fr.cluttered_scene(w, h, 2))
    with pytest.raises(OrbxError):
        a.stereo_match(b, 47.9, 0.11)                                         # different pyramid depth
    with pytest.raises(OrbxError):
        a.stereo_match(a, 47.9, 0.0)                                          # mb must be positive (maxD = mbf / mb)
    with pytest.raises(OrbxError):
        a.stereo_match(a, 47.9, 0.11, left_frames=[1], right_frames=[0])      # frame index beyond the last batch
    with pytest.raises(OrbxError):
        a.stereo_match(a, 47.9, 0.11, left_frames=[0, 0], right_frames=[0, 0])   # a left frame may appear once per call
    a.close(); b.close()
