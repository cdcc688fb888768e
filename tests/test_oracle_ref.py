"""The numpy/cv2 oracle (oracle/orb_oracle.py) against the reference's OWN code: the unmodified
/root/reference/src/ORBextractor.cc compiled on the OpenCV shim (oracle/_ref).  This is what pins
the oracle (the reference ships no golden vectors)."""
import numpy as np
import pytest

from oracle import orb_oracle
from orbslam2_with_quadrics_b200 import frames as fr


def _same(ro, kp, ds, pyr):
    assert ro.n == len(kp)
    for f in ro.keypoints.dtype.names:
        assert np.array_equal(ro.keypoints[f], kp[f]), f
    assert np.array_equal(ro.descriptors, ds)
    for a, b in zip(pyr, ro.pyramid):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["mono_tum", "stereo_euroc", "stereo_kitti", "rgbd_1080p", "mono_4k"])
@pytest.mark.parametrize("seed", [1234, 2234])
def test_full_path_matches_reference_tu(ref_available, name, seed):
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS[name]
    if name in ("rgbd_1080p", "mono_4k") and seed != 1234:
        pytest.skip("one 1080p / 4K seed keeps the CPU suite short")
    img = fr.cluttered_scene(w, h, seed)
    ro = orb_oracle.ORBextractor(nf, sf, nl, it, mt)(img)
    r = ref_available.RefORBextractor(nf, sf, nl, it, mt)
    kp, ds = r(img)
    _same(ro, kp, ds, r.pyramid())
    assert abs(ro.n - nf) <= 2 * nl + 8          # quota overshoot is small (App. B-6)


@pytest.mark.parametrize("kind", ["noise", "checker", "flat"])
def test_adversarial_frames_match_reference_tu(ref_available, kind):
    img = {"noise": fr.noise_frame(640, 480, 5), "checker": fr.checker_frame(640, 480, 5), "flat": fr.flat_frame(640, 480)}[kind]
    ro = orb_oracle.ORBextractor(1000, 1.2, 8, 20, 7)(img)
    r = ref_available.RefORBextractor(1000, 1.2, 8, 20, 7)
    kp, ds = r(img)
    _same(ro, kp, ds, r.pyramid())
    if kind == "flat":
        assert ro.n == 0 and ro.stats["retries"] == ro.stats["windows"]     # zero-keypoint path (:1064-1065)


def test_strided_input_and_odd_sizes(ref_available):
    big = fr.cluttered_scene(700, 500, 77)
    view = big[10:10 + 333, 20:20 + 517]                      # ROI with a row stride (treated as isolated, App. A-5)
    ro = orb_oracle.ORBextractor(500, 1.2, 6, 20, 7)(np.ascontiguousarray(view))
    r = ref_available.RefORBextractor(500, 1.2, 6, 20, 7)
    kp, ds = r(view)
    _same(ro, kp, ds, r.pyramid())


def test_empty_image_leaves_outputs_untouched(ref_available):
    assert orb_oracle.ORBextractor(1000, 1.2, 8, 20, 7)(np.zeros((0, 0), np.uint8)) is None
    assert ref_available.RefORBextractor(1000, 1.2, 8, 20, 7)(np.zeros((0, 0), np.uint8)) is None


@pytest.mark.parametrize("args", [(1000, 1.2, 8), (1200, 1.2, 8), (2000, 1.2, 8), (4000, 1.2, 10), (500, 1.5, 4), (300, 1.1, 12)])
def test_constructor_tables_match_reference_tu(ref_available, args):
    nf, sf, nl = args
    p = orb_oracle.OrbParams(nf, sf, nl, 20, 7)
    a, b, c, d, q, u = ref_available.RefORBextractor(nf, sf, nl, 20, 7).tables()
    assert np.array_equal(a.view(np.uint32), p.mvScaleFactor.view(np.uint32))
    assert np.array_equal(b.view(np.uint32), p.mvInvScaleFactor.view(np.uint32))
    assert np.array_equal(c.view(np.uint32), p.mvLevelSigma2.view(np.uint32))
    assert np.array_equal(d.view(np.uint32), p.mvInvLevelSigma2.view(np.uint32))
    assert list(q) == p.mnFeaturesPerLevel
    assert list(u) == p.umax == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def test_quota_known_answers():
    """SURVEY.md §8(a1)."""
    assert orb_oracle.OrbParams(1000, 1.2, 8, 20, 7).mnFeaturesPerLevel == [217, 181, 151, 126, 105, 87, 73, 60]
    assert orb_oracle.OrbParams(1200, 1.2, 8, 20, 7).mnFeaturesPerLevel == [261, 217, 181, 151, 126, 105, 87, 72]
    assert orb_oracle.OrbParams(2000, 1.2, 8, 20, 7).mnFeaturesPerLevel == [434, 362, 302, 251, 209, 175, 145, 122]
    assert orb_oracle.OrbParams(4000, 1.2, 10, 20, 7).mnFeaturesPerLevel == [795, 663, 552, 460, 383, 320, 266, 222, 185, 154]


def _random_candidates(rng, W, H, M, clustered):
    if clustered:
        cx = rng.integers(3, W - 3, 12)
        cy = rng.integers(3, H - 3, 12)
        k = rng.integers(0, 12, M)
        xs = np.clip(cx[k] + rng.integers(-9, 10, M), 3, W - 4)
        ys = np.clip(cy[k] + rng.integers(-9, 10, M), 3, H - 4)
    else:
        xs = rng.integers(3, W - 3, M)
        ys = rng.integers(3, H - 3, M)
    key = np.unique(ys.astype(np.int64) * 100000 + xs)             # a pixel holds at most one keypoint
    rng.shuffle(key)
    xs, ys = (key % 100000).astype(np.int32), (key // 100000).astype(np.int32)
    rs = rng.integers(7, 40, len(xs)).astype(np.int32)             # few distinct responses -> many ties
    return xs, ys, rs


@pytest.mark.parametrize("case", range(24))
def test_octree_restatement_matches_reference_code(ref_available, case):
    """DistributeOctTree (src/ORBextractor.cc:539-763) fuzz: restatement vs the reference's own
    function under canonical rule B-1, including 1..4 root nodes, N larger than the candidate
    count, premature termination (App. B-5) and heavy ties."""
    rng = np.random.default_rng(100 + case)
    W, H = [(608, 448), (1209, 344), (400, 410), (1888, 1048), (330, 90), (72, 200)][case % 6]
    M = [0, 1, 2, 5, 60, 400, 3000, 9000][case % 8]
    N = [1, 7, 60, 217, 434, 795][(case // 2) % 6]
    xs, ys, rs = _random_candidates(rng, W, H, M, clustered=case % 3 == 0)
    if W / H < 0.5:
        pytest.skip("nIni = 0: division by zero in the reference (App. B-7)")
    r = ref_available.RefORBextractor(1000, 1.2, 8, 20, 7)
    got = orb_oracle.distribute_octree(xs, ys, rs, 16, 16 + W, 16, 16 + H, N) if len(xs) else np.zeros(0, np.int64)
    want = r.distribute(xs, ys, rs, 16, 16 + W, 16, 16 + H, N) if len(xs) else np.zeros(0, np.int64)
    assert np.array_equal(got, want)
