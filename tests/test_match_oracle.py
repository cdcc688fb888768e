"""CPU tests of SURVEY.md §8(f) row 4, the descriptor consumer ORBmatcher::SearchByProjection(Frame&, const Frame&, th,
bMono) (reference src/ORBmatcher.cc:1328-1470, with Frame::GetFeaturesInArea src/Frame.cc:327-380, DescriptorDistance and
ComputeThreeMaxima): the restatement (real cv2.gemm + float32 numpy) against the reference's own lines compiled against a
stub, plus the stub's two gemm formulas against the real cv2.gemm."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle import match_oracle, stereo_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_match_golden", os.path.join(HERE, "golden", "make_match_golden.py"))
mmg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mmg)
GOLD = json.load(open(os.path.join(HERE, "golden", "match_golden.json")))
CASES, scenario = mmg.CASES, mmg.scenario


@pytest.fixture(scope="module")
def ref():
    try:
        stereo_oracle.ref_build()
    except Exception:
        pass
    if not match_oracle.ref_available():
        pytest.skip("oracle/_ref/libstereoref.so is not built and /root/reference is absent")
    return match_oracle


@pytest.fixture(scope="module")
def current_frame():
    return mmg.current_frame()


@pytest.mark.parametrize("case", CASES)
def test_restatement_matches_golden_of_the_reference_lines(case, current_frame):
    """Needs neither /root/reference nor oracle/_ref: the digests were made by the reference's own lines."""
    sc = scenario(current_frame, *case[:4])
    for check in (True, False):
        n, m = match_oracle.search_by_projection(th=case[4], mono=case[5], check_orientation=check, **sc)
        assert mmg.digest(sc, n, m) == GOLD[mmg.key(case, check)]


@pytest.mark.parametrize("seed,n_last,motion,stereo,th,mono", CASES)
def test_restatement_matches_reference_lines(seed, n_last, motion, stereo, th, mono, ref, current_frame):
    sc = scenario(current_frame, seed, n_last, motion, stereo)
    for check in (True, False):
        n1, m1 = match_oracle.search_by_projection(th=th, mono=mono, check_orientation=check, **sc)
        n2, m2 = ref.ref_search_by_projection(th=th, mono=mono, check_orientation=check, **sc)
        assert n1 == n2 and np.array_equal(m1, m2)
    fwd, bwd = match_oracle.motion_flags(sc["Tcw_cur"], sc["Tcw_last"], sc["mb"], mono)
    assert (fwd, bwd) == (motion == "forward" and not mono, motion == "backward" and not mono)
    assert n1 > 0.2 * n_last                               # the scenario really matches
    holders = m1[m1 >= 0]
    assert len(holders) <= n1                               # overwritten matches stay counted (the reference's nmatches)


def test_contention_is_exercised(current_frame):
    """The scenarios contain what makes the function sequential: a keypoint claimed by a map point with observations is
    skipped by later points, one claimed by a point without observations is overwritten."""
    sc = scenario(current_frame, 1, 900, "still", False)
    n, m = match_oracle.search_by_projection(th=15.0, mono=True, check_orientation=False, **sc)
    assert n > len(m[m >= 0])                               # at least one overwrite happened


def test_shim_gemm_formulas_match_cv2():
    """The two products the stub evaluates (oracle/shim_stereo/stereo_shim.h) are cv2.gemm bit for bit."""
    import cv2
    rng = np.random.default_rng(5)
    f32 = np.float32
    for _ in range(3000):
        R = rng.standard_normal((3, 3)).astype(f32)
        x = (rng.standard_normal((3, 1)) * 5).astype(f32)
        t = rng.standard_normal((3, 1)).astype(f32)
        y = cv2.gemm(R, x, 1, t, 1)
        a = [f32(np.float64(f32(f32(f32(R[r, 0] * x[0, 0]) + f32(R[r, 1] * x[1, 0])) + f32(R[r, 2] * x[2, 0]))) + np.float64(t[r, 0]))
             for r in range(3)]
        assert np.array_equal(np.array(a, f32).reshape(3, 1), y)
        y = cv2.gemm(R, t, -1, None, 0, flags=cv2.GEMM_1_T)
        b = [f32(-(np.float64(R[0, r]) * np.float64(t[0, 0]) + np.float64(R[1, r]) * np.float64(t[1, 0])
                   + np.float64(R[2, r]) * np.float64(t[2, 0]))) for r in range(3)]
        assert np.array_equal(np.array(b, f32).reshape(3, 1), y)


# ----------------------------------------------------------------------------- SearchLocalPoints' matcher (:45-129)
@pytest.mark.parametrize("case", mmg.LOCAL_CASES)
def test_local_points_restatement_matches_reference_lines_and_golden(case, current_frame):
    sc = mmg.local_scenario(current_frame, *case[:3])
    n1, m1 = match_oracle.search_local_points(th=case[3], **sc)
    assert mmg.local_digest(sc, n1, m1) == GOLD[mmg.local_key(case)]
    if match_oracle.ref_available():
        n2, m2 = match_oracle.ref_search_local_points(th=case[3], **sc)
        assert n1 == n2 and np.array_equal(m1, m2)
    assert n1 >= int((m1 >= 0).sum()) > 0.2 * case[1]


def test_local_points_ratio_test_and_held_points_are_exercised(current_frame):
    sc = mmg.local_scenario(current_frame, 2, 2000, True)
    n, m = match_oracle.search_local_points(th=3.0, **sc)
    n_no_ratio, _ = match_oracle.search_local_points(th=3.0, nnratio=0.0, **sc)      # best > 0 * best2 rejects every same-level pair
    n_never, _ = match_oracle.search_local_points(th=3.0, nnratio=1e9, **sc)        # never rejects
    assert n_no_ratio < n < n_never
    assert not np.any((m >= 0) & (sc["cur_obs"] > 0))           # keypoints holding an observed point are never taken (:96-98)
    assert np.any((m >= 0) & (sc["cur_obs"] == 0))              # ... those holding an unobserved one are
    assert match_oracle.radius_by_viewing_cos(np.float32(0.998)) == 2.5 and match_oracle.radius_by_viewing_cos(0.99799) == 4.0


# ----------------------------------------------------------------------------- SearchByBoW(KeyFrame*, Frame&) (:159-288)
@pytest.mark.parametrize("case", mmg.BOW_CASES)
def test_search_by_bow_restatement_matches_reference_lines_and_golden(case, current_frame):
    _, sc = mmg.bow_scenario(current_frame, case)
    n1, m1 = match_oracle.search_by_bow(nnratio=case[5], check_orientation=case[6], **sc)
    assert mmg.bow_digest(sc, n1, m1) == GOLD[mmg.bow_key(case)]
    if match_oracle.ref_available():
        n2, m2 = match_oracle.ref_search_by_bow(nnratio=case[5], check_orientation=case[6], **sc)
        assert n1 == n2 and np.array_equal(m1, m2)
    assert n1 == int((m1 >= 0).sum()) > 0.15 * case[4]          # every counted match holds its keypoint (claims are exclusive)
    kf_of = m1[m1 >= 0]
    assert np.all(sc["kf_valid"][kf_of] == 1)                   # features without a good map point never match (:196-202)


# ----------------------------------------------------------------------------- Relocalization's SearchByProjection(Frame&, KeyFrame*) (:1472-1599)
@pytest.mark.parametrize("case", mmg.KF_CASES)
def test_keyframe_projection_restatement_matches_reference_lines_and_golden(case, current_frame):
    sc = mmg.kf_scenario(current_frame, *case[:2])
    lsf = mmg.kf_log_scale(current_frame)
    inr, pred = match_oracle.kf_prepare(sc["valid"], sc["world"], sc["min_dist"], sc["max_dist"], sc["Tcw_cur"], lsf, current_frame["nlevels"])
    args = {k: v for k, v in sc.items() if k not in ("min_dist", "max_dist")}
    n1, m1 = match_oracle.search_by_projection_kf(in_range=inr, pred_level=pred, th=case[2], orb_dist=case[3], check_orientation=case[4], **args)
    assert mmg.kf_digest(sc, n1, m1, inr, pred) == GOLD[mmg.kf_key(case)]
    if match_oracle.ref_has("matchref_search_by_projection_kf"):
        n2, m2, inr2, pred2 = match_oracle.ref_search_by_projection_kf(log_scale_factor=lsf, th=case[2], orb_dist=case[3],
                                                                       check_orientation=case[4], **sc)
        assert n1 == n2 and np.array_equal(m1, m2)
        assert np.array_equal(inr, inr2) and np.array_equal(pred[inr2 > 0], pred2[inr2 > 0])
    assert n1 == int((m1 >= 0).sum()) > 0.1 * case[1]           # every keypoint is taken at most once (any holder blocks, :1540-1541)
    assert not np.any((m1 >= 0) & (sc["cur_held"] > 0))         # keypoints the frame already holds are never taken
    assert np.all(sc["valid"][m1[m1 >= 0]] == 1)                # bad / already found / missing points never match (:1493-1495)
    assert 0 < int(inr.sum()) < int((sc["valid"] > 0).sum())    # the distance-invariance test rejects some points
    assert len(set(pred[inr > 0].tolist())) > 3                 # several predicted levels occur


def test_shim_norm_matches_cv2():
    """cv::norm of a 3 x 1 float vector as the stub evaluates it (sum of squares in double, sqrt) is cv2.norm bit for bit."""
    import cv2
    rng = np.random.default_rng(9)
    for _ in range(3000):
        v = (rng.standard_normal((3, 1)) * rng.choice([0.01, 1.0, 30.0])).astype(np.float32)
        s = np.float64(0)
        for k in range(3):
            s += np.float64(v[k, 0]) * np.float64(v[k, 0])
        assert np.sqrt(s) == cv2.norm(v)


# ----------------------------------------------------------------------------- SearchForInitialization (:405-520)
@pytest.mark.parametrize("case", mmg.INIT_CASES)
def test_search_for_initialization_restatement_matches_reference_lines_and_golden(case, current_frame):
    sc = mmg.init_scenario(current_frame, *case[:2])
    n1, m1, p1 = match_oracle.search_for_initialization(nnratio=case[2], check_orientation=case[3], window=case[4], **sc)
    assert mmg.init_digest(sc, n1, m1, p1) == GOLD[mmg.init_key(case)]
    if match_oracle.ref_has("matchref_search_for_initialization"):
        n2, m2, p2 = match_oracle.ref_search_for_initialization(sf=current_frame["sf"], nnratio=case[2], check_orientation=case[3],
                                                                window=case[4], **sc)
        assert n1 == n2 and np.array_equal(m1, m2) and np.array_equal(p1, p2)
    assert n1 == int((m1 >= 0).sum()) > 0.1 * case[1]
    assert np.all(sc["octave1"][m1 >= 0] == 0)                  # only level-0 keypoints of F1 are matched (:424-426)
    taken = m1[m1 >= 0]
    assert len(set(taken.tolist())) == len(taken)               # a second-frame keypoint has one owner (:471-475)


def test_search_for_initialization_stealing_is_exercised(current_frame):
    """A later F1 keypoint takes an F2 keypoint from an earlier one when it is strictly closer (:447-448, :471-478)."""
    sc = mmg.init_scenario(current_frame, 1, 1000)
    n, m, _ = match_oracle.search_for_initialization(nnratio=0.9, check_orientation=False, window=100, **sc)
    # replay without the stealing rule: more F1 keypoints would keep a match
    import oracle.match_oracle as mo
    d1 = np.ascontiguousarray(sc["desc1"]).view(np.uint64).reshape(-1, 4)
    d2 = np.ascontiguousarray(sc["desc2"]).view(np.uint64).reshape(-1, 4)
    owners = {}
    for i1 in np.nonzero(m >= 0)[0]:
        owners[int(m[i1])] = int(i1)
    stolen = 0
    for i1 in range(len(m)):
        if m[i1] >= 0 or sc["octave1"][i1] > 0:
            continue
        cands = mo.features_in_area(sc["prev_matched"][i1, 0], sc["prev_matched"][i1, 1], np.float32(100), 0, 0, sc["xy_un2"],
                                    sc["octave2"], sc["cell_start"], sc["cell_items"], sc["bounds"])
        ds = [(int(sum(bin(int(a ^ b)).count("1") for a, b in zip(d1[i1], d2[i2]))), i2) for i2 in cands]
        if ds and min(ds)[0] <= mo.TH_LOW and min(ds)[1] in owners and owners[min(ds)[1]] > i1:
            stolen += 1
    assert stolen > 0


# ----------------------------------------------------------------------------- Frame::isInFrustum (src/Frame.cc:269-325)
@pytest.mark.parametrize("case", mmg.FRUSTUM_CASES)
def test_is_in_frustum_restatement_matches_reference_lines_and_golden(case, current_frame):
    sc = mmg.frustum_scenario(current_frame, *case[:2])
    r1 = match_oracle.is_in_frustum(cos_limit=case[2], mbf=case[3], **sc)
    assert mmg.frustum_digest(sc, *r1) == GOLD[mmg.frustum_key(case)]
    if match_oracle.ref_has("matchref_is_in_frustum"):
        r2 = match_oracle.ref_is_in_frustum(cos_limit=case[2], mbf=case[3], **sc)
        for a, b in zip(r1, r2):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8))          # floats by their bits
    in_view, proj, level, vcos = r1
    assert not np.any(in_view[sc["consider"] == 0])                              # skipped points are never marked (src/Tracking.cc:1169-1172)
    if case[1] >= 800:
        assert 0.1 * case[1] < int(in_view.sum()) < 0.6 * case[1]               # every rejection reason removes something
        assert set(level[in_view > 0].tolist()) == set(range(current_frame["nlevels"]))
        assert float(vcos[in_view > 0].min()) < case[2] + 0.05                  # the cosine limit cuts through the scenario
