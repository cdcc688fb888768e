"""CPU tests of SURVEY.md §8(f) row 4, the descriptor consumer ORBmatcher::SearchByProjection(Frame&, const Frame&, th,
bMono) (reference src/ORBmatcher.cc:1328-1470, with Frame::GetFeaturesInArea src/Frame.cc:327-380, DescriptorDistance and
ComputeThreeMaxima): the restatement (real cv2.gemm + float32 numpy) against the reference's own lines compiled against a
stub, plus the stub's two gemm formulas against the real cv2.gemm."""
import numpy as np
import pytest

from oracle import frame_oracle, match_oracle, orb_oracle, stereo_oracle
from orbslam2_with_quadrics_b200 import frames as fr
from orbslam2_with_quadrics_b200 import match_cases as mc

K_TUM1 = (517.306408, 516.469215, 318.643040, 255.313989)
D_TUM1 = (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)


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
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    ex = orb_oracle.ORBextractor(nf, sf, nl, it, mt)
    res = ex(fr.cluttered_scene(w, h, 77))
    xy, start, items, b = frame_oracle.undistort_and_grid(res.keypoints, K_TUM1, D_TUM1, w, h)
    return dict(xy_un=xy, cur_octave=res.keypoints["octave"].astype(np.int32), cur_angle=res.keypoints["angle"].astype(np.float32),
                desc=res.descriptors, cell_start=start, cell_items=items, bounds=b, sf=np.asarray(ex.GetScaleFactors(), np.float32),
                nlevels=nl)


def scenario(cf, seed, n_last, motion, stereo):
    rng = np.random.default_rng(seed)
    Tc = mc.pose(rng)
    tz = {"still": 0.0, "forward": 0.6, "backward": -0.6}[motion]
    Tl = mc.pose(rng, t=(0.0, 0.0, tz))
    last = mc.make_last_frame(rng, cf["xy_un"], cf["cur_octave"], cf["cur_angle"], cf["desc"], K_TUM1, Tc, n_last, cf["nlevels"])
    u_right = None
    if stereo:
        u_right = np.where(rng.random(len(cf["desc"])) < 0.6, cf["xy_un"][:, 0] - rng.uniform(2, 40, len(cf["desc"])), -1).astype(np.float32)
    return dict(**last, Tcw_cur=Tc, Tcw_last=Tl, xy_un=cf["xy_un"], cur_octave=cf["cur_octave"], cur_angle=cf["cur_angle"],
                desc=cf["desc"], u_right=u_right, cell_start=cf["cell_start"], cell_items=cf["cell_items"], bounds=cf["bounds"],
                K4=K_TUM1, mbf=40.0, mb=0.08, sf=cf["sf"])


CASES = [(1, 900, "still", False, 15.0, True), (2, 1500, "forward", True, 7.0, False), (3, 1500, "backward", True, 7.0, False),
         (4, 1200, "still", True, 14.0, False), (5, 600, "forward", False, 30.0, True), (6, 40, "still", False, 15.0, True)]


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
