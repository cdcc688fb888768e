"""CPU tests of SURVEY.md §8(f) row 3, Frame::UndistortKeyPoints + AssignFeaturesToGrid (reference src/Frame.cc:404-434,
:230-245, :382-392, :436-464): the restatement (real cv2.undistortPoints + float32 grid arithmetic) against the
reference's own lines compiled against a stub whose undistortPoints is the published 5-iteration algorithm in double."""
import numpy as np
import pytest

from oracle import frame_oracle, orb_oracle, stereo_oracle
from orbslam2_with_quadrics_b200 import frames as fr

# Examples/Monocular/TUM1.yaml:9-17, TUM2.yaml, and an undistorted (rectified) camera
CAMERAS = {"tum1": ((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
           "tum2": ((520.908620, 521.007327, 325.141442, 249.701764), (0.231222, -0.784899, -0.003257, -0.000105, 0.917205)),
           "four": ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
           "rectified": ((435.2046959714599, 435.2046959714599, 367.4517211914062, 252.2008514404297), (0.0, 0.0, 0.0, 0.0, 0.0))}


@pytest.fixture(scope="module")
def ref():
    try:
        stereo_oracle.ref_build()
    except Exception:
        pass
    if not stereo_oracle.ref_available():
        pytest.skip("oracle/_ref/libstereoref.so is not built and /root/reference is absent")
    return frame_oracle


@pytest.mark.parametrize("cam", list(CAMERAS))
def test_restatement_matches_reference_lines(cam, ref):
    K4, D = CAMERAS[cam]
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum" if cam != "four" else "stereo_euroc"]
    res = orb_oracle.ORBextractor(nf, sf, nl, it, mt)(fr.cluttered_scene(w, h, 321))
    xy, start, items, b = frame_oracle.undistort_and_grid(res.keypoints, K4, D, w, h)
    xy2, start2, items2, b2 = ref.ref_undistort_grid(res.keypoints, K4, D, w, h)
    assert np.array_equal(b.view(np.uint32), b2.view(np.uint32))
    assert np.array_equal(xy.view(np.uint32), xy2.view(np.uint32))
    assert np.array_equal(start, start2) and np.array_equal(items, items2)
    assert start[-1] == len(items) <= res.n and len(items) > 0.9 * res.n
    if D[0] != 0.0:
        assert float(np.abs(xy - np.stack([res.keypoints["x"], res.keypoints["y"]], 1)).max()) > 0.5    # it really moves points
    else:
        assert np.array_equal(xy[:, 0], res.keypoints["x"]) and start[-1] == res.n


def test_undistort_shim_matches_cv2_on_random_points(ref):
    """The stub's undistortPoints (what the reference lines call) is bit-identical to the real cv2 one."""
    rng = np.random.default_rng(3)
    kp = np.zeros(20000, orb_oracle.KP_DTYPE)
    kp["x"] = (np.round(rng.uniform(0, 640, len(kp)) * 8) / 8).astype(np.float32)
    kp["y"] = (np.round(rng.uniform(0, 480, len(kp)) * 8) / 8).astype(np.float32)
    for cam in ("tum1", "tum2", "four"):
        K4, D = CAMERAS[cam]
        xy, _, _, _ = frame_oracle.undistort_and_grid(kp, K4, D, 640, 480)
        xy2, _, _, _ = ref.ref_undistort_grid(kp, K4, D, 640, 480)
        assert np.array_equal(xy.view(np.uint32), xy2.view(np.uint32)), cam
