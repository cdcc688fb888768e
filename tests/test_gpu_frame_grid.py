"""GPU parity of orbx_undistort_grid (Frame::UndistortKeyPoints + AssignFeaturesToGrid, reference src/Frame.cc:404-434,
:230-245) against oracle/frame_oracle.py (real cv2.undistortPoints + the float32 grid arithmetic, itself pinned against
the reference's own lines).  Bar: undistorted coordinates bit-exact, grid cells and their index order identical."""
import numpy as np
import pytest

from oracle import frame_oracle
from orbslam2_with_quadrics_b200 import ORBextractor, OrbxError
from orbslam2_with_quadrics_b200 import frames as fr

pytestmark = pytest.mark.gpu
CAMERAS = {"tum1": ((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
           "tum2": ((520.908620, 521.007327, 325.141442, 249.701764), (0.231222, -0.784899, -0.003257, -0.000105, 0.917205)),
           "four": ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
           "rectified": ((435.2046959714599, 435.2046959714599, 367.4517211914062, 252.2008514404297), (0.0, 0.0, 0.0, 0.0, 0.0))}


def same(got, want):
    xy, st, it, b = got
    xo, so, io, bo = want
    assert np.array_equal(b.view(np.uint32), bo.view(np.uint32))
    assert np.array_equal(xy.view(np.uint32), xo.view(np.uint32))
    assert np.array_equal(st, so) and np.array_equal(it, io)


@pytest.mark.parametrize("cam", list(CAMERAS))
def test_undistort_and_grid_match_oracle(cam):
    K4, D = CAMERAS[cam]
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum" if cam != "four" else "stereo_euroc"]
    imgs = [fr.cluttered_scene(w, h, 600 + i) for i in range(3)] + [fr.flat_frame(w, h)]
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=4)
    res = gx.extract_batch(imgs)
    out = gx.undistort_grid(K4, D)
    assert len(out) == 4
    for (kps, _), got in zip(res, out):
        same(got, frame_oracle.undistort_and_grid(kps, K4, D, w, h))
    assert len(out[3][0]) == 0 and out[3][1][-1] == 0                   # no keypoints: empty grid
    # a subset of frames, in another order
    out2 = gx.undistort_grid(K4, D, frames=[2, 0])
    same(out2[0], frame_oracle.undistort_and_grid(res[2][0], K4, D, w, h))
    same(out2[1], frame_oracle.undistort_and_grid(res[0][0], K4, D, w, h))
    gx.close()


def test_undistort_grid_4k_and_bad_arguments():
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_4k"]
    K4, D = (3100.0, 3098.0, 1915.5, 1082.25), (0.12, -0.31, 0.0007, -0.0004, 0.09)
    gx = ORBextractor(nf, sf, nl, it, mt)
    with pytest.raises(OrbxError):
        gx.undistort_grid(K4, D, frames=[0])                            # nothing extracted yet
    kps, _ = gx(fr.cluttered_scene(w, h, 9))
    (got,) = gx.undistort_grid(K4, D)
    same(got, frame_oracle.undistort_and_grid(kps, K4, D, w, h))
    with pytest.raises(OrbxError):
        gx.undistort_grid(K4, D[:3])                                    # 4 or 5 distortion coefficients
    with pytest.raises(OrbxError):
        gx.undistort_grid(K4, D, frames=[0, 0])
    gx.close()
