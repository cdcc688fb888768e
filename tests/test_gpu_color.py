"""GPU parity of the colour entry points (cvtColor of Tracking::GrabImage*, reference
src/Tracking.cc:172-255, moved in front of the path on the device).  Oracle = the real cv2.cvtColor followed by the
extractor oracle; bar: everything bit-exact, including the level-0 plane (= the converted image plus border)."""
import cv2
import numpy as np
import pytest

from oracle import orb_oracle
from orbslam2_with_quadrics_b200 import ORBextractor, OrbxError, _capi
from orbslam2_with_quadrics_b200 import frames as fr

pytestmark = pytest.mark.gpu
CODES = {_capi.BGR8: cv2.COLOR_BGR2GRAY, _capi.RGB8: cv2.COLOR_RGB2GRAY, _capi.BGRA8: cv2.COLOR_BGRA2GRAY, _capi.RGBA8: cv2.COLOR_RGBA2GRAY}


def colour_scene(w, h, ch, seed):
    planes = [fr.cluttered_scene(w, h, seed + 17 * c) for c in range(ch)]
    return np.ascontiguousarray(np.stack(planes, axis=2))


@pytest.mark.parametrize("fmt", [_capi.BGR8, _capi.RGB8, _capi.BGRA8, _capi.RGBA8])
@pytest.mark.parametrize("shape", [(480, 640), (333, 517)])
def test_colour_frames_match_cvtcolor_then_oracle(fmt, shape):
    h, w = shape
    ch = 3 if fmt in (_capi.BGR8, _capi.RGB8) else 4
    img = colour_scene(w, h, ch, 40 + fmt)
    gray = cv2.cvtColor(img, CODES[fmt])
    args = (800, 1.2, 6, 20, 7)
    ro = orb_oracle.ORBextractor(*args)(gray)
    gx = ORBextractor(*args, max_batch=2)
    big = np.zeros((h + 5, w + 9, ch), np.uint8)              # a view with a row stride of its own
    big[2:2 + h, 3:3 + w] = img
    for frames in ([img], [big[2:2 + h, 3:3 + w], img]):
        out = gx.extract_batch_color(frames, fmt)
        for kps, desc in out:
            assert len(kps) == ro.n
            for f in kps.dtype.names:
                assert np.array_equal(kps[f], ro.keypoints[f]), f
            assert np.array_equal(desc, ro.descriptors)
        assert np.array_equal(gx.stage_dump(0, 0, _capi.STAGE_PYRAMID), ro.pyramid[0])
    gx.close()


def test_colour_bad_arguments():
    gx = ORBextractor(500, 1.2, 4, 20, 7)
    img = colour_scene(320, 240, 3, 1)
    with pytest.raises(OrbxError):
        gx.extract_batch_color([img], _capi.GRAY8)            # not a colour format
    with pytest.raises(OrbxError):
        gx.extract_batch_color([img], 9)
    gx.close()
