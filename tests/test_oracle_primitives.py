"""Primitive known-answer tests: the C restatements in oracle/shim/cvshim.cpp (which the
unmodified reference TU runs on) against the real OpenCV 4.13 wheel."""
import ctypes as C

import cv2
import numpy as np
import pytest

from orbslam2_with_quadrics_b200 import frames as fr

cv2.setNumThreads(1)


def images():
    rng = np.random.default_rng(7)
    out = [("noise", rng.integers(0, 256, (97, 131), dtype=np.uint8)),
           ("scene", fr.cluttered_scene(320, 200, 11)),
           ("checker", fr.checker_frame(200, 150, 3)),
           ("flat", fr.flat_frame(80, 70)),
           ("ramp", (np.add.outer(np.arange(90) * 2, np.arange(120)) % 256).astype(np.uint8)),
           ("extremes", (rng.integers(0, 2, (64, 75)) * 255).astype(np.uint8))]
    return out


@pytest.mark.parametrize("name,img", images())
def test_resize_matches_cv2(ref_available, name, img):
    L = ref_available.lib()
    h, w = img.shape
    for dw, dh in [(int(round(w / 1.2)), int(round(h / 1.2))), (w - 1, h - 1), (w // 2 + 3, h // 2 + 1), (w, h)]:
        dst = np.zeros((dh, dw), np.uint8)
        L.cvshim_resize(img.ctypes.data, w, h, img.strides[0], dst.ctypes.data, dw, dh)
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(dst, ref), (name, dw, dh, int((dst != ref).sum()))


@pytest.mark.parametrize("w,h", [(640, 480), (752, 480), (1241, 376), (1920, 1080)])
def test_resize_pyramid_chain_matches_cv2(ref_available, w, h):
    """The exact ratios ComputePyramid produces (src/ORBextractor.cc:1111-1120)."""
    from oracle.orb_oracle import OrbParams
    L = ref_available.lib()
    p = OrbParams(1000, 1.2, 8, 20, 7)
    img = fr.cluttered_scene(w, h, 5)
    prev = img
    for lw, lh in p.level_sizes(w, h)[1:]:
        dst = np.zeros((lh, lw), np.uint8)
        L.cvshim_resize(prev.ctypes.data, prev.shape[1], prev.shape[0], prev.strides[0], dst.ctypes.data, lw, lh)
        ref = cv2.resize(prev, (lw, lh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(dst, ref)
        prev = ref


@pytest.mark.parametrize("name,img", images())
def test_border_matches_cv2(ref_available, name, img):
    L = ref_available.lib()
    h, w = img.shape
    dst = np.zeros((h + 38, w + 38), np.uint8)
    L.cvshim_border(img.ctypes.data, w, h, img.strides[0], dst.ctypes.data, 19)
    assert np.array_equal(dst, cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))


@pytest.mark.parametrize("name,img", images())
def test_blur_matches_cv2(ref_available, name, img):
    L = ref_available.lib()
    h, w = img.shape
    dst = np.zeros((h, w), np.uint8)
    L.cvshim_blur(img.ctypes.data, w, h, img.strides[0], dst.ctypes.data)
    assert np.array_equal(dst, cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))


def _cv_fast(img, t):
    det = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    return [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(img, None)]


@pytest.mark.parametrize("name,img", images())
@pytest.mark.parametrize("t", [20, 7])
def test_fast_matches_cv2(ref_available, name, img, t):
    L = ref_available.lib()
    h, w = img.shape
    out = np.zeros((w * h, 3), np.int32)
    n = L.cvshim_fast(img.ctypes.data, w, h, img.strides[0], t, len(out), out.ctypes.data)
    assert [tuple(r) for r in out[:n].tolist()] == _cv_fast(img, t)


def test_fast_small_windows(ref_available):
    """Clipped last-row windows can be 4..6 px tall (src/ORBextractor.cc:794): FAST must return nothing below 7."""
    L = ref_available.lib()
    rng = np.random.default_rng(3)
    for h in range(3, 10):
        for w in (5, 6, 7, 37):
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
            out = np.zeros((w * h + 1, 3), np.int32)
            n = L.cvshim_fast(img.ctypes.data, w, h, img.strides[0], 7, len(out), out.ctypes.data)
            assert [tuple(r) for r in out[:n].tolist()] == _cv_fast(img, 7)


def test_fast_on_subview_equals_copy(ref_available):
    """cv::FAST on a cell window sees no pixel outside the window (App. A-3)."""
    L = ref_available.lib()
    img = fr.cluttered_scene(200, 160, 9)
    win = img[30:68, 41:78]
    out = np.zeros((win.size, 3), np.int32)
    n = L.cvshim_fast(win.ctypes.data, win.shape[1], win.shape[0], win.strides[0], 20, len(out), out.ctypes.data)
    assert [tuple(r) for r in out[:n].tolist()] == _cv_fast(np.ascontiguousarray(win), 20)


def test_fast_atan2_matches_cv2(ref_available):
    L = ref_available.lib()
    rng = np.random.default_rng(1)
    ys = np.concatenate([rng.integers(-60000, 60000, 5000), [0, 0, 1, -1, 0, 5, -5]])
    xs = np.concatenate([rng.integers(-60000, 60000, 5000), [0, 1, 0, 0, -1, 5, -5]])
    for y, x in zip(ys, xs):
        a = np.float32(L.cvshim_fast_atan2(float(y), float(x)))
        b = np.float32(cv2.fastAtan2(float(y), float(x)))
        assert a.view(np.uint32) == b.view(np.uint32), (y, x, a, b)


def test_cvtcolor_gray_closed_form_matches_cv2():
    """SURVEY.md §8(f) row 2 (reference src/Tracking.cc:172-255): the arithmetic the CUDA conversion kernel restates,
    gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15, against the real cv2.cvtColor for all four colour orders."""
    import cv2
    rng = np.random.default_rng(5)
    for ch, codes in ((3, (cv2.COLOR_BGR2GRAY, cv2.COLOR_RGB2GRAY)), (4, (cv2.COLOR_BGRA2GRAY, cv2.COLOR_RGBA2GRAY))):
        img = rng.integers(0, 256, (301, 517, ch), dtype=np.uint8)
        c = [img[..., i].astype(np.int64) for i in range(3)]
        for code, (b, g, r) in zip(codes, ((c[0], c[1], c[2]), (c[2], c[1], c[0]))):
            want = cv2.cvtColor(img, code)
            got = ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)
            assert np.array_equal(want, got)
