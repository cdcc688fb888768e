"""Deterministic synthetic 8-bit gray frames for parity tests and benchmarks.

SURVEY.md §8(d) asks for a seeded "cluttered scene" generator that does not
depend on cv2's rasteriser: mid-gray canvas, W*H/900 filled rectangles and
ellipses (side 4..60 px, uniform gray), a small integer blur, integer noise
of +-3, clipped to u8.  Everything here is integer numpy arithmetic, so a seed
produces the same bytes on every machine.

Frame i of camera stream s uses ``seed = 1234 + 1000*s + i`` (``stream_seed``).
Three adversarial frames exercise the reference's corner cases
(src/ORBextractor.cc:812 retry, :1064 zero-keypoint path, candidate overflow).
"""
from __future__ import annotations

import numpy as np

# The five BASELINE.json configs: name -> (width, height, nfeatures, scale, nlevels, iniTh, minTh, images per frame)
CONFIGS = {
    "mono_tum": (640, 480, 1000, 1.2, 8, 20, 7, 1),        # Examples/Monocular/TUM1.yaml:30-43
    "stereo_euroc": (752, 480, 1200, 1.2, 8, 20, 7, 2),    # Examples/Stereo/EuRoC.yaml:88-101
    "stereo_kitti": (1241, 376, 2000, 1.2, 8, 20, 7, 2),   # Examples/Stereo/KITTI00-02.yaml:38-51
    "rgbd_1080p": (1920, 1080, 2000, 1.2, 8, 20, 7, 1),
    "mono_4k": (3840, 2160, 4000, 1.2, 10, 20, 7, 1),
}


def stream_seed(stream: int, frame: int) -> int:
    return 1234 + 1000 * stream + frame


def _binomial_blur(img: np.ndarray) -> np.ndarray:
    """5x5 binomial ([1,4,6,4,1]/16 per axis) integer blur, edge-replicated."""
    a = np.pad(img.astype(np.int32), 2, mode="edge")
    h = (a[:, 0:-4] + 4 * a[:, 1:-3] + 6 * a[:, 2:-2] + 4 * a[:, 3:-1] + a[:, 4:])
    v = (h[0:-4] + 4 * h[1:-3] + 6 * h[2:-2] + 4 * h[3:-1] + h[4:])
    return ((v + 128) >> 8).astype(np.uint8)


def cluttered_scene(width: int, height: int, seed: int) -> np.ndarray:
    """Seeded cluttered scene, (height, width) uint8, C-contiguous."""
    rng = np.random.default_rng(seed)
    img = np.full((height, width), 128, np.uint8)
    n = int(width * height / 900)
    xs = rng.integers(0, width, n)
    ys = rng.integers(0, height, n)
    aa = rng.integers(4, 60, n)
    bb = rng.integers(4, 60, n)
    cc = rng.integers(0, 256, n)
    kind = rng.random(n) < 0.5
    # integer "rotation": ellipses get one of 8 orientations via an integer shear-free quadratic form
    rot = rng.integers(0, 8, n)
    cs = np.array([(16, 0), (15, 6), (11, 11), (6, 15), (0, 16), (-6, 15), (-11, 11), (-15, 6)], np.int64)
    for i in range(n):
        x, y, a, b, c = int(xs[i]), int(ys[i]), int(aa[i]), int(bb[i]), int(cc[i])
        if kind[i]:
            img[y:min(y + b + 1, height), x:min(x + a + 1, width)] = c
        else:
            ra, rb = a // 2 + 1, b // 2 + 1
            r = max(ra, rb)
            x0, x1 = max(x - r, 0), min(x + r + 1, width)
            y0, y1 = max(y - r, 0), min(y + r + 1, height)
            if x0 >= x1 or y0 >= y1:
                continue
            dx = np.arange(x0, x1, dtype=np.int64)[None, :] - x
            dy = np.arange(y0, y1, dtype=np.int64)[:, None] - y
            co, si = cs[rot[i]]
            u = co * dx + si * dy          # scaled by 16
            v = -si * dx + co * dy
            inside = (u * u) * (rb * rb) + (v * v) * (ra * ra) <= 256 * (ra * ra) * (rb * rb)
            img[y0:y1, x0:x1][inside] = c
    img = _binomial_blur(img)
    noise = rng.integers(-3, 4, (height, width))
    return np.clip(img.astype(np.int32) + noise, 0, 255).astype(np.uint8)


def flat_frame(width: int, height: int, value: int = 97) -> np.ndarray:
    """All-flat frame: every FAST window is retried, zero keypoints (src/ORBextractor.cc:1064-1065)."""
    return np.full((height, width), value, np.uint8)


def noise_frame(width: int, height: int, seed: int) -> np.ndarray:
    """Uniform noise: ~9% of level-0 pixels become candidates (candidate-buffer stress)."""
    return np.random.default_rng(seed).integers(0, 256, (height, width), dtype=np.uint8)


def checker_frame(width: int, height: int, seed: int, cell: int = 9) -> np.ndarray:
    """4-gray-level checkerboard with many equal FAST scores (NMS ties / minThFAST retry path)."""
    rng = np.random.default_rng(seed)
    levels = np.array([40, 58, 76, 120], np.uint8)   # low contrast in places -> retries at minThFAST
    gy, gx = (height + cell - 1) // cell, (width + cell - 1) // cell
    grid = levels[rng.integers(0, 4, (gy, gx))]
    return np.ascontiguousarray(np.kron(grid, np.ones((cell, cell), np.uint8))[:height, :width])


def stereo_pair(width: int, height: int, seed: int, bands: int = 6, max_disparity: int = 40):
    """Rectified synthetic stereo pair for the ComputeStereoMatches row (src/Frame.cc:466-640): the right image is
    the left scene shifted by an integer disparity that is constant inside each of `bands` horizontal bands
    (right(x) = left(x + d), so a feature at uL appears at uR = uL - d), plus its own +-2 sensor noise."""
    rng = np.random.default_rng(seed + 777)
    margin = max_disparity + 8
    scene = cluttered_scene(width + margin, height, seed)
    left = np.ascontiguousarray(scene[:, :width])
    right = np.empty_like(left)
    edges = np.linspace(0, height, bands + 1).astype(int)
    disp = rng.integers(2, max_disparity + 1, bands)
    for b in range(bands):
        d = int(disp[b])
        right[edges[b]:edges[b + 1]] = scene[edges[b]:edges[b + 1], d:d + width]
    noise = rng.integers(-2, 3, right.shape)
    right = np.clip(right.astype(np.int32) + noise, 0, 255).astype(np.uint8)
    return left, right


def config_frames(name: str, stream: int = 0, frame: int = 0):
    """Images of one frame of a BASELINE config (2 for stereo: right uses stream+1)."""
    w, h, *_rest, nimg = CONFIGS[name]
    return [cluttered_scene(w, h, stream_seed(stream + k, frame)) for k in range(nimg)]
