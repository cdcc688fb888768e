"""CPU restatement of Frame::UndistortKeyPoints + Frame::AssignFeaturesToGrid with their set-up
(reference src/Frame.cc:404-434, :230-245, PosInGrid :382-392, ComputeImageBounds :436-464, grid scale :155-156).

TEST INFRASTRUCTURE: the checker of ``orbx_undistort_grid``; only tests/ may import it.  The undistortion itself is
the REAL OpenCV (cv2.undistortPoints, 4.13.0); everything around it is restated in float32.  Pinned against the
reference's own lines compiled against a stub (oracle/_ref/libstereoref.so, ``ref_undistort_grid`` below) by
tests/test_frame_oracle.py.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

f32 = np.float32
GRID_COLS, GRID_ROWS = 64, 48          # include/Frame.h:39-40
_HERE = os.path.dirname(os.path.abspath(__file__))


def _undistort(pts, K4, D):
    import cv2
    fx, fy, cx, cy = [float(f32(v)) for v in K4]
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    out = cv2.undistortPoints(np.ascontiguousarray(pts, np.float32).reshape(-1, 1, 2), K, np.asarray(D, np.float32), None, K)
    return out.reshape(-1, 2)


def _round_half_away(x) -> int:
    x = float(x)
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def undistort_and_grid(kps, K4, D, width, height):
    """kps: structured keypoints (x, y ...); K4 = (fx, fy, cx, cy); D = (k1, k2, p1, p2[, k3]).
    Returns (xy_un (n, 2) float32, cell_start (64*48+1,), cell_items, bounds (mnMinX, mnMaxX, mnMinY, mnMaxY))."""
    n = len(kps)
    D = np.asarray(D, np.float32)
    xy = np.stack([kps["x"], kps["y"]], axis=1).astype(np.float32) if n else np.zeros((0, 2), np.float32)
    if D[0] != 0.0:                                                   # ComputeImageBounds (:438-455)
        c = _undistort(np.array([[0, 0], [width, 0], [0, height], [width, height]], np.float32), K4, D)
        mnMinX, mnMaxX = min(c[0, 0], c[2, 0]), max(c[1, 0], c[3, 0])
        mnMinY, mnMaxY = min(c[0, 1], c[1, 1]), max(c[2, 1], c[3, 1])
        xy_un = _undistort(xy, K4, D) if n else xy                    # UndistortKeyPoints (:412-433)
    else:
        mnMinX, mnMaxX, mnMinY, mnMaxY = f32(0), f32(width), f32(0), f32(height)
        xy_un = xy.copy()                                             # (:406-410)
    mnMinX, mnMaxX, mnMinY, mnMaxY = f32(mnMinX), f32(mnMaxX), f32(mnMinY), f32(mnMaxY)
    winv = f32(f32(GRID_COLS) / f32(mnMaxX - mnMinX))                 # (:155-156)
    hinv = f32(f32(GRID_ROWS) / f32(mnMaxY - mnMinY))
    cells = [[] for _ in range(GRID_COLS * GRID_ROWS)]
    for i in range(n):                                                # AssignFeaturesToGrid (:237-244), PosInGrid (:384-391)
        px = _round_half_away(f32(f32(xy_un[i, 0] - mnMinX) * winv))
        py = _round_half_away(f32(f32(xy_un[i, 1] - mnMinY) * hinv))
        if 0 <= px < GRID_COLS and 0 <= py < GRID_ROWS:
            cells[px * GRID_ROWS + py].append(i)
    start = np.zeros(GRID_COLS * GRID_ROWS + 1, np.int32)
    start[1:] = np.cumsum([len(c) for c in cells])
    items = np.array([i for c in cells for i in c], np.int32)
    return xy_un.astype(np.float32), start, items, np.array([mnMinX, mnMaxX, mnMinY, mnMaxY], np.float32)


_ref = None


def ref_undistort_grid(kps, K4, D, width, height):
    """The same through the reference's own lines (oracle/_ref/libstereoref.so)."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libstereoref.so"))
        _ref.frameref_undistort_grid.restype = C.c_int
    n = len(kps)
    k = np.ascontiguousarray(kps).view(np.float32).reshape(-1, 7) if n else np.zeros((0, 7), f32)
    K = np.ascontiguousarray(K4, f32)
    Dv = np.ascontiguousarray(D, f32)
    xy = np.zeros((max(n, 1), 2), f32)
    start = np.zeros(GRID_COLS * GRID_ROWS + 1, np.int32)
    items = np.zeros(max(n, 1), np.int32)
    b = np.zeros(4, f32)
    m = _ref.frameref_undistort_grid(C.c_int(n), C.c_void_p(k.ctypes.data), C.c_void_p(K.ctypes.data), C.c_void_p(Dv.ctypes.data),
                                     C.c_int(len(Dv)), C.c_int(width), C.c_int(height), C.c_void_p(xy.ctypes.data),
                                     C.c_void_p(start.ctypes.data), C.c_void_p(items.ctypes.data), C.c_void_p(b.ctypes.data))
    return xy[:n], start, items[:m], b
