"""CPU ORACLE for ORB_SLAM2::ORBextractor::operator() -- TEST INFRASTRUCTURE ONLY.

This file restates /root/reference/src/ORBextractor.cc stage by stage.  The five
primitives the reference takes from its un-vendored OpenCV dependency (cv::resize,
cv::copyMakeBorder, cv::FAST, cv::GaussianBlur, cv::fastAtan2) are called through
the real OpenCV that ships in this image (opencv-python-headless 4.13.0), so the
external arithmetic is OpenCV's own; everything the reference itself computes
(scale tables, cell grid, retry, quadtree distribution, intensity centroid,
rotated BRIEF) is restated here in numpy float32 / integer arithmetic.

Parity status: the reference has no tests or golden vectors of its own
(SURVEY.md §4), so parity is pinned by (i) this oracle on cv2 4.13.0 and (ii) the
reference's own, unmodified translation unit compiled in oracle/_ref against the
OpenCV shim in oracle/shim (see oracle/Makefile); tests/test_oracle_ref.py checks
that (i) and (ii) agree bit for bit.  Two canonical rules are declared because the
reference's output is otherwise not a function of its input:
  B-1  equal-size nodes in DistributeOctTree's sort are ordered by node creation
       sequence (what the reference does under a monotonic allocator);
  B-2  float32 without FMA contraction; cos/sin = float32(double cos/sin).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.  The product (orbslam2_with_quadrics_b200) never does.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

try:  # cv2 is part of the image; the oracle is unusable without it
    import cv2
    cv2.setNumThreads(1)
except Exception as e:  # pragma: no cover
    cv2 = None
    _cv2_error = e

f32 = np.float32

PATCH_SIZE = 31            # src/ORBextractor.cc:72
HALF_PATCH_SIZE = 15       # :73
EDGE_THRESHOLD = 19        # :74
CELL_W = 30                # :769

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_pattern() -> np.ndarray:
    """512x2 int32 (x, y) -- bit_pattern_31_, src/ORBextractor.cc:150-408."""
    p = np.loadtxt(os.path.join(_HERE, "bit_pattern_31.txt"), dtype=np.int32)
    return p.reshape(512, 2)


def cv_round(x) -> int:
    """cvRound: round half to even."""
    return int(np.rint(x))


# --------------------------------------------------------------------------- ctor
@dataclass
class OrbParams:
    """ORBextractor::ORBextractor, src/ORBextractor.cc:410-470."""
    nfeatures: int
    scaleFactor: float
    nlevels: int
    iniThFAST: int
    minThFAST: int
    mvScaleFactor: np.ndarray = field(init=False)
    mvInvScaleFactor: np.ndarray = field(init=False)
    mvLevelSigma2: np.ndarray = field(init=False)
    mvInvLevelSigma2: np.ndarray = field(init=False)
    mnFeaturesPerLevel: list = field(init=False)
    umax: list = field(init=False)

    def __post_init__(self):
        # the member is `double scaleFactor` initialised from a float argument (ORBextractor.h:98)
        sfd = float(f32(self.scaleFactor))
        self.scaleFactor = sfd
        n = self.nlevels
        sf = np.ones(n, f32)
        sg = np.ones(n, f32)
        for i in range(1, n):
            sf[i] = f32(float(sf[i - 1]) * sfd)        # float*double -> double -> float  (:421)
            sg[i] = f32(sf[i] * sf[i])                  # (:422)
        self.mvScaleFactor = sf
        self.mvLevelSigma2 = sg
        self.mvInvScaleFactor = (f32(1.0) / sf).astype(f32)      # (:429)
        self.mvInvLevelSigma2 = (f32(1.0) / sg).astype(f32)      # (:430)
        factor = f32(1.0 / sfd)                                  # (:436)
        denom = f32(f32(1) - f32(math.pow(float(factor), float(n))))
        nd = f32(f32(f32(self.nfeatures) * f32(f32(1) - factor)) / denom)   # (:437)
        per = []
        s = 0
        for _ in range(n - 1):
            per.append(cv_round(nd))                             # (:442)
            s += per[-1]
            nd = f32(nd * factor)
        per.append(max(self.nfeatures - s, 0))                   # (:446)
        self.mnFeaturesPerLevel = per
        # umax (:454-469)
        hp = HALF_PATCH_SIZE
        root2 = f32(np.sqrt(f32(2.0)))
        vmax = int(math.floor(float(f32(f32(f32(hp) * root2) / f32(2)) + f32(1))))
        vmin = int(math.ceil(float(f32(f32(hp) * root2) / f32(2))))
        um = [0] * (hp + 1)
        for v in range(vmax + 1):
            um[v] = cv_round(math.sqrt(hp * hp - v * v))
        v0 = 0
        for v in range(hp, vmin - 1, -1):
            while um[v0] == um[v0 + 1]:
                v0 += 1
            um[v] = v0
            v0 += 1
        self.umax = um

    def level_sizes(self, width: int, height: int):
        """Level sizes as ComputePyramid computes them (:1111-1112): always from the original size."""
        out = []
        for l in range(self.nlevels):
            s = self.mvInvScaleFactor[l]
            out.append((cv_round(f32(f32(width) * s)), cv_round(f32(f32(height) * s))))
        return out


# --------------------------------------------------------------------------- stages
def compute_pyramid(p: OrbParams, image: np.ndarray):
    """ORBextractor::ComputePyramid (:1107-1132).  Returns the padded planes
    ((h+38) x (w+38) u8); level l's image is plane[19:-19, 19:-19]."""
    E = EDGE_THRESHOLD
    planes = []
    h0, w0 = image.shape
    for l, (w, h) in enumerate(p.level_sizes(w0, h0)):
        if l == 0:
            planes.append(cv2.copyMakeBorder(image, E, E, E, E, cv2.BORDER_REFLECT_101))
        else:
            prev = np.ascontiguousarray(planes[l - 1][E:-E, E:-E])
            lvl = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)       # (:1120)
            planes.append(cv2.copyMakeBorder(lvl, E, E, E, E, cv2.BORDER_REFLECT_101))  # (:1122)
    return planes


_fast_cache = {}


def _fast(window: np.ndarray, t: int):
    det = _fast_cache.get(t)
    if det is None:
        det = cv2.FastFeatureDetector_create(threshold=int(t), nonmaxSuppression=True,
                                             type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        _fast_cache[t] = det
    return det.detect(np.ascontiguousarray(window), None)


def cell_grid(cols: int, rows: int):
    """Grid geometry of ComputeKeyPointsOctTree for a level of cols x rows (:771-787)."""
    minBX = EDGE_THRESHOLD - 3
    minBY = minBX
    maxBX = cols - EDGE_THRESHOLD + 3
    maxBY = rows - EDGE_THRESHOLD + 3
    width = f32(maxBX - minBX)
    height = f32(maxBY - minBY)
    nCols = int(width / f32(CELL_W))
    nRows = int(height / f32(CELL_W))
    wCell = int(math.ceil(float(width / f32(nCols))))
    hCell = int(math.ceil(float(height / f32(nRows))))
    return minBX, minBY, maxBX, maxBY, nCols, nRows, wCell, hCell


def fast_candidates(p: OrbParams, level_img: np.ndarray, stats=None):
    """The cell loop of ComputeKeyPointsOctTree (:789-829).  level_img is the level
    (no border).  Returns int32 arrays (x, y, response) in the reference's push order,
    coordinates relative to (minBorderX, minBorderY)."""
    rows, cols = level_img.shape
    minBX, minBY, maxBX, maxBY, nCols, nRows, wCell, hCell = cell_grid(cols, rows)
    xs, ys, rs = [], [], []
    nwin = nretry = 0
    for i in range(nRows):
        iniY = minBY + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        if maxY > maxBY:
            maxY = maxBY
        for j in range(nCols):
            iniX = minBX + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            if maxX > maxBX:
                maxX = maxBX
            win = level_img[iniY:maxY, iniX:maxX]
            nwin += 1
            kps = _fast(win, p.iniThFAST)
            if len(kps) == 0:
                nretry += 1
                kps = _fast(win, p.minThFAST)
            for kp in kps:
                xs.append(int(kp.pt[0]) + j * wCell)
                ys.append(int(kp.pt[1]) + i * hCell)
                rs.append(int(kp.response))
    if stats is not None:
        stats["windows"] = stats.get("windows", 0) + nwin
        stats["retries"] = stats.get("retries", 0) + nretry
    return (np.asarray(xs, np.int32), np.asarray(ys, np.int32), np.asarray(rs, np.int32))


class _Node:
    __slots__ = ("ulx", "urx", "uly", "bry", "keys", "nomore", "seq")

    def __init__(self, ulx, urx, uly, bry, keys):
        self.ulx, self.urx, self.uly, self.bry = ulx, urx, uly, bry
        self.keys = keys            # int index array into the candidate list, in list order
        self.nomore = False
        self.seq = -1


def _divide(n: _Node, xs, ys):
    """ExtractorNode::DivideNode (:481-537)."""
    halfX = int(math.ceil(float(f32(n.urx - n.ulx) / f32(2))))
    halfY = int(math.ceil(float(f32(n.bry - n.uly) / f32(2))))
    mx, my = n.ulx + halfX, n.uly + halfY
    k = n.keys
    left = xs[k] < mx
    top = ys[k] < my
    c1 = _Node(n.ulx, mx, n.uly, my, k[left & top])
    c2 = _Node(mx, n.urx, n.uly, my, k[~left & top])
    c3 = _Node(n.ulx, mx, my, n.bry, k[left & ~top])
    c4 = _Node(mx, n.urx, my, n.bry, k[~left & ~top])
    for c in (c1, c2, c3, c4):
        if len(c.keys) == 1:
            c.nomore = True
    return c1, c2, c3, c4


def distribute_octree(xs, ys, rs, minX, maxX, minY, maxY, N, trace=None):
    """ORBextractor::DistributeOctTree (:539-763) with canonical rule B-1.
    xs, ys, rs: candidate list (box coordinates).  Returns indices of the kept
    candidates in the reference's output order."""
    xs = np.asarray(xs)
    ys = np.asarray(ys)
    rs = np.asarray(rs)
    M = len(xs)
    ratio = f32(maxX - minX) / f32(maxY - minY)
    nIni = int(math.floor(float(ratio) + 0.5))          # std::round(float), half away from zero (:543)
    hX = f32(f32(maxX - minX) / f32(nIni))               # (:545)
    seq = [0]

    def stamp(n):
        n.seq = seq[0]
        seq[0] += 1

    roots = []
    root_of = (xs.astype(f32) / hX).astype(np.int32) if M else np.zeros(0, np.int32)   # (:569) truncation
    allidx = np.arange(M)
    for i in range(nIni):
        ulx = int(f32(hX * f32(i)))                       # (:555)
        urx = int(f32(hX * f32(i + 1)))                   # (:556)
        n = _Node(ulx, urx, 0, maxY - minY, allidx[root_of == i])
        stamp(n)
        roots.append(n)
    lst = []
    for n in roots:                                       # (:574-585)
        if len(n.keys) == 1:
            n.nomore = True
            lst.append(n)
        elif len(n.keys) > 1:
            lst.append(n)
    finish = False
    while not finish:
        prev_size = len(lst)
        n_expand = 0
        size_and_node = []
        new_front = []
        kept = []
        for n in lst:                                     # phase A (:606-665)
            if n.nomore:
                kept.append(n)
                continue
            for c in _divide(n, xs, ys):
                if len(c.keys) > 0:
                    stamp(c)
                    new_front.insert(0, c)
                    if len(c.keys) > 1:
                        n_expand += 1
                        size_and_node.append(c)
        lst = new_front + kept
        if trace is not None:
            trace.append(("A", len(lst)))
        if len(lst) >= N or len(lst) == prev_size:        # (:669)
            finish = True
        elif len(lst) + n_expand * 3 > N:                 # (:673)
            while not finish:                             # phase B (:676-737)
                prev_size = len(lst)
                prev_nodes = sorted(size_and_node, key=lambda c: (len(c.keys), c.seq))   # rule B-1 (:684)
                size_and_node = []
                for n in reversed(prev_nodes):
                    for c in _divide(n, xs, ys):
                        if len(c.keys) > 0:
                            stamp(c)
                            lst.insert(0, c)
                            if len(c.keys) > 1:
                                size_and_node.append(c)
                    for q in range(len(lst)):
                        if lst[q] is n:
                            del lst[q]
                            break
                    if len(lst) >= N:                     # (:730)
                        break
                if trace is not None:
                    trace.append(("B", len(lst)))
                if len(lst) >= N or len(lst) == prev_size:
                    finish = True
    out = []
    for n in lst:                                         # (:741-760) first max wins
        k = n.keys
        out.append(int(k[int(np.argmax(rs[k]))]))
    return np.asarray(out, np.int64)


def ic_angles(p: OrbParams, plane: np.ndarray, kx: np.ndarray, ky: np.ndarray) -> np.ndarray:
    """IC_Angle (:77-104) for keypoints at level coordinates (kx, ky); plane is the padded level."""
    E = EDGE_THRESHOLD
    n = len(kx)
    out = np.zeros(n, f32)
    if n == 0:
        return out
    hp = HALF_PATCH_SIZE
    d = np.arange(-hp, hp + 1)
    um = np.asarray(p.umax)
    mask = (np.abs(d)[None, :] <= um[np.abs(d)][:, None])            # [v, u]
    yy = (ky[:, None, None] + E + d[None, :, None])
    xx = (kx[:, None, None] + E + d[None, None, :])
    patch = plane[yy, xx].astype(np.int64) * mask[None]
    m10 = (patch * d[None, None, :]).sum(axis=(1, 2))
    m01 = (patch * d[None, :, None]).sum(axis=(1, 2))
    for i in range(n):
        out[i] = f32(cv2.fastAtan2(float(m01[i]), float(m10[i])))    # (:103)
    return out


def blur_level(level_img: np.ndarray) -> np.ndarray:
    """clone + GaussianBlur(7x7, 2, 2, BORDER_REFLECT_101) (:1085-1086)."""
    return cv2.GaussianBlur(np.ascontiguousarray(level_img), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)


FACTOR_PI = f32(math.pi / float(f32(180.0)))     # (:107)


def orb_descriptors(blurred: np.ndarray, kx, ky, angles, pattern: np.ndarray) -> np.ndarray:
    """computeOrbDescriptor (:108-147) under canonical rule B-2."""
    n = len(kx)
    desc = np.zeros((n, 32), np.uint8)
    if n == 0:
        return desc
    ang = (angles.astype(f32) * FACTOR_PI).astype(f32)
    a = np.cos(ang.astype(np.float64)).astype(f32)[:, None]
    b = np.sin(ang.astype(np.float64)).astype(f32)[:, None]
    px = pattern[:, 0].astype(f32)[None, :]
    py = pattern[:, 1].astype(f32)[None, :]
    row = np.rint((px * b).astype(f32) + (py * a).astype(f32)).astype(np.int64)     # (:119)
    col = np.rint((px * a).astype(f32) - (py * b).astype(f32)).astype(np.int64)     # (:120)
    val = blurred[ky[:, None] + row, kx[:, None] + col].astype(np.int32)             # [n, 512]
    bits = (val[:, 0::2] < val[:, 1::2]).astype(np.uint8)                            # [n, 256]
    bits = bits.reshape(n, 32, 8)
    desc = (bits << np.arange(8, dtype=np.uint8)[None, None, :]).sum(axis=2).astype(np.uint8)
    return desc


# --------------------------------------------------------------------------- whole path
@dataclass
class OracleResult:
    n: int
    keypoints: np.ndarray            # [n] structured: x,y,size,angle,response f32, octave,class_id i32 (cv::KeyPoint order)
    descriptors: np.ndarray          # [n,32] u8
    pyramid: list                    # padded planes
    candidates: list                 # per level (x, y, response) int32, box coordinates, push order
    kept: list                       # per level (x, y, response) int32, LEVEL coordinates (box + 16), octree order
    angles: list                     # per level f32
    blurred: list                    # per level u8 level image (None where the level had no keypoints)
    stats: dict


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


class ORBextractor:
    """Mirror of ORB_SLAM2::ORBextractor's public surface (include/ORBextractor.h:45-85)."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST):
        if cv2 is None:  # pragma: no cover
            raise RuntimeError("oracle needs cv2: %r" % (_cv2_error,))
        self.p = OrbParams(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        self.pattern = load_pattern()
        self.mvImagePyramid = [None] * nlevels

    def GetLevels(self): return self.p.nlevels
    def GetScaleFactor(self): return float(f32(self.p.scaleFactor))
    def GetScaleFactors(self): return self.p.mvScaleFactor.copy()
    def GetInverseScaleFactors(self): return self.p.mvInvScaleFactor.copy()
    def GetScaleSigmaSquares(self): return self.p.mvLevelSigma2.copy()
    def GetInverseScaleSigmaSquares(self): return self.p.mvInvLevelSigma2.copy()

    def __call__(self, image: np.ndarray, mask=None) -> OracleResult | None:
        """ORBextractor::operator() (:1043-1105).  Returns None for an empty image (:1046-1047)."""
        if image is None or image.size == 0:
            return None
        assert image.dtype == np.uint8 and image.ndim == 2
        p = self.p
        E = EDGE_THRESHOLD
        planes = compute_pyramid(p, image)
        self.mvImagePyramid = [pl[E:-E, E:-E] for pl in planes]
        stats = {}
        cands, kept, angles, blurred = [], [], [], []
        kps_all, desc_all = [], []
        for l in range(p.nlevels):
            lvl = self.mvImagePyramid[l]
            rows, cols = lvl.shape
            minBX, minBY, maxBX, maxBY, *_ = cell_grid(cols, rows)
            cx, cy, cr = fast_candidates(p, lvl, stats)
            cands.append((cx, cy, cr))
            sel = distribute_octree(cx, cy, cr, minBX, maxBX, minBY, maxBY, p.mnFeaturesPerLevel[l]) \
                if len(cx) else np.zeros(0, np.int64)
            kx = (cx[sel] + minBX).astype(np.int64)            # (:843-844)
            ky = (cy[sel] + minBY).astype(np.int64)
            kr = cr[sel]
            kept.append((kx.astype(np.int32), ky.astype(np.int32), kr.astype(np.int32)))
            ang = ic_angles(p, planes[l], kx, ky)              # (:851-852), un-blurred level
            angles.append(ang)
            if len(kx) == 0:                                   # (:1081-1082)
                blurred.append(None)
                continue
            bl = blur_level(lvl)
            blurred.append(bl)
            desc_all.append(orb_descriptors(bl, kx, ky, ang, self.pattern))
            k = np.zeros(len(kx), KP_DTYPE)
            sc = p.mvScaleFactor[l]
            k["x"] = kx.astype(f32) * sc if l else kx.astype(f32)       # (:1095-1101)
            k["y"] = ky.astype(f32) * sc if l else ky.astype(f32)
            k["size"] = f32(int(f32(f32(PATCH_SIZE) * sc)))             # (:837,846)
            k["angle"] = ang
            k["response"] = kr.astype(f32)
            k["octave"] = l
            k["class_id"] = -1
            kps_all.append(k)
        if kps_all:
            kps = np.concatenate(kps_all)
            desc = np.concatenate(desc_all)
        else:
            kps = np.zeros(0, KP_DTYPE)
            desc = np.zeros((0, 32), np.uint8)
        return OracleResult(len(kps), kps, desc, planes, cands, kept, angles, blurred, stats)
