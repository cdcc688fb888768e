"""CPU restatement of ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)
(reference src/ORBmatcher.cc:1328-1470) with Frame::GetFeaturesInArea (src/Frame.cc:327-380), DescriptorDistance
(src/ORBmatcher.cc:1647-1663) and ComputeThreeMaxima (:1601-1642): the per-frame descriptor consumer of
Tracking::TrackWithMotionModel (src/Tracking.cc:885, :891).

TEST INFRASTRUCTURE: the checker of ``orbx_search_by_projection``; only tests/ may import it.  The three small matrix
products are the REAL OpenCV (cv2.gemm, 4.13.0), everything else is float32 numpy scalars without FMA (rule B-2).
Pinned against the reference's own lines compiled against a stub (oracle/_ref/libstereoref.so,
``ref_search_by_projection`` below) by tests/test_match_oracle.py.

State convention (the reference's only call sites fill CurrentFrame.mvpMapPoints with NULL first, src/Tracking.cc:871,
:890): the current frame starts without map points.  The result names, per current keypoint, the index of the LastFrame
keypoint whose map point it holds at return, or -1.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

f32 = np.float32
GRID_COLS, GRID_ROWS = 64, 48          # include/Frame.h:39-40
TH_HIGH, HISTO_LENGTH = 100, 30        # src/ORBmatcher.cc:37, :39
_HERE = os.path.dirname(os.path.abspath(__file__))


def _gemm(A, B, Cm=None, alpha=1.0, flags=0):
    import cv2
    return cv2.gemm(np.ascontiguousarray(A, f32), np.ascontiguousarray(B, f32), alpha, None if Cm is None else
                    np.ascontiguousarray(Cm, f32), 0.0 if Cm is None else 1.0, flags=flags)


def motion_flags(Tcw_cur, Tcw_last, mb, mono):
    """bForward / bBackward (:1337-1349)."""
    import cv2
    Tc = np.asarray(Tcw_cur, f32).reshape(4, 4)
    Tl = np.asarray(Tcw_last, f32).reshape(4, 4)
    twc = _gemm(Tc[:3, :3], Tc[:3, 3:4], None, -1.0, cv2.GEMM_1_T)            # -Rcw.t() * tcw
    tlc = _gemm(Tl[:3, :3], twc, Tl[:3, 3:4])                                  # Rlw * twc + tlw
    z = f32(tlc[2, 0])
    return bool(z > f32(mb) and not mono), bool(-z > f32(mb) and not mono)


def _round_half_away(x) -> int:
    x = float(x)
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def features_in_area(x, y, r, min_level, max_level, xy_un, octave, cell_start, cell_items, bounds):
    """Frame::GetFeaturesInArea (src/Frame.cc:327-380), candidates in the reference's visiting order."""
    mnx, mxx, mny, mxy = [f32(b) for b in bounds]
    winv = f32(f32(GRID_COLS) / f32(mxx - mnx))
    hinv = f32(f32(GRID_ROWS) / f32(mxy - mny))
    x, y, r = f32(x), f32(y), f32(r)
    out = []
    c0 = max(0, int(math.floor(f32(f32(f32(x - mnx) - r) * winv))))
    if c0 >= GRID_COLS:
        return out
    c1 = min(GRID_COLS - 1, int(math.ceil(f32(f32(f32(x - mnx) + r) * winv))))
    if c1 < 0:
        return out
    r0 = max(0, int(math.floor(f32(f32(f32(y - mny) - r) * hinv))))
    if r0 >= GRID_ROWS:
        return out
    r1 = min(GRID_ROWS - 1, int(math.ceil(f32(f32(f32(y - mny) + r) * hinv))))
    if r1 < 0:
        return out
    check = (min_level > 0) or (max_level >= 0)
    for ix in range(c0, c1 + 1):
        for iy in range(r0, r1 + 1):
            c = ix * GRID_ROWS + iy
            for k in cell_items[cell_start[c]:cell_start[c + 1]]:
                if check:
                    if octave[k] < min_level:
                        continue
                    if max_level >= 0 and octave[k] > max_level:
                        continue
                if abs(f32(xy_un[k, 0] - x)) < r and abs(f32(xy_un[k, 1] - y)) < r:
                    out.append(int(k))
    return out


def search_by_projection(world, mp_desc, mp_obs, outlier, last_octave, last_angle, Tcw_cur, Tcw_last, xy_un, cur_octave,
                         cur_angle, desc, u_right, cell_start, cell_items, bounds, K4, mbf, mb, sf, th, mono,
                         check_orientation=True):
    """Returns (nmatches, cur_match int32[nC])."""
    nL, nC = len(mp_obs), len(desc)
    Tc = np.asarray(Tcw_cur, f32).reshape(4, 4)
    fx, fy, cx, cy = [f32(v) for v in K4]
    mnx, mxx, mny, mxy = [f32(b) for b in bounds]
    fwd, bwd = motion_flags(Tcw_cur, Tcw_last, mb, mono)
    factor = f32(f32(1.0) / f32(HISTO_LENGTH))
    desc64 = np.ascontiguousarray(desc).view(np.uint64).reshape(nC, 4) if nC else np.zeros((0, 4), np.uint64)
    q64 = np.ascontiguousarray(mp_desc).view(np.uint64).reshape(nL, 4) if nL else np.zeros((0, 4), np.uint64)
    holder = np.full(nC, -1, np.int64)            # CurrentFrame.mvpMapPoints as LastFrame indices
    rot_hist = [[] for _ in range(HISTO_LENGTH)]
    nmatches = 0
    W = np.asarray(world, f32).reshape(nL, 3)
    for i in range(nL):
        if mp_obs[i] < 0 or outlier[i]:
            continue
        X = _gemm(Tc[:3, :3], W[i].reshape(3, 1), Tc[:3, 3:4])                 # Rcw*x3Dw+tcw, one 3x1 product as in the reference
        xc, yc, zc = f32(X[0, 0]), f32(X[1, 0]), f32(X[2, 0])
        with np.errstate(divide="ignore"):
            invzc = f32(np.float64(1.0) / np.float64(zc))
        if invzc < 0:
            continue
        u = f32(f32(f32(fx * xc) * invzc) + cx)
        v = f32(f32(f32(fy * yc) * invzc) + cy)
        if u < mnx or u > mxx or v < mny or v > mxy:
            continue
        if not (np.isfinite(u) and np.isfinite(v)):
            continue                               # NaN would index the grid with an undefined int cast; defined as "no candidates"
        lo = int(last_octave[i])
        radius = f32(f32(th) * f32(sf[lo]))
        if fwd:
            cands = features_in_area(u, v, radius, lo, -1, xy_un, cur_octave, cell_start, cell_items, bounds)
        elif bwd:
            cands = features_in_area(u, v, radius, 0, lo, xy_un, cur_octave, cell_start, cell_items, bounds)
        else:
            cands = features_in_area(u, v, radius, lo - 1, lo + 1, xy_un, cur_octave, cell_start, cell_items, bounds)
        if not cands:
            continue
        best, best_idx = 256, -1
        for i2 in cands:
            if holder[i2] >= 0 and mp_obs[holder[i2]] > 0:
                continue
            if u_right is not None and u_right[i2] > 0:
                ur = f32(u - f32(f32(mbf) * invzc))
                if abs(f32(ur - f32(u_right[i2]))) > radius:
                    continue
            d = int(sum(bin(int(a ^ b)).count("1") for a, b in zip(q64[i], desc64[i2])))
            if d < best:
                best, best_idx = d, i2
        if best <= TH_HIGH:
            holder[best_idx] = i
            nmatches += 1
            if check_orientation:
                rot = f32(f32(last_angle[i]) - f32(cur_angle[best_idx]))
                if rot < 0.0:
                    rot = f32(rot + f32(360.0))
                b = _round_half_away(f32(rot * factor))
                if b == HISTO_LENGTH:
                    b = 0
                rot_hist[b].append(best_idx)
    if check_orientation:
        ind = compute_three_maxima([len(h) for h in rot_hist])
        for b in range(HISTO_LENGTH):
            if b not in ind:
                for i2 in rot_hist[b]:
                    holder[i2] = -1
                    nmatches -= 1
    return nmatches, holder.astype(np.int32)


def compute_three_maxima(sizes):
    """ORBmatcher::ComputeThreeMaxima (:1601-1642) on the bin sizes; returns (ind1, ind2, ind3), -1 = none."""
    max1 = max2 = max3 = 0
    ind1 = ind2 = ind3 = -1
    for i, s in enumerate(sizes):
        if s > max1:
            max3, max2, max1 = max2, max1, s
            ind3, ind2, ind1 = ind2, ind1, i
        elif s > max2:
            max3, max2 = max2, s
            ind3, ind2 = ind2, i
        elif s > max3:
            max3, ind3 = s, i
    if f32(max2) < f32(f32(0.1) * f32(max1)):
        ind2 = ind3 = -1
    elif f32(max3) < f32(f32(0.1) * f32(max1)):
        ind3 = -1
    return ind1, ind2, ind3


# ----------------------------------------------------------------------------- the reference's own lines
_ref = None


def ref_available() -> bool:
    if not os.path.exists(os.path.join(_HERE, "_ref", "libstereoref.so")):
        return False
    try:
        return hasattr(C.CDLL(os.path.join(_HERE, "_ref", "libstereoref.so")), "matchref_search_by_projection")
    except OSError:
        return False


def ref_search_by_projection(world, mp_desc, mp_obs, outlier, last_octave, last_angle, Tcw_cur, Tcw_last, xy_un, cur_octave,
                             cur_angle, desc, u_right, cell_start, cell_items, bounds, K4, mbf, mb, sf, th, mono,
                             check_orientation=True):
    """Same call shape, executed by the reference's own lines (oracle/_ref/libstereoref.so)."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libstereoref.so"))
        _ref.matchref_search_by_projection.restype = C.c_int
    nL, nC = len(mp_obs), len(desc)
    kp = np.zeros((nC, 7), f32)
    kp[:, 0:2] = np.asarray(xy_un, f32).reshape(nC, 2)
    kp[:, 3] = np.asarray(cur_angle, f32)
    kp[:, 5] = np.asarray(cur_octave, np.int32).view(f32)
    arrs = dict(world=np.ascontiguousarray(world, f32), mp_desc=np.ascontiguousarray(mp_desc, np.uint8),
                mp_obs=np.ascontiguousarray(mp_obs, np.int32), outlier=np.ascontiguousarray(outlier, np.uint8),
                lo=np.ascontiguousarray(last_octave, np.int32), la=np.ascontiguousarray(last_angle, f32),
                tc=np.ascontiguousarray(Tcw_cur, f32).reshape(16), tl=np.ascontiguousarray(Tcw_last, f32).reshape(16),
                desc=np.ascontiguousarray(desc, np.uint8), cs=np.ascontiguousarray(cell_start, np.int32),
                ci=np.ascontiguousarray(cell_items, np.int32), b=np.ascontiguousarray(bounds, f32),
                k=np.ascontiguousarray(K4, f32), sf=np.ascontiguousarray(sf, f32))
    ur = None if u_right is None else np.ascontiguousarray(u_right, f32)
    out = np.full(nC, -1, np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    n = _ref.matchref_search_by_projection(
        C.c_int(nL), p(arrs["world"]), p(arrs["mp_desc"]), p(arrs["mp_obs"]), p(arrs["outlier"]), p(arrs["lo"]), p(arrs["la"]),
        p(arrs["tc"]), p(arrs["tl"]), C.c_int(nC), p(kp), p(arrs["desc"]), C.c_void_p(0) if ur is None else p(ur),
        p(arrs["cs"]), p(arrs["ci"]), p(arrs["b"]), p(arrs["k"]), C.c_float(mbf), C.c_float(mb), p(arrs["sf"]),
        C.c_int(len(arrs["sf"])), C.c_float(th), C.c_int(int(mono)), C.c_int(int(check_orientation)), p(out))
    return int(n), out


# ============================================================================= SearchLocalPoints' matcher
def radius_by_viewing_cos(view_cos):
    """ORBmatcher::RadiusByViewingCos (src/ORBmatcher.cc:131-137): the float is compared with the double 0.998."""
    return f32(2.5) if float(f32(view_cos)) > 0.998 else f32(4.0)


def search_local_points(in_view, proj_x, proj_y, proj_xr, scale_level, view_cos, mp_desc, mp_obs, xy_un, cur_octave, desc, u_right,
                        cur_obs, cell_start, cell_items, bounds, sf, th, nnratio=0.8):
    """ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th) (src/ORBmatcher.cc:45-129).
    cur_obs[i2] < 0: F.mvpMapPoints[i2] is NULL, else the Observations() of the point it holds.
    Returns (nmatches, new_match int32[nC]: index into vpMapPoints F.mvpMapPoints[i2] was set to, else -1)."""
    nP, nC = len(mp_obs), len(desc)
    desc64 = np.ascontiguousarray(desc).view(np.uint64).reshape(nC, 4) if nC else np.zeros((0, 4), np.uint64)
    q64 = np.ascontiguousarray(mp_desc).view(np.uint64).reshape(nP, 4) if nP else np.zeros((0, 4), np.uint64)
    holder_obs = np.asarray(cur_obs, np.int64).copy()          # Observations() of what each keypoint holds, -1 = NULL
    new_match = np.full(nC, -1, np.int32)
    b_factor = float(f32(th)) != 1.0
    nmatches = 0
    for i in range(nP):
        if not in_view[i]:
            continue
        lvl = int(scale_level[i])
        r = radius_by_viewing_cos(view_cos[i])
        if b_factor:
            r = f32(r * f32(th))
        rad = f32(r * f32(sf[lvl]))
        cands = features_in_area(proj_x[i], proj_y[i], rad, lvl - 1, lvl, xy_un, cur_octave, cell_start, cell_items, bounds)
        if not cands:
            continue
        best = best2 = 256
        lvl1 = lvl2 = -1
        best_idx = -1
        for i2 in cands:
            if holder_obs[i2] > 0:
                continue
            if u_right is not None and u_right[i2] > 0:
                if abs(f32(f32(proj_xr[i]) - f32(u_right[i2]))) > rad:
                    continue
            d = int(sum(bin(int(a ^ b)).count("1") for a, b in zip(q64[i], desc64[i2])))
            if d < best:
                best2, best = best, d
                lvl2, lvl1 = lvl1, int(cur_octave[i2])
                best_idx = i2
            elif d < best2:
                lvl2 = int(cur_octave[i2])
                best2 = d
        if best <= TH_HIGH:
            if lvl1 == lvl2 and f32(best) > f32(f32(nnratio) * f32(best2)):
                continue
            holder_obs[best_idx] = mp_obs[i]
            new_match[best_idx] = i
            nmatches += 1
    return nmatches, new_match


def ref_search_local_points(in_view, proj_x, proj_y, proj_xr, scale_level, view_cos, mp_desc, mp_obs, xy_un, cur_octave, desc,
                            u_right, cur_obs, cell_start, cell_items, bounds, sf, th, nnratio=0.8):
    """Same call shape, executed by the reference's own lines (oracle/_ref/libstereoref.so)."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libstereoref.so"))
        _ref.matchref_search_by_projection.restype = C.c_int
    _ref.matchref_search_local_points.restype = C.c_int
    nP, nC = len(mp_obs), len(desc)
    kp = np.zeros((nC, 7), f32)
    kp[:, 0:2] = np.asarray(xy_un, f32).reshape(nC, 2)
    kp[:, 5] = np.asarray(cur_octave, np.int32).view(f32)
    a = [np.ascontiguousarray(in_view, np.uint8), np.ascontiguousarray(proj_x, f32), np.ascontiguousarray(proj_y, f32),
         np.ascontiguousarray(proj_xr, f32), np.ascontiguousarray(scale_level, np.int32), np.ascontiguousarray(view_cos, f32),
         np.ascontiguousarray(mp_desc, np.uint8), np.ascontiguousarray(mp_obs, np.int32)]
    b = [np.ascontiguousarray(desc, np.uint8), None if u_right is None else np.ascontiguousarray(u_right, f32),
         np.ascontiguousarray(cur_obs, np.int32), np.ascontiguousarray(cell_start, np.int32), np.ascontiguousarray(cell_items, np.int32),
         np.ascontiguousarray(bounds, f32), np.ascontiguousarray(sf, f32)]
    out = np.full(nC, -1, np.int32)
    p = lambda x: C.c_void_p(0) if x is None else C.c_void_p(x.ctypes.data)
    n = _ref.matchref_search_local_points(C.c_int(nP), *[p(x) for x in a], C.c_int(nC), p(kp), p(b[0]), p(b[1]), p(b[2]), p(b[3]),
                                          p(b[4]), p(b[5]), p(b[6]), C.c_int(len(b[6])), C.c_float(th), C.c_float(nnratio), p(out))
    return int(n), out


# ============================================================================= SearchByBoW(KeyFrame*, Frame&)
TH_LOW = 50                               # src/ORBmatcher.cc:38


def search_by_bow(kf_desc, kf_valid, kf_angle, kf_fv_nodes, kf_fv_features, f_desc, f_angle, f_fv_nodes, f_fv_features, nnratio=0.7,
                  check_orientation=True):
    """ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches) (src/ORBmatcher.cc:159-288).  kf_valid: 0 = no map
    point, 1 = good, 2 = isBad().  FeatureVectors as (node, feature) pairs in map / push_back order.
    Returns (nmatches, match int32[F.N]: KeyFrame feature index or -1)."""
    nF = len(f_desc)
    kd = np.ascontiguousarray(kf_desc).view(np.uint64).reshape(-1, 4)
    fd = np.ascontiguousarray(f_desc).view(np.uint64).reshape(-1, 4) if nF else np.zeros((0, 4), np.uint64)
    match = np.full(nF, -1, np.int32)
    rot_hist = [[] for _ in range(HISTO_LENGTH)]
    factor = f32(f32(1.0) / f32(HISTO_LENGTH))
    nmatches = 0
    kf_groups, f_groups = {}, {}
    for n, i in zip(kf_fv_nodes, kf_fv_features):
        kf_groups.setdefault(int(n), []).append(int(i))
    for n, i in zip(f_fv_nodes, f_fv_features):
        f_groups.setdefault(int(n), []).append(int(i))
    for node in sorted(kf_groups):                        # the merge walk visits the common nodes in ascending order
        if node not in f_groups:
            continue
        for ikf in kf_groups[node]:
            if kf_valid[ikf] != 1:
                continue
            best1 = best2 = 256
            best_idx = -1
            for i_f in f_groups[node]:
                if match[i_f] >= 0:
                    continue
                d = int(sum(bin(int(a ^ b)).count("1") for a, b in zip(kd[ikf], fd[i_f])))
                if d < best1:
                    best2, best1, best_idx = best1, d, i_f
                elif d < best2:
                    best2 = d
            if best1 <= TH_LOW and f32(best1) < f32(f32(nnratio) * f32(best2)):
                match[best_idx] = ikf
                if check_orientation:
                    rot = f32(f32(kf_angle[ikf]) - f32(f_angle[best_idx]))
                    if rot < 0.0:
                        rot = f32(rot + f32(360.0))
                    b = _round_half_away(f32(rot * factor))
                    if b == HISTO_LENGTH:
                        b = 0
                    rot_hist[b].append(best_idx)
                nmatches += 1
    if check_orientation:
        ind = compute_three_maxima([len(h) for h in rot_hist])
        for b in range(HISTO_LENGTH):
            if b in ind:
                continue
            for i_f in rot_hist[b]:
                match[i_f] = -1
                nmatches -= 1
    return nmatches, match


def ref_search_by_bow(kf_desc, kf_valid, kf_angle, kf_fv_nodes, kf_fv_features, f_desc, f_angle, f_fv_nodes, f_fv_features,
                      nnratio=0.7, check_orientation=True):
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libstereoref.so"))
        _ref.matchref_search_by_projection.restype = C.c_int
    _ref.matchref_search_by_bow.restype = C.c_int
    a = [np.ascontiguousarray(kf_desc, np.uint8), np.ascontiguousarray(kf_valid, np.uint8), np.ascontiguousarray(kf_angle, f32)]
    b = [np.ascontiguousarray(kf_fv_nodes, np.uint32), np.ascontiguousarray(kf_fv_features, np.uint32)]
    c = [np.ascontiguousarray(f_desc, np.uint8), np.ascontiguousarray(f_angle, f32)]
    d = [np.ascontiguousarray(f_fv_nodes, np.uint32), np.ascontiguousarray(f_fv_features, np.uint32)]
    out = np.full(len(c[0]), -1, np.int32)
    p = lambda x: C.c_void_p(x.ctypes.data)
    n = _ref.matchref_search_by_bow(C.c_int(len(a[1])), p(a[0]), p(a[1]), p(a[2]), C.c_int(len(b[0])), p(b[0]), p(b[1]),
                                    C.c_int(len(c[0])), p(c[0]), p(c[1]), C.c_int(len(d[0])), p(d[0]), p(d[1]), C.c_float(nnratio),
                                    C.c_int(int(check_orientation)), p(out))
    return int(n), out


# ============================================================================= Relocalization's SearchByProjection(Frame&, KeyFrame*)
_libm = None


def _logf(x) -> np.float32:
    """std::log(float) of the C library the reference links (MapPoint.cc:410: `log(ratio)` with a float argument under
    `using namespace std` resolves to the float overload)."""
    global _libm
    if _libm is None:
        _libm = C.CDLL("libm.so.6")
        _libm.logf.restype = C.c_float
        _libm.logf.argtypes = [C.c_float]
    return f32(_libm.logf(C.c_float(float(x))))


def kf_prepare(valid, world, min_dist, max_dist, Tcw_cur, log_scale_factor, nlevels):
    """What a caller of orbx_search_by_projection_kf stages per KeyFrame map point, by the arithmetic of the reference's lines:
    Ow = -Rcw.t()*tcw (src/ORBmatcher.cc:1476-1478), dist3D = cv::norm(x3Dw - Ow) and the distance-invariance test (:1510-1518,
    src/MapPoint.cc:373-383), nPredictedLevel = pMP->PredictScale(dist3D, &CurrentFrame) (:1520, src/MapPoint.cc:402-417).
    Returns (in_range uint8[n], pred_level int32[n])."""
    import cv2
    n = len(valid)
    Tc = np.asarray(Tcw_cur, f32).reshape(4, 4)
    Ow = _gemm(Tc[:3, :3], Tc[:3, 3:4], None, -1.0, cv2.GEMM_1_T)
    W = np.asarray(world, f32).reshape(n, 3)
    in_range = np.zeros(n, np.uint8)
    pred = np.zeros(n, np.int32)
    for i in range(n):
        if not valid[i]:
            continue
        PO = (W[i].reshape(3, 1) - Ow).astype(f32)
        dist3D = f32(cv2.norm(PO))
        if dist3D < f32(f32(0.8) * f32(min_dist[i])) or dist3D > f32(f32(1.2) * f32(max_dist[i])):
            continue
        in_range[i] = 1
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = f32(f32(max_dist[i]) / dist3D)
        q = f32(_logf(ratio) / f32(log_scale_factor))
        ns = int(math.ceil(float(q))) if np.isfinite(q) else 0
        pred[i] = 0 if ns < 0 else (nlevels - 1 if ns >= nlevels else ns)
    return in_range, pred


def search_by_projection_kf(valid, world, mp_desc, in_range, pred_level, kf_angle, Tcw_cur, xy_un, cur_octave, cur_angle, desc,
                            cur_held, cell_start, cell_items, bounds, K4, sf, th, orb_dist, check_orientation=True):
    """ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1472-1599).
    valid[i]: 0 = no map point, 1 = good, 2 = isBad(), 3 = in sAlreadyFound; (in_range, pred_level) from kf_prepare;
    cur_held[i2] != 0: CurrentFrame.mvpMapPoints[i2] is not NULL at entry.
    Returns (nmatches, new_match int32[nC]: index of the KeyFrame point the call set mvpMapPoints[i2] to and kept, else -1)."""
    nP, nC = len(valid), len(desc)
    Tc = np.asarray(Tcw_cur, f32).reshape(4, 4)
    fx, fy, cx, cy = [f32(v) for v in K4]
    mnx, mxx, mny, mxy = [f32(b) for b in bounds]
    factor = f32(f32(1.0) / f32(HISTO_LENGTH))
    desc64 = np.ascontiguousarray(desc).view(np.uint64).reshape(nC, 4) if nC else np.zeros((0, 4), np.uint64)
    q64 = np.ascontiguousarray(mp_desc).view(np.uint64).reshape(nP, 4) if nP else np.zeros((0, 4), np.uint64)
    held = np.asarray(cur_held).astype(bool).copy()
    new_match = np.full(nC, -1, np.int32)
    rot_hist = [[] for _ in range(HISTO_LENGTH)]
    nmatches = 0
    W = np.asarray(world, f32).reshape(nP, 3)
    for i in range(nP):
        if valid[i] != 1:
            continue
        X = _gemm(Tc[:3, :3], W[i].reshape(3, 1), Tc[:3, 3:4])
        xc, yc, zc = f32(X[0, 0]), f32(X[1, 0]), f32(X[2, 0])
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            invzc = f32(np.float64(1.0) / np.float64(zc))
            u = f32(f32(f32(fx * xc) * invzc) + cx)
            v = f32(f32(f32(fy * yc) * invzc) + cy)
        if u < mnx or u > mxx or v < mny or v > mxy:
            continue
        if not (np.isfinite(u) and np.isfinite(v)):
            continue                               # NaN: undefined grid index in the reference; defined as "no candidates"
        if not in_range[i]:
            continue
        lvl = int(pred_level[i])
        radius = f32(f32(th) * f32(sf[lvl]))
        cands = features_in_area(u, v, radius, lvl - 1, lvl + 1, xy_un, cur_octave, cell_start, cell_items, bounds)
        if not cands:
            continue
        best, best_idx = 256, -1
        for i2 in cands:
            if held[i2]:
                continue
            d = int(sum(bin(int(a ^ b)).count("1") for a, b in zip(q64[i], desc64[i2])))
            if d < best:
                best, best_idx = d, i2
        if best <= orb_dist:
            held[best_idx] = True
            new_match[best_idx] = i
            nmatches += 1
            if check_orientation:
                rot = f32(f32(kf_angle[i]) - f32(cur_angle[best_idx]))
                if rot < 0.0:
                    rot = f32(rot + f32(360.0))
                b = _round_half_away(f32(rot * factor))
                if b == HISTO_LENGTH:
                    b = 0
                rot_hist[b].append(best_idx)
    if check_orientation:
        ind = compute_three_maxima([len(h) for h in rot_hist])
        for b in range(HISTO_LENGTH):
            if b not in ind:
                for i2 in rot_hist[b]:
                    new_match[i2] = -1
                    nmatches -= 1
    return nmatches, new_match


def _kp7(xy_un, angle, octave):
    n = len(octave)
    kp = np.zeros((n, 7), f32)
    kp[:, 0:2] = np.asarray(xy_un, f32).reshape(n, 2)
    kp[:, 3] = np.asarray(angle, f32)
    kp[:, 5] = np.asarray(octave, np.int32).view(f32)
    return kp


def _load_ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libstereoref.so"))
        _ref.matchref_search_by_projection.restype = C.c_int
    return _ref


def ref_has(name: str) -> bool:
    return ref_available() and hasattr(_load_ref(), name)


def ref_search_by_projection_kf(valid, world, mp_desc, min_dist, max_dist, kf_angle, Tcw_cur, xy_un, cur_octave, cur_angle, desc,
                                cur_held, cell_start, cell_items, bounds, K4, sf, log_scale_factor, th, orb_dist,
                                check_orientation=True):
    """The reference's own lines (oracle/_ref/libstereoref.so).  Returns (nmatches, new_match, in_range, pred_level): the last two
    are what its own distance test and MapPoint::PredictScale say per point."""
    r = _load_ref()
    r.matchref_search_by_projection_kf.restype = C.c_int
    nP, nC = len(valid), len(desc)
    kp = _kp7(xy_un, cur_angle, cur_octave)
    a = [np.ascontiguousarray(valid, np.uint8), np.ascontiguousarray(world, f32), np.ascontiguousarray(mp_desc, np.uint8),
         np.ascontiguousarray(min_dist, f32), np.ascontiguousarray(max_dist, f32), np.ascontiguousarray(kf_angle, f32),
         np.ascontiguousarray(Tcw_cur, f32).reshape(16)]
    b = [np.ascontiguousarray(desc, np.uint8), np.ascontiguousarray(cur_held, np.uint8), np.ascontiguousarray(cell_start, np.int32),
         np.ascontiguousarray(cell_items, np.int32), np.ascontiguousarray(bounds, f32), np.ascontiguousarray(K4, f32),
         np.ascontiguousarray(sf, f32)]
    out = np.full(nC, -1, np.int32)
    pred = np.zeros(nP, np.int32)
    inr = np.zeros(nP, np.uint8)
    p = lambda x: C.c_void_p(x.ctypes.data)
    n = r.matchref_search_by_projection_kf(C.c_int(nP), *[p(x) for x in a], C.c_int(nC), p(kp), *[p(x) for x in b], C.c_int(len(b[6])),
                                           C.c_float(log_scale_factor), C.c_float(th), C.c_int(int(orb_dist)),
                                           C.c_int(int(check_orientation)), p(out), p(pred), p(inr))
    return int(n), out, inr, pred


# ============================================================================= MonocularInitialization's SearchForInitialization
INT_MAX = 2 ** 31 - 1


def search_for_initialization(xy_un1, octave1, angle1, desc1, prev_matched, xy_un2, octave2, angle2, desc2, cell_start, cell_items,
                              bounds, nnratio=0.9, check_orientation=True, window=100):
    """ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:405-520).
    Returns (nmatches, matches12 int32[n1], prev_matched float32[n1, 2] as updated at :515-517)."""
    n1, n2 = len(desc1), len(desc2)
    d1 = np.ascontiguousarray(desc1).view(np.uint64).reshape(n1, 4) if n1 else np.zeros((0, 4), np.uint64)
    d2 = np.ascontiguousarray(desc2).view(np.uint64).reshape(n2, 4) if n2 else np.zeros((0, 4), np.uint64)
    prev = np.asarray(prev_matched, f32).reshape(n1, 2).copy()
    m12 = np.full(n1, -1, np.int32)
    m21 = np.full(n2, -1, np.int64)
    matched_dist = np.full(n2, INT_MAX, np.int64)
    rot_hist = [[] for _ in range(HISTO_LENGTH)]
    factor = f32(f32(1.0) / f32(HISTO_LENGTH))
    nmatches = 0
    for i1 in range(n1):
        level1 = int(octave1[i1])
        if level1 > 0:
            continue
        cands = features_in_area(prev[i1, 0], prev[i1, 1], f32(window), level1, level1, xy_un2, octave2, cell_start, cell_items, bounds)
        if not cands:
            continue
        best = best2 = INT_MAX
        best_idx = -1
        for i2 in cands:
            d = int(sum(bin(int(a ^ b)).count("1") for a, b in zip(d1[i1], d2[i2])))
            if matched_dist[i2] <= d:
                continue
            if d < best:
                best2, best, best_idx = best, d, i2
            elif d < best2:
                best2 = d
        if best <= TH_LOW:
            if f32(best) < f32(f32(best2) * f32(nnratio)):
                if m21[best_idx] >= 0:
                    m12[m21[best_idx]] = -1
                    nmatches -= 1
                m12[i1] = best_idx
                m21[best_idx] = i1
                matched_dist[best_idx] = best
                nmatches += 1
                if check_orientation:
                    rot = f32(f32(angle1[i1]) - f32(angle2[best_idx]))
                    if rot < 0.0:
                        rot = f32(rot + f32(360.0))
                    b = _round_half_away(f32(rot * factor))
                    if b == HISTO_LENGTH:
                        b = 0
                    rot_hist[b].append(i1)
    if check_orientation:
        ind = compute_three_maxima([len(h) for h in rot_hist])
        for b in range(HISTO_LENGTH):
            if b in ind:
                continue
            for idx1 in rot_hist[b]:
                if m12[idx1] >= 0:
                    m12[idx1] = -1
                    nmatches -= 1
    xy2 = np.asarray(xy_un2, f32).reshape(n2, 2)
    for i1 in range(n1):
        if m12[i1] >= 0:
            prev[i1] = xy2[m12[i1]]
    return nmatches, m12, prev


def ref_search_for_initialization(xy_un1, octave1, angle1, desc1, prev_matched, xy_un2, octave2, angle2, desc2, cell_start,
                                  cell_items, bounds, sf, nnratio=0.9, check_orientation=True, window=100):
    r = _load_ref()
    r.matchref_search_for_initialization.restype = C.c_int
    n1, n2 = len(desc1), len(desc2)
    kp1, kp2 = _kp7(xy_un1, angle1, octave1), _kp7(xy_un2, angle2, octave2)
    prev = np.ascontiguousarray(prev_matched, f32).reshape(n1, 2).copy()
    a = [np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8), np.ascontiguousarray(cell_start, np.int32),
         np.ascontiguousarray(cell_items, np.int32), np.ascontiguousarray(bounds, f32), np.ascontiguousarray(sf, f32)]
    out = np.full(n1, -1, np.int32)
    p = lambda x: C.c_void_p(x.ctypes.data)
    n = r.matchref_search_for_initialization(C.c_int(n1), p(kp1), p(a[0]), p(prev), C.c_int(n2), p(kp2), p(a[1]), p(a[2]), p(a[3]),
                                             p(a[4]), p(a[5]), C.c_int(len(a[5])), C.c_float(nnratio), C.c_int(int(check_orientation)),
                                             C.c_int(int(window)), p(out))
    return int(n), out, prev


# ============================================================================= Frame::isInFrustum (the producer of SearchLocalPoints' inputs)
def predict_scale(max_dist, dist, log_scale_factor, nlevels) -> int:
    """MapPoint::PredictScale(currentDist, Frame*) (src/MapPoint.cc:402-417): ratio = mfMaxDistance / dist in float,
    ceil(logf(ratio) / mfLogScaleFactor), clamped to [0, nlevels - 1].  A non-finite quotient (the reference then converts
    inf / NaN to int, which is undefined; x86 yields INT_MIN -> 0) gives level 0."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        ratio = f32(f32(max_dist) / f32(dist))
        q = f32(_logf(ratio) / f32(log_scale_factor))
    ns = int(math.ceil(float(q))) if np.isfinite(q) else 0
    return 0 if ns < 0 else (nlevels - 1 if ns >= nlevels else ns)


def is_in_frustum(consider, world, normal, min_dist, max_dist, Tcw, K4, mbf, bounds, log_scale_factor, nlevels, cos_limit=0.5):
    """Frame::isInFrustum(MapPoint*, viewingCosLimit) (src/Frame.cc:269-325) over a list of map points, with the pose split of
    Frame::UpdatePoseMatrices (:260-267: mRcw, mtcw, mOw = -mRcw.t() * mtcw).  consider[i] = 0: the point is not handed to
    isInFrustum at all (Tracking::SearchLocalPoints skips points already seen in the frame and bad points, src/Tracking.cc:1169-1172)
    and its mbTrackInView stays false.  The matrix products and the norm are the REAL OpenCV (cv2.gemm, cv2.norm); Mat::dot has no
    Python binding and is restated from OpenCV's published dotProd_32f (modules/core/src/matmul.simd.hpp: for 3 elements no SIMD
    block is reached and the tail adds the float products, exact in double, to a double accumulator in element order).
    Returns (in_view uint8[n], proj float32[n, 3] = (mTrackProjX, mTrackProjY, mTrackProjXR), scale_level int32[n], view_cos float32[n])."""
    import cv2
    n = len(world)
    T = np.asarray(Tcw, f32).reshape(4, 4)
    Rcw, tcw = T[:3, :3], T[:3, 3:4]
    Ow = _gemm(Rcw, tcw, None, -1.0, cv2.GEMM_1_T)                             # (:266)
    fx, fy, cx, cy = [f32(v) for v in K4]
    minx, maxx, miny, maxy = [f32(v) for v in bounds]
    W = np.asarray(world, f32).reshape(n, 3)
    Nn = np.asarray(normal, f32).reshape(n, 3)
    in_view = np.zeros(n, np.uint8)
    proj = np.zeros((n, 3), f32)
    level = np.zeros(n, np.int32)
    vcos = np.zeros(n, f32)
    for i in range(n):
        if consider is not None and not consider[i]:
            continue
        P = W[i].reshape(3, 1)
        Pc = _gemm(Rcw, P, tcw)                                                # mRcw*P+mtcw (:277)
        pcx, pcy, pcz = f32(Pc[0, 0]), f32(Pc[1, 0]), f32(Pc[2, 0])
        if pcz < f32(0.0):                                                     # (:283)
            continue
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            invz = f32(f32(1.0) / pcz)                                         # (:287)
            u = f32(f32(f32(fx * pcx) * invz) + cx)
            v = f32(f32(f32(fy * pcy) * invz) + cy)
        if u < minx or u > maxx or v < miny or v > maxy:                       # (:291-294)
            continue
        with np.errstate(over="ignore", invalid="ignore"):
            maxd = f32(f32(1.2) * f32(max_dist[i]))                            # GetMaxDistanceInvariance (src/MapPoint.cc:379-383)
            mind = f32(f32(0.8) * f32(min_dist[i]))
        PO = (P - Ow).astype(f32)
        dist = f32(cv2.norm(PO))                                               # (:300)
        if dist < mind or dist > maxd:
            continue
        dot = float(np.float64(PO[0, 0]) * np.float64(Nn[i, 0]))               # Mat::dot: float products accumulated in double
        dot = float(np.float64(dot) + np.float64(PO[1, 0]) * np.float64(Nn[i, 1]))
        dot = float(np.float64(dot) + np.float64(PO[2, 0]) * np.float64(Nn[i, 2]))
        with np.errstate(divide="ignore", invalid="ignore"):
            vc = f32(np.float64(dot) / np.float64(dist))                       # (:308) double / float -> double -> float
        if vc < f32(cos_limit):
            continue
        level[i] = predict_scale(max_dist[i], dist, log_scale_factor, nlevels)
        in_view[i] = 1
        proj[i, 0] = u
        with np.errstate(over="ignore", invalid="ignore"):
            proj[i, 2] = f32(u - f32(f32(mbf) * invz))                         # (:319)
        proj[i, 1] = v
        vcos[i] = vc
    return in_view, proj, level, vcos


def ref_is_in_frustum(consider, world, normal, min_dist, max_dist, Tcw, K4, mbf, bounds, log_scale_factor, nlevels, cos_limit=0.5):
    """Same call shape, executed by the reference's own lines of Frame::isInFrustum and MapPoint::PredictScale
    (oracle/_ref/libstereoref.so)."""
    r = _load_ref()
    n = len(world)
    cons = np.ones(n, np.uint8) if consider is None else np.ascontiguousarray(consider, np.uint8)
    a = [cons, np.ascontiguousarray(world, f32), np.ascontiguousarray(normal, f32), np.ascontiguousarray(min_dist, f32),
         np.ascontiguousarray(max_dist, f32), np.ascontiguousarray(Tcw, f32).reshape(16), np.ascontiguousarray(K4, f32),
         np.ascontiguousarray(bounds, f32)]
    in_view = np.zeros(n, np.uint8)
    proj = np.zeros((n, 3), f32)
    level = np.zeros(n, np.int32)
    vcos = np.zeros(n, f32)
    p = lambda x: C.c_void_p(x.ctypes.data)
    r.matchref_is_in_frustum.restype = C.c_int
    r.matchref_is_in_frustum(C.c_int(n), *[p(x) for x in a], C.c_float(mbf), C.c_float(log_scale_factor), C.c_int(nlevels),
                             C.c_float(cos_limit), p(in_view), p(proj), p(level), p(vcos))
    return in_view, proj, level, vcos
