"""CPU restatement of Frame::ComputeStereoMatches (reference src/Frame.cc:466-640).

TEST INFRASTRUCTURE: the checker of ``orbx_stereo_match``; only tests/ (and bench.py's CPU legs) may
import it.  Pinned against the reference's OWN lines compiled here (oracle/build_stereo_ref.sh ->
oracle/_ref/libstereoref.so, bound below as ``ref_compute_stereo_matches``) by tests/test_stereo_oracle.py.

Arithmetic notes (all float32, nothing contracted):
  * row table (:476-496): right keypoint iR is listed in rows floor(y - r) .. ceil(y + r), r = 2 * mvScaleFactors[octave];
    a left keypoint looks at row (size_t)vL, candidates in iR order, strict `<` keeps the first best (:522-546);
  * DescriptorDistance (src/ORBmatcher.cc:1647-1663) is the Hamming distance of the 256-bit descriptors;
  * the 11x11 SAD (:552-596) subtracts each window's own centre; every value is an integer <= 61 710, so the float
    L1 norm is exact; `dist < bestDist` is strict, the first best incR wins;
  * parabola (:602-610), re-scaling (:613), disparity clamp (:617-622: `bestuR = uL - 0.01` is evaluated in double);
  * final filter (:629-640): median = sorted SAD list [size / 2], thDist = 1.5f * 1.4f * median, entries with
    dist >= thDist are reset to -1.  An empty list is undefined behaviour in the reference; here nothing is filtered.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

f32 = np.float32
TH_HIGH, TH_LOW = 100, 50          # src/ORBmatcher.cc:37-38
_HERE = os.path.dirname(os.path.abspath(__file__))


def _hamming(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a: (32,) u8, b: (m, 32) u8 -> (m,) int"""
    return np.unpackbits(np.bitwise_xor(b, a[None, :]), axis=1).sum(axis=1).astype(np.int64)


def _round(x) -> int:
    """C round() of a non-negative float32 (exact in double)."""
    return int(math.floor(float(x) + 0.5))


def compute_stereo_matches(kpL, descL, pyrL, kpR, descR, pyrR, sf, inv_sf, mbf, mb, edge=19):
    """kp*: structured keypoint arrays (oracle KP_DTYPE); pyr*: PADDED planes per level (border `edge`);
    returns (mvuRight, mvDepth, best SAD per left keypoint or -1)."""
    N, Nr = len(kpL), len(kpR)
    sf = np.asarray(sf, f32)
    inv_sf = np.asarray(inv_sf, f32)
    mbf, mb = f32(mbf), f32(mb)
    uRight = np.full(N, -1.0, f32)
    depth = np.full(N, -1.0, f32)
    sad = np.full(N, -1, np.int64)
    thOrbDist = (TH_HIGH + TH_LOW) // 2
    nRows = pyrL[0].shape[0] - 2 * edge
    # row table (:476-496), kept as per-keypoint row intervals
    yR = kpR["y"].astype(f32)
    rR = (f32(2.0) * sf[kpR["octave"]]).astype(f32)
    maxr = np.ceil((yR + rR).astype(f32)).astype(np.int64)
    minr = np.floor((yR - rR).astype(f32)).astype(np.int64)
    xR = kpR["x"].astype(f32)
    octR = kpR["octave"].astype(np.int64)
    maxD = f32(mbf / mb)                                      # (:499-501) minZ = mb, maxD = mbf / minZ
    minD = f32(0)
    accepted = []
    for iL in range(N):
        levelL = int(kpL["octave"][iL])
        vL, uL = f32(kpL["y"][iL]), f32(kpL["x"][iL])
        row = int(vL)                                         # vRowIndices[vL]: float -> size_t
        assert 0 <= row < nRows
        cand = (minr <= row) & (row <= maxr)
        if not cand.any():
            continue
        minU, maxU = f32(uL - maxD), f32(uL - minD)
        if maxU < 0:
            continue
        ok = cand & (octR >= levelL - 1) & (octR <= levelL + 1) & (xR >= minU) & (xR <= maxU)
        idx = np.nonzero(ok)[0]
        bestDist, bestIdxR = TH_HIGH, 0
        if len(idx):
            d = _hamming(descL[iL], descR[idx])
            j = int(np.argmin(d))                             # first minimum = first in iR order
            if d[j] < bestDist:
                bestDist, bestIdxR = int(d[j]), int(idx[j])
        if bestDist >= thOrbDist:
            continue
        uR0 = f32(xR[bestIdxR])
        scaleFactor = inv_sf[levelL]
        scaleduL = _round(f32(uL * scaleFactor))              # (:556-558) round(): half away from zero
        scaledvL = _round(f32(vL * scaleFactor))
        scaleduR0 = _round(f32(uR0 * scaleFactor))
        w = L = 5
        PL = pyrL[levelL][edge:-edge, edge:-edge] if edge else pyrL[levelL]
        PR = pyrR[levelL][edge:-edge, edge:-edge] if edge else pyrR[levelL]
        IL = PL[scaledvL - w:scaledvL + w + 1, scaleduL - w:scaleduL + w + 1].astype(np.int64)
        IL = IL - IL[w, w]
        iniu, endu = scaleduR0 + L - w, scaleduR0 + L + w + 1
        if iniu < 0 or endu >= PR.shape[1]:
            continue
        best, bestinc = 2 ** 31 - 1, 0
        vd = np.zeros(2 * L + 1, np.int64)
        for inc in range(-L, L + 1):
            c0 = scaleduR0 + inc - w
            assert c0 >= 0, "negative colRange start throws in OpenCV"
            IR = PR[scaledvL - w:scaledvL + w + 1, c0:c0 + 2 * w + 1].astype(np.int64)
            IR = IR - IR[w, w]
            dist = int(np.abs(IL - IR).sum())
            if dist < best:
                best, bestinc = dist, inc
            vd[L + inc] = dist
        if bestinc == -L or bestinc == L:
            continue
        d1, d2, d3 = f32(vd[L + bestinc - 1]), f32(vd[L + bestinc]), f32(vd[L + bestinc + 1])
        with np.errstate(divide="ignore", invalid="ignore"):
            deltaR = f32(f32(d1 - d3) / f32(f32(2.0) * f32(f32(d1 + d3) - f32(f32(2.0) * d2))))
        if deltaR < -1 or deltaR > 1:
            continue
        bestuR = f32(sf[levelL] * f32(f32(f32(scaleduR0) + f32(bestinc)) + deltaR))
        disparity = f32(uL - bestuR)
        if disparity >= minD and disparity < maxD:
            if disparity <= 0:
                disparity = f32(0.01)
                bestuR = f32(np.float64(uL) - 0.01)
            depth[iL] = f32(mbf / disparity)
            uRight[iL] = bestuR
            sad[iL] = best
            accepted.append((best, iL))
    if accepted:
        accepted.sort()
        median = f32(accepted[len(accepted) // 2][0])
        thDist = f32(f32(f32(1.5) * f32(1.4)) * median)
        for dist, iL in reversed(accepted):
            if f32(dist) < thDist:
                break
            uRight[iL] = -1
            depth[iL] = -1
    return uRight, depth, sad


# ----------------------------------------------------------------------------- the reference's own lines
_ref = None


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libstereoref.so"))


def ref_build() -> bool:
    import subprocess
    if not os.path.isdir("/root/reference/src"):
        return ref_available()
    subprocess.run(["sh", os.path.join(_HERE, "build_stereo_ref.sh")], check=True, stdout=subprocess.DEVNULL)
    return True


def ref_compute_stereo_matches(kpL, descL, pyrL, kpR, descR, pyrR, sf, inv_sf, mbf, mb, edge=19):
    """Same call shape, executed by the reference's own lines (oracle/_ref/libstereoref.so)."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libstereoref.so"))
        _ref.stereoref_match.restype = C.c_int
    nl = len(pyrL)
    kl = np.ascontiguousarray(kpL).view(np.float32).reshape(-1, 7) if len(kpL) else np.zeros((0, 7), f32)
    kr = np.ascontiguousarray(kpR).view(np.float32).reshape(-1, 7) if len(kpR) else np.zeros((0, 7), f32)
    dl, dr = np.ascontiguousarray(descL), np.ascontiguousarray(descR)
    PL = [np.ascontiguousarray(p) for p in pyrL]
    PR = [np.ascontiguousarray(p) for p in pyrR]
    ptr = lambda a, off: a.ctypes.data + off
    pl = (C.c_void_p * nl)(*[ptr(p, edge * p.shape[1] + edge) for p in PL])
    pr = (C.c_void_p * nl)(*[ptr(p, edge * p.shape[1] + edge) for p in PR])
    lw = (C.c_int * nl)(*[p.shape[1] - 2 * edge for p in PL])
    lh = (C.c_int * nl)(*[p.shape[0] - 2 * edge for p in PL])
    sl = (C.c_size_t * nl)(*[p.shape[1] for p in PL])
    sr = (C.c_size_t * nl)(*[p.shape[1] for p in PR])
    sfa = np.ascontiguousarray(sf, f32)
    isf = np.ascontiguousarray(inv_sf, f32)
    u = np.zeros(len(kpL), f32)
    d = np.zeros(len(kpL), f32)
    _ref.stereoref_match(C.c_int(len(kpL)), C.c_void_p(kl.ctypes.data), C.c_void_p(dl.ctypes.data), C.c_int(len(kpR)),
                         C.c_void_p(kr.ctypes.data), C.c_void_p(dr.ctypes.data), C.c_int(nl), pl, pr, lw, lh, sl, sr,
                         C.c_void_p(sfa.ctypes.data), C.c_void_p(isf.ctypes.data), C.c_float(mbf), C.c_float(mb),
                         C.c_void_p(u.ctypes.data), C.c_void_p(d.ctypes.data))
    return u, d
