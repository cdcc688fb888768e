"""Python mirror of ORB_SLAM2::ORBextractor (reference include/ORBextractor.h:45-111) on top of
the C ABI (include/orbx.h).  Same constructor arguments, same getters, `mvImagePyramid`, and a
call operator that returns (keypoints, descriptors) instead of filling output arguments.

All compute happens in liborbx.so's CUDA kernels; if the library is missing, importing this
module works but constructing an extractor raises."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import KP_DTYPE, OrbxConfig, OrbxError, OrbxResult, check, lib


class ORBextractor:
    HARRIS_SCORE = 0     # include/ORBextractor.h:49 (unused by the reference as well)
    FAST_SCORE = 1

    def __init__(self, nfeatures: int, scaleFactor: float, nlevels: int, iniThFAST: int, minThFAST: int,
                 device: int = 0, max_batch: int = 1, download_pyramid: bool = True, candidate_divisor: int = 0,
                 device_chunks: int = 0):
        self._L = lib()
        self._h = C.c_void_p()
        cfg = OrbxConfig(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device, max_batch,
                         1 if download_pyramid else 0, candidate_divisor, device_chunks)
        check(self._L.orbx_create(C.byref(cfg), C.byref(self._h)))
        self.nfeatures, self.nlevels, self.max_batch = nfeatures, nlevels, max_batch
        self.download_pyramid = bool(download_pyramid)
        self.mvImagePyramid: List[np.ndarray] = [np.zeros((0, 0), np.uint8) for _ in range(nlevels)]
        self._pyramids: List[List[np.ndarray]] = []
        self._last_n = 0
        self._last_shape: Optional[Tuple[int, int]] = None

    # ---------------------------------------------------------------- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.orbx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- getters (ORBextractor.h:63-83)
    def GetLevels(self) -> int:
        return self._L.orbx_get_levels(self._h)

    def GetScaleFactor(self) -> float:
        return float(self._L.orbx_get_scale_factor(self._h))

    def _tables(self):
        p = [C.POINTER(C.c_float)() for _ in range(4)]
        check(self._L.orbx_scale_tables(self._h, *[C.byref(x) for x in p]))
        return [np.ctypeslib.as_array(x, shape=(self.nlevels,)).copy() for x in p]

    def GetScaleFactors(self): return self._tables()[0]
    def GetInverseScaleFactors(self): return self._tables()[1]
    def GetScaleSigmaSquares(self): return self._tables()[2]
    def GetInverseScaleSigmaSquares(self): return self._tables()[3]

    def level_quotas(self):
        q = (C.c_int * self.nlevels)()
        u = (C.c_int * 16)()
        check(self._L.orbx_level_quotas(self._h, q, u))
        return list(q), list(u)

    def level_sizes(self, width: int, height: int):
        w = (C.c_int * self.nlevels)()
        h = (C.c_int * self.nlevels)()
        check(self._L.orbx_level_sizes(self._h, width, height, w, h))
        return list(zip(w, h))

    def algorithmic_bytes(self, width: int, height: int) -> int:
        return int(self._L.orbx_algorithmic_bytes(self._h, width, height))

    # ---------------------------------------------------------------- operator()
    def __call__(self, image: np.ndarray, mask=None):
        """ORBextractor::operator() (src/ORBextractor.cc:1043).  mask is ignored, as in the reference.
        Returns None for an empty image (the reference returns without touching its outputs)."""
        if image is None or image.size == 0:
            return None
        return self.extract_batch([image])[0]

    def extract_batch(self, images: Sequence[np.ndarray], copy: bool = True):
        """n frames of identical shape in one launch sequence.  With copy=False the returned arrays are
        views of the library's pinned result buffers, valid until the next call (the C ABI's contract)."""
        n = len(images)
        if n == 0:
            return []
        h, w = images[0].shape
        ptrs = (C.c_void_p * n)()
        strides = (C.c_size_t * n)()
        keep = []
        for i, im in enumerate(images):
            if im.dtype != np.uint8 or im.ndim != 2 or im.shape != (h, w):
                raise ValueError("images must be 2-D uint8 arrays of identical shape (CV_8UC1)")
            if im.strides[1] != 1:
                im = np.ascontiguousarray(im)
            keep.append(im)
            ptrs[i] = im.__array_interface__["data"][0]
            strides[i] = im.strides[0]
        res = (OrbxResult * n)()
        check(self._L.orbx_extract_batch(self._h, n, ptrs, w, h, strides, res), self._h)
        self._last_n, self._last_shape = n, (h, w)
        out = self._copy_results(res, n, copy)
        self._refresh_pyramids(n)
        return out

    def extract_batch_color(self, images: Sequence[np.ndarray], fmt: int):
        """Colour frames (H x W x 3 or 4, uint8; fmt = _capi.BGR8 / RGB8 / BGRA8 / RGBA8): cvtColor -> GRAY on the device
        (reference src/Tracking.cc:172-255), then the normal path."""
        n = len(images)
        if n == 0:
            return []
        h, w, ch = images[0].shape
        ptrs = (C.c_void_p * n)()
        strides = (C.c_size_t * n)()
        keep = []
        for i, im in enumerate(images):
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape != (h, w, ch):
                raise ValueError("images must be H x W x C uint8 arrays of identical shape")
            if im.strides[2] != 1 or im.strides[1] != ch:
                im = np.ascontiguousarray(im)
            keep.append(im)
            ptrs[i] = im.__array_interface__["data"][0]
            strides[i] = im.strides[0]
        res = (OrbxResult * n)()
        check(self._L.orbx_extract_batch_color(self._h, n, ptrs, w, h, strides, fmt, res), self._h)
        self._last_n, self._last_shape = n, (h, w)
        out = self._copy_results(res, n, True)
        self._refresh_pyramids(n)
        return out

    @staticmethod
    def _copy_results(res, n: int, copy: bool = True):
        """Copies the per-frame results out of the library's pinned buffers (two numpy views over the
        whole batch, then one slice copy per frame)."""
        if n == 0:
            return []
        stride_kp = (res[1].kps - res[0].kps) if n > 1 else max(res[0].n, 1) * 28       # bytes between frames
        cap = stride_kp // 28
        kp_all = np.ctypeslib.as_array(C.cast(res[0].kps, C.POINTER(C.c_uint8)), shape=(n * cap * 28,)).view(KP_DTYPE).reshape(n, cap)
        ds_all = np.ctypeslib.as_array(C.cast(res[0].desc, C.POINTER(C.c_uint8)), shape=(n, cap, 32))
        out = []
        for f in range(n):
            k = res[f].n          # 0: descriptors released (src/ORBextractor.cc:1064-1065)
            out.append((kp_all[f, :k].copy(), ds_all[f, :k].copy()) if copy else (kp_all[f, :k], ds_all[f, :k]))
        return out

    def _refresh_pyramids(self, n: int):
        if not self.download_pyramid:
            return
        self.mvImagePyramid = [self._pyramid_view(0, l) for l in range(self.nlevels)]

    def _pyramid_view(self, frame: int, level: int, padded: bool = False) -> np.ndarray:
        p = C.c_void_p()
        w, h, step = C.c_int(), C.c_int(), C.c_size_t()
        check(self._L.orbx_pyramid_level(self._h, frame, level, C.byref(p), C.byref(w), C.byref(h), C.byref(step)))
        E = 19
        base = p.value - E * step.value - E
        rows, cols = h.value + 2 * E, w.value + 2 * E
        buf = (C.c_uint8 * (rows * step.value)).from_address(base)
        plane = np.ctypeslib.as_array(buf).reshape(rows, step.value)[:, :cols]
        return plane if padded else plane[E:-E, E:-E]

    def pyramid(self, frame: int = 0, padded: bool = True):
        """Host views (valid until the next call) of the padded planes of `frame`."""
        return [self._pyramid_view(frame, l, padded) for l in range(self.nlevels)]

    # ---------------------------------------------------------------- device-resident path
    def extract_device(self, d_ptr: int, n: int, width: int, height: int, pitch: int, frame_stride: int):
        """Enqueue the path on n images already in HBM (roofline timing); no host sync."""
        check(self._L.orbx_extract_device(self._h, n, C.c_void_p(d_ptr), width, height, pitch, frame_stride), self._h)
        self._last_n, self._last_shape = n, (height, width)

    def fetch_results(self, n: int):
        res = (OrbxResult * n)()
        check(self._L.orbx_fetch_results(self._h, n, res), self._h)
        out = self._copy_results(res, n)
        self._refresh_pyramids(n)
        return out

    def fast_stats(self, frame: int = 0):
        """(FAST corners handed to the quadtree per level, cells re-run at minThFAST per level) of a fetched frame."""
        c = (C.c_int * self.nlevels)()
        r = (C.c_int * self.nlevels)()
        check(self._L.orbx_fast_stats(self._h, frame, c, r), self._h)
        return list(c), list(r)

    def synchronize(self):
        check(self._L.orbx_synchronize(self._h), self._h)

    @property
    def stream(self) -> int:
        return int(self._L.orbx_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._L.orbx_launch_count(self._h))

    def stage_timing(self, enable: bool):
        check(self._L.orbx_stage_timing_enable(self._h, 1 if enable else 0), self._h)

    def stage_times(self):
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        ln = (C.c_int * 16)()
        k = self._L.orbx_stage_timing_read(self._h, 16, names, ms, ln)
        return [(names[i].decode(), float(ms[i]), int(ln[i])) for i in range(k)]

    # ---------------------------------------------------------------- Frame::ComputeStereoMatches (src/Frame.cc:466-640)
    def stereo_match(self, right: "ORBextractor", mbf: float, mb: float, left_frames=None, right_frames=None,
                     npairs: Optional[int] = None):
        """Matches frames of the last extract on `self` (left eye) against frames of the last extract on `right`
        (which may be `self` when one batch holds both eyes).  Returns [(mvuRight, mvDepth)] per pair, one float per
        left keypoint, -1 where the reference leaves -1."""
        if npairs is None:
            npairs = len(left_frames) if left_frames is not None else min(self._last_n, right._last_n)
        lf = (C.c_int * npairs)(*left_frames) if left_frames is not None else None
        rf = (C.c_int * npairs)(*right_frames) if right_frames is not None else None
        res = (_capi.OrbxStereoResult * npairs)()
        check(self._L.orbx_stereo_match(self._h, right._h, npairs, lf, rf, mbf, mb, res), self._h)
        out = []
        for i in range(npairs):
            n = res[i].n
            u = np.ctypeslib.as_array(C.cast(res[i].u_right, C.POINTER(C.c_float)), shape=(max(n, 1),))[:n].copy()
            d = np.ctypeslib.as_array(C.cast(res[i].depth, C.POINTER(C.c_float)), shape=(max(n, 1),))[:n].copy()
            out.append((u, d))
        return out

    def stereo_match_device(self, right: "ORBextractor", mbf: float, mb: float, left_frames, right_frames):
        """Enqueue only (device-side timing)."""
        n = len(left_frames)
        check(self._L.orbx_stereo_match_device(self._h, right._h, n, (C.c_int * n)(*left_frames), (C.c_int * n)(*right_frames),
                                               mbf, mb), self._h)

    # ---------------------------------------------------------------- Frame::UndistortKeyPoints + AssignFeaturesToGrid
    def undistort_grid(self, K4, dist, frames=None):
        """(src/Frame.cc:404-434, :230-245) on the keypoints of the last extract.  K4 = (fx, fy, cx, cy); dist = (k1, k2,
        p1, p2[, k3]).  Returns per frame (xy_un (n, 2), cell_start (64*48+1,), cell_items, bounds (4,))."""
        n = len(frames) if frames is not None else self._last_n
        fr_ = (C.c_int * n)(*frames) if frames is not None else None
        k = (C.c_float * 4)(*K4)
        d = (C.c_float * len(dist))(*dist)
        res = (_capi.OrbxGridResult * n)()
        check(self._L.orbx_undistort_grid(self._h, n, fr_, k, d, len(dist), res), self._h)
        out = []
        for r in res:
            xy = np.ctypeslib.as_array(C.cast(r.xy_un, C.POINTER(C.c_float)), shape=(max(r.n, 1) * 2,))[:2 * r.n].reshape(-1, 2).copy()
            st = np.ctypeslib.as_array(C.cast(r.cell_start, C.POINTER(C.c_int32)), shape=(64 * 48 + 1,)).copy()
            it = np.ctypeslib.as_array(C.cast(r.cell_items, C.POINTER(C.c_int32)), shape=(max(r.n_in_grid, 1),))[:r.n_in_grid].copy()
            out.append((xy, st, it, np.array(list(r.bounds), np.float32)))
        return out

    # ---------------------------------------------------------------- ORBmatcher::SearchByProjection(Frame&, const Frame&)
    def _projection_queries(self, queries):
        """queries: dicts with cur_frame, world (n, 3), mp_desc (n, 32), mp_obs (n,), outlier (n,) or None, last_octave,
        last_angle, Tcw_cur, Tcw_last.  Returns (ctypes array, keep-alive list)."""
        qs = (_capi.OrbxProjectionQuery * len(queries))()
        keep = []
        for q, d in zip(qs, queries):
            a = dict(world=np.ascontiguousarray(d["world"], np.float32), mp_desc=np.ascontiguousarray(d["mp_desc"], np.uint8),
                     mp_obs=np.ascontiguousarray(d["mp_obs"], np.int32), octave=np.ascontiguousarray(d["last_octave"], np.int32),
                     angle=np.ascontiguousarray(d["last_angle"], np.float32))
            out = None if d.get("outlier") is None else np.ascontiguousarray(d["outlier"], np.uint8)
            keep.append((a, out))
            q.cur_frame, q.n_last = int(d.get("cur_frame", 0)), len(a["mp_obs"])
            q.world_pos, q.mp_desc, q.mp_obs = a["world"].ctypes.data, a["mp_desc"].ctypes.data, a["mp_obs"].ctypes.data
            q.outlier = None if out is None else out.ctypes.data
            q.octave, q.angle = a["octave"].ctypes.data, a["angle"].ctypes.data
            q.Tcw_cur = (C.c_float * 16)(*np.asarray(d["Tcw_cur"], np.float32).reshape(16))
            q.Tcw_last = (C.c_float * 16)(*np.asarray(d["Tcw_last"], np.float32).reshape(16))
        return qs, keep

    def search_by_projection(self, queries, K4, mbf: float, mb: float, th: float, mono: bool, check_orientation: bool = True,
                             use_stereo: bool = False):
        """(src/ORBmatcher.cc:1328-1470) for each query against frame `cur_frame` of the last extract (orbx_undistort_grid
        must have run with the same K4).  Returns [(nmatches, match int32[N], rounds)]."""
        qs, keep = self._projection_queries(queries)
        res = (_capi.OrbxProjectionResult * len(queries))()
        check(self._L.orbx_search_by_projection(self._h, len(queries), qs, (C.c_float * 4)(*K4), mbf, mb, th, int(mono),
                                                int(check_orientation), int(use_stereo), res), self._h)
        out = []
        for r in res:
            m = np.ctypeslib.as_array(C.cast(r.match, C.POINTER(C.c_int32)), shape=(max(r.n, 1),))[:r.n].copy()
            out.append((r.nmatches, m, r.rounds))
        del keep
        return out

    def search_by_projection_device(self, prepared, K4, mbf, mb, th, mono, check_orientation=True, use_stereo=False):
        """Stage + enqueue only (device-side timing); `prepared` = self._projection_queries(queries)."""
        qs, _ = prepared
        check(self._L.orbx_search_by_projection_device(self._h, len(qs), qs, (C.c_float * 4)(*K4), mbf, mb, th, int(mono),
                                                       int(check_orientation), int(use_stereo)), self._h)

    # ---------------------------------------------------------------- ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th)
    def _local_queries(self, queries):
        qs = (_capi.OrbxLocalPointsQuery * len(queries))()
        keep = []
        for q, d in zip(qs, queries):
            a = dict(in_view=np.ascontiguousarray(d["in_view"], np.uint8),
                     w3=np.ascontiguousarray(np.stack([d["proj_x"], d["proj_y"], d["proj_xr"]], axis=1), np.float32),
                     lvl=np.ascontiguousarray(d["scale_level"], np.int32), vc=np.ascontiguousarray(d["view_cos"], np.float32),
                     desc=np.ascontiguousarray(d["mp_desc"], np.uint8), obs=np.ascontiguousarray(d["mp_obs"], np.int32))
            co = None if d.get("cur_obs") is None else np.ascontiguousarray(d["cur_obs"], np.int32)
            keep.append((a, co))
            q.cur_frame, q.n_points = int(d.get("cur_frame", 0)), len(a["obs"])
            q.in_view, q.proj_xy_xr, q.scale_level = a["in_view"].ctypes.data, a["w3"].ctypes.data, a["lvl"].ctypes.data
            q.view_cos, q.mp_desc, q.mp_obs = a["vc"].ctypes.data, a["desc"].ctypes.data, a["obs"].ctypes.data
            q.cur_obs = None if co is None else co.ctypes.data
        return qs, keep

    def search_local_points(self, queries, th: float, nnratio: float = 0.8, use_stereo: bool = False):
        """(src/ORBmatcher.cc:45-129) for each query: dicts with cur_frame, in_view, proj_x, proj_y, proj_xr, scale_level,
        view_cos, mp_desc, mp_obs, cur_obs (or None).  Returns [(nmatches, new_match int32[N], rounds)]."""
        qs, keep = self._local_queries(queries)
        res = (_capi.OrbxProjectionResult * len(queries))()
        check(self._L.orbx_search_local_points(self._h, len(queries), qs, th, nnratio, int(use_stereo), res), self._h)
        out = []
        for r in res:
            m = np.ctypeslib.as_array(C.cast(r.match, C.POINTER(C.c_int32)), shape=(max(r.n, 1),))[:r.n].copy()
            out.append((r.nmatches, m, r.rounds))
        del keep
        return out

    # ---------------------------------------------------------------- Frame::isInFrustum over the local map
    def is_in_frustum(self, queries, K4, mbf: float, bounds, log_scale_factor: float, cos_limit: float = 0.5):
        """(src/Frame.cc:269-325 with MapPoint::PredictScale, src/MapPoint.cc:402-417; Tracking::SearchLocalPoints) for each query:
        dicts with consider (uint8 or None), world (n x 3), normal (n x 3), min_dist, max_dist (mfMinDistance / mfMaxDistance), Tcw.
        Returns [(in_view uint8[n], proj float32[n, 3] = (mTrackProjX, mTrackProjY, mTrackProjXR), scale_level int32[n],
        view_cos float32[n])] -- the inputs of search_local_points."""
        qs = (_capi.OrbxFrustumQuery * len(queries))()
        keep = []
        for q, d in zip(qs, queries):
            a = dict(world=np.ascontiguousarray(d["world"], np.float32).reshape(-1, 3), normal=np.ascontiguousarray(d["normal"], np.float32).reshape(-1, 3),
                     mn=np.ascontiguousarray(d["min_dist"], np.float32), mx=np.ascontiguousarray(d["max_dist"], np.float32))
            co = None if d.get("consider") is None else np.ascontiguousarray(d["consider"], np.uint8)
            keep.append((a, co))
            q.n_points = len(a["mn"])
            q.consider = None if co is None else co.ctypes.data
            q.world_pos, q.normal, q.min_dist, q.max_dist = a["world"].ctypes.data, a["normal"].ctypes.data, a["mn"].ctypes.data, a["mx"].ctypes.data
            q.Tcw = (C.c_float * 16)(*np.asarray(d["Tcw"], np.float32).reshape(16).tolist())
        res = (_capi.OrbxFrustumResult * len(queries))()
        k4 = (C.c_float * 4)(*[float(v) for v in K4])
        bd = (C.c_float * 4)(*[float(v) for v in bounds])
        check(self._L.orbx_is_in_frustum(self._h, len(queries), qs, k4, mbf, bd, log_scale_factor, cos_limit, res), self._h)
        out = []
        for r in res:
            n = r.n
            if n == 0:
                out.append((np.zeros(0, np.uint8), np.zeros((0, 3), np.float32), np.zeros(0, np.int32), np.zeros(0, np.float32)))
                continue
            iv = np.ctypeslib.as_array(C.cast(r.in_view, C.POINTER(C.c_uint8)), shape=(n,)).copy()
            pj = np.ctypeslib.as_array(C.cast(r.proj_xy_xr, C.POINTER(C.c_float)), shape=(n, 3)).copy()
            lv = np.ctypeslib.as_array(C.cast(r.scale_level, C.POINTER(C.c_int32)), shape=(n,)).copy()
            vc = np.ctypeslib.as_array(C.cast(r.view_cos, C.POINTER(C.c_float)), shape=(n,)).copy()
            assert int(iv.sum()) == r.n_in_view
            out.append((iv, pj, lv, vc))
        del keep
        return out

    # ---------------------------------------------------------------- ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist)
    def search_by_projection_kf(self, queries, K4, th: float, orb_dist: int, check_orientation: bool = True):
        """(src/ORBmatcher.cc:1472-1599, Tracking::Relocalization) for each query: dicts with cur_frame, search (uint8: good, not yet
        found, inside the distance-invariance range), world (n x 3), pred_level (MapPoint::PredictScale), mp_desc, kf_angle, Tcw_cur,
        cur_held (or None).  Returns [(nmatches, new_match int32[N], rounds)]."""
        qs = (_capi.OrbxKeyframeProjectionQuery * len(queries))()
        keep = []
        for q, d in zip(qs, queries):
            a = dict(s=np.ascontiguousarray(d["search"], np.uint8), w=np.ascontiguousarray(d["world"], np.float32),
                     lvl=np.ascontiguousarray(d["pred_level"], np.int32), desc=np.ascontiguousarray(d["mp_desc"], np.uint8),
                     ang=np.ascontiguousarray(d["kf_angle"], np.float32))
            ch = None if d.get("cur_held") is None else np.ascontiguousarray(d["cur_held"], np.int32)
            keep.append((a, ch))
            q.cur_frame, q.n_points = int(d.get("cur_frame", 0)), len(a["s"])
            q.search, q.world_pos, q.pred_level = a["s"].ctypes.data, a["w"].ctypes.data, a["lvl"].ctypes.data
            q.mp_desc, q.kf_angle = a["desc"].ctypes.data, a["ang"].ctypes.data
            q.cur_held = None if ch is None else ch.ctypes.data
            q.Tcw_cur = (C.c_float * 16)(*np.asarray(d["Tcw_cur"], np.float32).reshape(16))
        res = (_capi.OrbxProjectionResult * len(queries))()
        check(self._L.orbx_search_by_projection_kf(self._h, len(queries), qs, (C.c_float * 4)(*K4), th, int(orb_dist),
                                                   int(check_orientation), res), self._h)
        out = []
        for r in res:
            m = np.ctypeslib.as_array(C.cast(r.match, C.POINTER(C.c_int32)), shape=(max(r.n, 1),))[:r.n].copy()
            out.append((r.nmatches, m, r.rounds))
        del keep
        return out

    # ---------------------------------------------------------------- ORBmatcher::SearchForInitialization
    def search_for_initialization(self, queries, nnratio: float = 0.9, check_orientation: bool = True, window: int = 100):
        """(src/ORBmatcher.cc:405-520, Tracking::MonocularInitialization) for each query: dicts with cur_frame (F2, on the device),
        octave1, angle1, desc1 (F1's undistorted keypoints), prev_matched (n1 x 2).  Returns [(nmatches, matches12 int32[n1],
        prev_matched float32[n1, 2] as updated at :515-517)]."""
        qs = (_capi.OrbxInitializationQuery * len(queries))()
        keep = []
        for q, d in zip(qs, queries):
            a = dict(o=np.ascontiguousarray(d["octave1"], np.int32), a=np.ascontiguousarray(d["angle1"], np.float32),
                     d=np.ascontiguousarray(d["desc1"], np.uint8), p=np.ascontiguousarray(d["prev_matched"], np.float32))
            keep.append(a)
            q.cur_frame, q.n1 = int(d.get("cur_frame", 0)), len(a["o"])
            q.octave1, q.angle1, q.desc1, q.prev_matched = a["o"].ctypes.data, a["a"].ctypes.data, a["d"].ctypes.data, a["p"].ctypes.data
        res = (_capi.OrbxInitializationResult * len(queries))()
        check(self._L.orbx_search_for_initialization(self._h, len(queries), qs, nnratio, int(check_orientation), int(window), res), self._h)
        out = []
        for r in res:
            m = np.ctypeslib.as_array(C.cast(r.matches12, C.POINTER(C.c_int32)), shape=(max(r.n1, 1),))[:r.n1].copy()
            pm = np.ctypeslib.as_array(C.cast(r.prev_matched, C.POINTER(C.c_float)), shape=(max(r.n1, 1) * 2,))[:2 * r.n1].copy()
            out.append((r.nmatches, m, pm.reshape(-1, 2)))
        del keep
        return out

    # ---------------------------------------------------------------- Frame::ComputeBoW
    def compute_bow(self, voc: "Vocabulary", frames=None, levelsup: int = 4):
        """(src/Frame.cc:395-402) on the descriptors of the last extract.  Returns per frame (word_ids uint32[], word_values
        float64[], fv_nodes uint32[], fv_features uint32[])."""
        n = len(frames) if frames is not None else self._last_n
        fr_ = (C.c_int * n)(*frames) if frames is not None else None
        res = (_capi.OrbxBowResult * n)()
        check(self._L.orbx_compute_bow(self._h, voc._v, n, fr_, levelsup, res), self._h)
        out = []
        for r in res:
            arr = lambda p, t, k: np.ctypeslib.as_array(C.cast(p, C.POINTER(t)), shape=(max(k, 1),))[:k].copy()
            out.append((arr(r.word_ids, C.c_uint32, r.n_words), arr(r.word_values, C.c_double, r.n_words),
                        arr(r.fv_nodes, C.c_uint32, r.n_features), arr(r.fv_features, C.c_uint32, r.n_features)))
        return out

    def compute_bow_device(self, voc: "Vocabulary", frames, levelsup: int = 4):
        """Enqueue only (device-side timing)."""
        n = len(frames)
        check(self._L.orbx_compute_bow_device(self._h, voc._v, n, (C.c_int * n)(*frames), levelsup), self._h)

    # ---------------------------------------------------------------- ORBmatcher::SearchByBoW(KeyFrame*, Frame&)
    def _bow_queries(self, queries):
        qs = (_capi.OrbxBowMatchQuery * len(queries))()
        keep = []
        for q, d in zip(qs, queries):
            a = dict(desc=np.ascontiguousarray(d["kf_desc"], np.uint8), valid=np.ascontiguousarray(d["kf_valid"], np.uint8),
                     angle=np.ascontiguousarray(d["kf_angle"], np.float32), fn=np.ascontiguousarray(d["kf_fv_nodes"], np.uint32),
                     ff=np.ascontiguousarray(d["kf_fv_features"], np.uint32))
            keep.append(a)
            q.cur_frame, q.n_kf, q.n_kf_fv = int(d.get("cur_frame", 0)), len(a["valid"]), len(a["fn"])
            q.kf_desc, q.kf_valid, q.kf_angle = a["desc"].ctypes.data, a["valid"].ctypes.data, a["angle"].ctypes.data
            q.kf_fv_nodes, q.kf_fv_features = a["fn"].ctypes.data, a["ff"].ctypes.data
        return qs, keep

    def search_by_bow(self, queries, nnratio: float = 0.7, check_orientation: bool = True):
        """(src/ORBmatcher.cc:159-288) for each query (dicts with cur_frame, kf_desc, kf_valid, kf_angle, kf_fv_nodes,
        kf_fv_features) against the frame's FeatureVector of the last compute_bow.  Returns [(nmatches, match int32[N])]."""
        qs, keep = self._bow_queries(queries)
        res = (_capi.OrbxProjectionResult * len(queries))()
        check(self._L.orbx_search_by_bow(self._h, len(queries), qs, nnratio, int(check_orientation), res), self._h)
        out = []
        for r in res:
            out.append((r.nmatches, np.ctypeslib.as_array(C.cast(r.match, C.POINTER(C.c_int32)), shape=(max(r.n, 1),))[:r.n].copy()))
        del keep
        return out

    # ---------------------------------------------------------------- stage dumps (parity tests)
    def stage_dump(self, frame: int, level: int, stage: int):
        nbytes = C.c_size_t()
        check(self._L.orbx_stage_dump(self._h, frame, level, stage, None, 0, C.byref(nbytes)), self._h)
        buf = np.zeros(max(nbytes.value, 1), np.uint8)
        check(self._L.orbx_stage_dump(self._h, frame, level, stage, buf.ctypes.data, buf.size, C.byref(nbytes)), self._h)
        buf = buf[:nbytes.value]
        h, w = self._last_shape
        lw, lh = self.level_sizes(w, h)[level]
        if stage == _capi.STAGE_PYRAMID:
            return buf.reshape(lh + 38, lw + 38)
        if stage == _capi.STAGE_BLURRED:
            return buf.reshape(lh, lw)
        if stage in (_capi.STAGE_CANDIDATES, _capi.STAGE_KEPT):
            return buf.view(np.int32).reshape(-1, 3)
        if stage == _capi.STAGE_ANGLES:
            return buf.view(np.float32)
        raise ValueError(stage)


class Vocabulary:
    """An ORB vocabulary (DBoW2 tree) resident in HBM: `voc` is the flat layout of orbslam2_with_quadrics_b200.vocabulary."""

    def __init__(self, voc: dict, device: int = 0):
        self._L = _capi.lib()
        self._v = C.c_void_p()
        a = [np.ascontiguousarray(voc["child_start"], np.int32), np.ascontiguousarray(voc["child_items"], np.int32),
             np.ascontiguousarray(voc["node_desc"], np.uint8), np.ascontiguousarray(voc["node_weight"], np.float64),
             np.ascontiguousarray(voc["node_word"], np.int32)]
        check(self._L.orbx_vocabulary_create(device, int(voc["n_nodes"]), int(voc["L"]), a[0].ctypes.data, a[1].ctypes.data,
                                             a[2].ctypes.data, a[3].ctypes.data, a[4].ctypes.data, C.byref(self._v)), None)

    def close(self):
        if self._v:
            self._L.orbx_vocabulary_destroy(self._v)
            self._v = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
