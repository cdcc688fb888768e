"""ctypes binding of liborbx.so (include/orbx.h).  Fails loudly when the CUDA library is
missing or cannot be loaded: there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORBX_LIB") or os.path.join(PKG, "liborbx.so")     # ORBX_LIB: A/B builds of the same library

ORBX_OK = 0
ERR_BAD_ARGS, ERR_BAD_GEOMETRY, ERR_CUDA, ERR_CANDIDATE_OVERFLOW, ERR_NO_DEVICE, ERR_OOM, ERR_EMPTY_IMAGE = \
    -1, -2, -3, -4, -5, -6, -7
STAGE_PYRAMID, STAGE_CANDIDATES, STAGE_KEPT, STAGE_ANGLES, STAGE_BLURRED = range(5)

# every symbol include/orbx.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "orbx_create", "orbx_destroy", "orbx_extract", "orbx_extract_batch", "orbx_extract_device",
    "orbx_fetch_results", "orbx_alloc_host", "orbx_free_host", "orbx_pyramid_level", "orbx_scale_tables",
    "orbx_get_levels", "orbx_get_scale_factor", "orbx_level_quotas", "orbx_level_sizes", "orbx_stage_dump",
    "orbx_stream", "orbx_synchronize", "orbx_stage_timing_enable", "orbx_stage_timing_read",
    "orbx_launch_count", "orbx_algorithmic_bytes", "orbx_strerror", "orbx_last_cuda_error", "orbx_version",
    "orbx_stereo_match", "orbx_stereo_match_device", "orbx_stereo_fetch",
    "orbx_extract_batch_color", "orbx_extract_device_color", "orbx_undistort_grid", "orbx_fast_stats",
    "orbx_search_by_projection", "orbx_search_by_projection_device", "orbx_search_by_projection_fetch",
    "orbx_search_local_points", "orbx_search_local_points_device",
    "orbx_search_by_projection_kf", "orbx_search_by_projection_kf_device", "orbx_search_for_initialization",
    "orbx_search_by_bow", "orbx_search_by_bow_device", "orbx_vocabulary_create", "orbx_vocabulary_destroy", "orbx_compute_bow", "orbx_compute_bow_device",
    "orbx_bind_thread_to_device", "orbx_is_in_frustum",
]
GRAY8, BGR8, RGB8, BGRA8, RGBA8 = range(5)

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


class OrbxConfig(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int),
                ("ini_th_fast", C.c_int), ("min_th_fast", C.c_int), ("device", C.c_int),
                ("max_batch", C.c_int), ("download_pyramid", C.c_int), ("candidate_divisor", C.c_int),
                ("device_chunks", C.c_int), ("reserved", C.c_int * 6)]


class OrbxResult(C.Structure):
    _fields_ = [("n", C.c_int), ("status", C.c_int), ("kps", C.c_void_p), ("desc", C.c_void_p)]


class OrbxStereoResult(C.Structure):
    _fields_ = [("n", C.c_int), ("u_right", C.c_void_p), ("depth", C.c_void_p)]


class OrbxGridResult(C.Structure):
    _fields_ = [("n", C.c_int), ("n_in_grid", C.c_int), ("xy_un", C.c_void_p), ("cell_start", C.c_void_p),
                ("cell_items", C.c_void_p), ("bounds", C.c_float * 4)]


class OrbxProjectionQuery(C.Structure):
    _fields_ = [("cur_frame", C.c_int), ("n_last", C.c_int), ("world_pos", C.c_void_p), ("mp_desc", C.c_void_p),
                ("mp_obs", C.c_void_p), ("outlier", C.c_void_p), ("octave", C.c_void_p), ("angle", C.c_void_p),
                ("Tcw_cur", C.c_float * 16), ("Tcw_last", C.c_float * 16)]


class OrbxLocalPointsQuery(C.Structure):
    _fields_ = [("cur_frame", C.c_int), ("n_points", C.c_int), ("in_view", C.c_void_p), ("proj_xy_xr", C.c_void_p),
                ("scale_level", C.c_void_p), ("view_cos", C.c_void_p), ("mp_desc", C.c_void_p), ("mp_obs", C.c_void_p),
                ("cur_obs", C.c_void_p)]


class OrbxFrustumQuery(C.Structure):
    _fields_ = [("n_points", C.c_int), ("consider", C.c_void_p), ("world_pos", C.c_void_p), ("normal", C.c_void_p),
                ("min_dist", C.c_void_p), ("max_dist", C.c_void_p), ("Tcw", C.c_float * 16)]


class OrbxFrustumResult(C.Structure):
    _fields_ = [("n", C.c_int), ("n_in_view", C.c_int), ("in_view", C.c_void_p), ("proj_xy_xr", C.c_void_p),
                ("scale_level", C.c_void_p), ("view_cos", C.c_void_p)]


class OrbxKeyframeProjectionQuery(C.Structure):
    _fields_ = [("cur_frame", C.c_int), ("n_points", C.c_int), ("search", C.c_void_p), ("world_pos", C.c_void_p),
                ("pred_level", C.c_void_p), ("mp_desc", C.c_void_p), ("kf_angle", C.c_void_p), ("cur_held", C.c_void_p),
                ("Tcw_cur", C.c_float * 16)]


class OrbxInitializationQuery(C.Structure):
    _fields_ = [("cur_frame", C.c_int), ("n1", C.c_int), ("octave1", C.c_void_p), ("angle1", C.c_void_p), ("desc1", C.c_void_p),
                ("prev_matched", C.c_void_p)]


class OrbxInitializationResult(C.Structure):
    _fields_ = [("n1", C.c_int), ("nmatches", C.c_int), ("matches12", C.c_void_p), ("prev_matched", C.c_void_p)]


class OrbxBowResult(C.Structure):
    _fields_ = [("n_words", C.c_int), ("word_ids", C.c_void_p), ("word_values", C.c_void_p), ("n_features", C.c_int),
                ("fv_nodes", C.c_void_p), ("fv_features", C.c_void_p)]


class OrbxBowMatchQuery(C.Structure):
    _fields_ = [("cur_frame", C.c_int), ("n_kf", C.c_int), ("kf_desc", C.c_void_p), ("kf_valid", C.c_void_p), ("kf_angle", C.c_void_p),
                ("n_kf_fv", C.c_int), ("kf_fv_nodes", C.c_void_p), ("kf_fv_features", C.c_void_p)]


class OrbxProjectionResult(C.Structure):
    _fields_ = [("n", C.c_int), ("nmatches", C.c_int), ("rounds", C.c_int), ("match", C.c_void_p)]


class OrbxError(RuntimeError):
    def __init__(self, status: int, detail: str = ""):
        self.status = status
        msg = lib().orbx_strerror(status).decode() if _lib is not None else str(status)
        super().__init__("orbx status %d: %s%s" % (status, msg, (" -- " + detail) if detail else ""))


_lib = None


def lib():
    """Loads liborbx.so; raises (never falls back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "liborbx.so is not built (%s). Run `python -m orbslam2_with_quadrics_b200.build`; "
            "this package has no CPU path." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    L.orbx_create.argtypes = [C.POINTER(OrbxConfig), C.POINTER(vp)]
    L.orbx_destroy.argtypes = [vp]
    L.orbx_extract.argtypes = [vp, vp, i, i, sz, C.POINTER(OrbxResult)]
    L.orbx_extract_batch.argtypes = [vp, i, C.POINTER(vp), i, i, C.POINTER(sz), C.POINTER(OrbxResult)]
    L.orbx_extract_device.argtypes = [vp, i, vp, i, i, sz, sz]
    L.orbx_fetch_results.argtypes = [vp, i, C.POINTER(OrbxResult)]
    L.orbx_alloc_host.argtypes = [sz, C.POINTER(vp)]
    L.orbx_free_host.argtypes = [vp]
    L.orbx_pyramid_level.argtypes = [vp, i, i, C.POINTER(vp), C.POINTER(i), C.POINTER(i), C.POINTER(sz)]
    L.orbx_scale_tables.argtypes = [vp] + [C.POINTER(C.POINTER(C.c_float))] * 4
    L.orbx_get_levels.argtypes = [vp]
    L.orbx_get_scale_factor.argtypes = [vp]
    L.orbx_get_scale_factor.restype = C.c_float
    L.orbx_level_quotas.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
    L.orbx_level_sizes.argtypes = [vp, i, i, C.POINTER(i), C.POINTER(i)]
    L.orbx_stage_dump.argtypes = [vp, i, i, i, vp, sz, C.POINTER(sz)]
    L.orbx_stream.argtypes = [vp]
    L.orbx_stream.restype = vp
    L.orbx_synchronize.argtypes = [vp]
    L.orbx_stage_timing_enable.argtypes = [vp, i]
    L.orbx_stage_timing_read.argtypes = [vp, i, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(i)]
    L.orbx_launch_count.argtypes = [vp]
    L.orbx_launch_count.restype = C.c_longlong
    L.orbx_algorithmic_bytes.argtypes = [vp, i, i]
    L.orbx_algorithmic_bytes.restype = C.c_longlong
    L.orbx_strerror.argtypes = [i]
    L.orbx_strerror.restype = C.c_char_p
    L.orbx_last_cuda_error.argtypes = [vp]
    L.orbx_last_cuda_error.restype = C.c_char_p
    L.orbx_version.restype = C.c_char_p
    L.orbx_stereo_match.argtypes = [vp, vp, i, C.POINTER(i), C.POINTER(i), C.c_float, C.c_float, C.POINTER(OrbxStereoResult)]
    L.orbx_stereo_match_device.argtypes = [vp, vp, i, C.POINTER(i), C.POINTER(i), C.c_float, C.c_float]
    L.orbx_stereo_fetch.argtypes = [vp, i, C.POINTER(i), C.POINTER(OrbxStereoResult)]
    L.orbx_extract_batch_color.argtypes = [vp, i, C.POINTER(vp), i, i, C.POINTER(sz), i, C.POINTER(OrbxResult)]
    L.orbx_extract_device_color.argtypes = [vp, i, vp, i, i, sz, sz, i]
    L.orbx_fast_stats.argtypes = [vp, i, C.POINTER(i), C.POINTER(i)]
    f = C.c_float
    L.orbx_search_by_projection.argtypes = [vp, i, C.POINTER(OrbxProjectionQuery), C.POINTER(f), f, f, f, i, i, i,
                                            C.POINTER(OrbxProjectionResult)]
    L.orbx_search_by_projection_device.argtypes = [vp, i, C.POINTER(OrbxProjectionQuery), C.POINTER(f), f, f, f, i, i, i]
    L.orbx_search_by_projection_fetch.argtypes = [vp, i, C.POINTER(OrbxProjectionQuery), C.POINTER(OrbxProjectionResult)]
    L.orbx_search_local_points.argtypes = [vp, i, C.POINTER(OrbxLocalPointsQuery), f, f, i, C.POINTER(OrbxProjectionResult)]
    L.orbx_search_local_points_device.argtypes = [vp, i, C.POINTER(OrbxLocalPointsQuery), f, f, i]
    L.orbx_is_in_frustum.argtypes = [vp, i, C.POINTER(OrbxFrustumQuery), C.POINTER(C.c_float), f, C.POINTER(C.c_float), f, f,
                                     C.POINTER(OrbxFrustumResult)]
    L.orbx_search_by_projection_kf.argtypes = [vp, i, C.POINTER(OrbxKeyframeProjectionQuery), C.POINTER(C.c_float), f, i, i,
                                               C.POINTER(OrbxProjectionResult)]
    L.orbx_search_by_projection_kf_device.argtypes = [vp, i, C.POINTER(OrbxKeyframeProjectionQuery), C.POINTER(C.c_float), f, i, i]
    L.orbx_search_for_initialization.argtypes = [vp, i, C.POINTER(OrbxInitializationQuery), f, i, i, C.POINTER(OrbxInitializationResult)]
    L.orbx_search_by_bow.argtypes = [vp, i, C.POINTER(OrbxBowMatchQuery), f, i, C.POINTER(OrbxProjectionResult)]
    L.orbx_search_by_bow_device.argtypes = [vp, i, C.POINTER(OrbxBowMatchQuery), f, i]
    L.orbx_vocabulary_create.argtypes = [i, i, i, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.orbx_vocabulary_destroy.argtypes = [vp]
    L.orbx_compute_bow.argtypes = [vp, vp, i, C.POINTER(i), i, C.POINTER(OrbxBowResult)]
    L.orbx_compute_bow_device.argtypes = [vp, vp, i, C.POINTER(i), i]
    L.orbx_undistort_grid.argtypes = [vp, i, C.POINTER(i), C.POINTER(C.c_float), C.POINTER(C.c_float), i, C.POINTER(OrbxGridResult)]
    L.orbx_bind_thread_to_device.argtypes = [i, C.POINTER(i), C.POINTER(i)]
    _lib = L
    return L


def check(status: int, handle=None):
    if status != ORBX_OK:
        detail = ""
        if handle is not None and status == ERR_CUDA:
            detail = lib().orbx_last_cuda_error(handle).decode()
        raise OrbxError(status, detail)
