"""ctypes binding of oracle/_ref/liborbref.so -- the reference's own, unmodified
ORBextractor translation unit compiled against oracle/shim (see oracle/Makefile).
TEST INFRASTRUCTURE: used by tests/, smoke() and bench.py's CPU-baseline legs only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .orb_oracle import KP_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = line.split()
                    return "avx2" in fl and "bmi2" in fl and "fma" in fl
    except OSError:
        pass
    return False


def build(force: bool = False) -> bool:
    """(Re)build oracle/_ref when the reference sources are present (build container only)."""
    if not os.path.isdir("/root/reference/src"):
        return os.path.exists(os.path.join(_HERE, "_ref", "liborbref.so"))
    cmd = ["make", "-C", _HERE] + (["-B"] if force else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return True


def available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "liborbref.so"))


def lib():
    global _lib
    if _lib is None:
        name = "liborbref.so" if _has_avx2() else "liborbref_v2.so"
        L = C.CDLL(os.path.join(_HERE, "_ref", name))
        L.orbref_create.restype = C.c_void_p
        L.orbref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orbref_destroy.argtypes = [C.c_void_p]
        L.orbref_extract.restype = C.c_int
        L.orbref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
        L.orbref_pyramid_level.restype = C.c_int
        L.orbref_pyramid_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orbref_distribute.restype = C.c_int
        L.orbref_distribute.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_int] * 5 + [C.c_void_p, C.c_int]
        L.orbref_tables.restype = C.c_int
        L.orbref_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orbref_scale_factor.restype = C.c_float
        L.orbref_scale_factor.argtypes = [C.c_void_p]
        for fn in ("cvshim_resize", "cvshim_border", "cvshim_blur", "cvshim_fast"):
            getattr(L, fn).restype = C.c_int if fn == "cvshim_fast" else None
        L.cvshim_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int]
        L.cvshim_border.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int]
        L.cvshim_blur.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p]
        L.cvshim_fast.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        L.cvshim_fast_atan2.restype = C.c_float
        L.cvshim_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib = L
    return _lib


class RefORBextractor:
    """The reference ORBextractor (unmodified TU) behind the same call shape as the oracle."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST):
        self.L = lib()
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = self.L.orbref_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orbref_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        a = [np.zeros(n, np.float32) for _ in range(4)]
        q = np.zeros(n, np.int32)
        u = np.zeros(16, np.int32)
        self.L.orbref_tables(self.h, *[x.ctypes.data for x in a], q.ctypes.data, u.ctypes.data)
        return a[0], a[1], a[2], a[3], q, u

    def __call__(self, image: np.ndarray):
        """Returns (keypoints structured array, descriptors) or None for an empty image."""
        if image.size == 0:
            h, w = (image.shape + (0, 0))[:2]
            r = self.L.orbref_extract(self.h, None, int(w), int(h), 0, 0, None, None)
            assert r == -1
            return None
        assert image.dtype == np.uint8 and image.ndim == 2 and image.strides[1] == 1
        cap = self.nfeatures + 64 * self.nlevels
        while True:
            kps = np.zeros(cap, KP_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            n = self.L.orbref_extract(self.h, image.ctypes.data, image.shape[1], image.shape[0], image.strides[0],
                                      cap, kps.ctypes.data, desc.ctypes.data)
            if n <= cap:
                return kps[:n].copy(), desc[:n].copy()
            cap = n

    def pyramid(self):
        out = []
        w = C.c_int()
        h = C.c_int()
        for l in range(self.nlevels):
            if self.L.orbref_pyramid_level(self.h, l, None, C.byref(w), C.byref(h)) != 0:
                return None
            buf = np.zeros((h.value + 38, w.value + 38), np.uint8)
            self.L.orbref_pyramid_level(self.h, l, buf.ctypes.data, C.byref(w), C.byref(h))
            out.append(buf)
        return out

    def distribute(self, xs, ys, rs, minX, maxX, minY, maxY, N):
        xyr = np.ascontiguousarray(np.stack([xs, ys, rs], axis=1).astype(np.int32))
        M = len(xyr)
        out = np.zeros(M + 8, np.int32)
        n = self.L.orbref_distribute(self.h, xyr.ctypes.data, M, minX, maxX, minY, maxY, N, out.ctypes.data, len(out))
        return out[:n].astype(np.int64)
