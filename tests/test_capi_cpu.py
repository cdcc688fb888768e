"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports exactly the symbols
include/orbx.h declares; without a GPU the product fails loudly (no CPU fallback); host geometry."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry
from orbslam2_with_quadrics_b200 import _capi, geometry
from orbslam2_with_quadrics_b200.frames import CONFIGS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return _capi.lib()


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "orbx.h")).read()
    return sorted(set(re.findall(r"ORBX_API[^;(]*?\b(orbx_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib):
    syms = header_symbols()
    assert len(syms) >= 24
    assert sorted(_capi.SYMBOLS) == syms
    for s in syms:
        assert hasattr(lib, s), s


def test_strerror_and_version(lib):
    assert lib.orbx_version().startswith(b"orbx-b200")
    for code in range(0, -8, -1):
        assert lib.orbx_strerror(code) not in (None, b"unknown status")
    assert lib.orbx_strerror(-99) == b"unknown status"


def test_bad_config_is_rejected_before_touching_cuda(lib):
    h = C.c_void_p()
    for bad in [dict(nlevels=0), dict(nlevels=17), dict(nfeatures=0), dict(scale_factor=1.0), dict(scale_factor=2.5), dict(ini_th_fast=0), dict(max_batch=0)]:
        kw = dict(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th_fast=20, min_th_fast=7, device=0, max_batch=1,
                  download_pyramid=1, candidate_divisor=0)
        kw.update(bad)
        cfg = _capi.OrbxConfig(**kw)
        assert lib.orbx_create(C.byref(cfg), C.byref(h)) == _capi.ERR_BAD_ARGS
    assert lib.orbx_create(None, C.byref(h)) == _capi.ERR_BAD_ARGS


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from orbslam2_with_quadrics_b200 import ORBextractor, OrbxError
    with pytest.raises(OrbxError) as e:
        ORBextractor(1000, 1.2, 8, 20, 7)
    assert e.value.status == _capi.ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "orbslam2_with_quadrics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cc", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "oracle/" not in txt.replace("oracle/ ", ""), f


def test_algorithmic_bytes_match_survey_table():
    want = {"mono_tum": 6564354, "stereo_euroc": 7731981, "stereo_kitti": 10587233, "rgbd_1080p": 37187353, "mono_4k": 146616792}
    for name, b in want.items():
        w, h, nf, sf, nl, *_ = CONFIGS[name]
        assert geometry.algorithmic_bytes(w, h, nf, sf, nl) == b
        assert sum(geometry.stage_algorithmic_bytes(w, h, nf, sf, nl).values()) == b


def test_geometry_tables_match_oracle():
    from oracle.orb_oracle import OrbParams
    for nf, sf, nl in [(1000, 1.2, 8), (4000, 1.2, 10), (500, 1.5, 4), (300, 1.1, 12)]:
        p = OrbParams(nf, sf, nl, 20, 7)
        t = geometry.scale_tables(sf, nl)
        assert np.array_equal(t[0].view(np.uint32), p.mvScaleFactor.view(np.uint32))
        assert np.array_equal(t[1].view(np.uint32), p.mvInvScaleFactor.view(np.uint32))
        assert geometry.level_quotas(nf, sf, nl) == p.mnFeaturesPerLevel
        assert geometry.level_sizes(1920, 1080, sf, nl) == p.level_sizes(1920, 1080)
    assert geometry.level_sizes(640, 480, 1.2, 8) == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]
