"""The OpenCV-typed C++ class ORB-SLAM2 links against (orbslam2_with_quadrics_b200/cpp), compiled here
against the OpenCV shim (real OpenCV headers do not exist in this image -- real-OpenCV linkage is
unverified).  CPU: it compiles and links against liborbx.so.  GPU: a small C++
program shaped like Frame::ExtractORB reproduces the oracle bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

import __graft_entry__ as entry
from orbslam2_with_quadrics_b200 import frames as fr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "orbslam2_with_quadrics_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "adapter_main")


def build_adapter():
    entry.build()
    src = [os.path.join(ROOT, "tests", "cpp", "adapter_main.cc"), os.path.join(PKG, "cpp", "ORBextractor.cc")]
    newest = max(os.path.getmtime(p) for p in src + [os.path.join(PKG, "cpp", "ORBextractor.h"), os.path.join(PKG, "liborbx.so")])
    if os.path.exists(EXE) and os.path.getmtime(EXE) >= newest:
        return EXE
    cmd = ["g++", "-std=c++11", "-O2", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(PKG, "cpp"),
           "-I" + os.path.join(ROOT, "include")] + src + ["-L" + PKG, "-lorbx", "-Wl,-rpath," + PKG, "-o", EXE]
    subprocess.run(cmd, check=True)
    return EXE


def test_adapter_compiles_against_the_reference_api_surface():
    exe = build_adapter()
    out = subprocess.run(["nm", "-C", "--defined-only", exe], capture_output=True, text=True).stdout
    for sym in ("ORB_SLAM2::ORBextractor::ORBextractor(int, float, int, int, int)",
                "ORB_SLAM2::ORBextractor::operator()(cv::_InputArray const&, cv::_InputArray const&, std::vector<cv::KeyPoint",
                "ORB_SLAM2::ORBextractor::GetScaleFactors()", "ORB_SLAM2::ORBextractor::GetInverseScaleSigmaSquares()",
                "ORB_SLAM2::ORBextractor::GetLevels()", "ORB_SLAM2::ORBextractor::GetScaleFactor()",
                "ORB_SLAM2::ORBextractor::ComputeStereoMatches(", "ORB_SLAM2::ORBextractor::UndistortAndAssignToGrid(",
                "ORB_SLAM2::ORBextractor::SetColorOrder(bool)"):
        assert sym in out, sym


def _adapter_vs_oracle(tmp_path, config, seed, env=None):
    from oracle import orb_oracle
    exe = build_adapter()
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS[config]
    img = fr.cluttered_scene(w, h, seed)
    raw, outp = tmp_path / "in.raw", tmp_path / "out.bin"
    img.tofile(raw)
    e = dict(os.environ)
    e.update(env or {})
    subprocess.run([exe, str(w), str(h), str(nf), str(sf), str(nl), str(it), str(mt), str(raw), str(outp)], check=True, env=e)
    ro = orb_oracle.ORBextractor(nf, sf, nl, it, mt)(img)
    b = outp.read_bytes()
    n, nlv = np.frombuffer(b, np.int32, 2)
    assert (n, nlv) == (ro.n, nl)
    off = 8
    kp = np.frombuffer(b, orb_oracle.KP_DTYPE, n, off); off += 28 * n
    ds = np.frombuffer(b, np.uint8, 32 * n, off).reshape(n, 32); off += 32 * n
    for f in kp.dtype.names:
        assert np.array_equal(kp[f], ro.keypoints[f]), f
    assert np.array_equal(ds, ro.descriptors)
    sfs = np.frombuffer(b, np.float32, nl, off); off += 4 * nl
    isg = np.frombuffer(b, np.float32, nl, off); off += 4 * nl
    p = orb_oracle.OrbParams(nf, sf, nl, it, mt)
    assert np.array_equal(sfs, p.mvScaleFactor) and np.array_equal(isg, p.mvInvLevelSigma2)
    for l in range(nl):
        lw, lh = np.frombuffer(b, np.int32, 2, off); off += 8
        plane = np.frombuffer(b, np.uint8, (lw + 38) * (lh + 38), off).reshape(lh + 38, lw + 38); off += plane.size
        assert np.array_equal(plane, ro.pyramid[l]), l


@pytest.mark.gpu
def test_adapter_matches_oracle(tmp_path):
    _adapter_vs_oracle(tmp_path, "stereo_euroc", 31)


@pytest.mark.gpu
def test_adapter_pageable_odd_stride_1080p(tmp_path):
    """What Frame::ExtractORB hands over in practice: an ordinary (pageable) cv::Mat whose rows do not start on any
    alignment boundary (step 1933, data pointer odd), at the size the headline number is quoted on.  Goes through the
    library's row-wise staging path; keypoints, descriptors and every padded pyramid plane bit-exact vs the oracle."""
    _adapter_vs_oracle(tmp_path, "rgbd_1080p", 77, env={"ORBX_TEST_STRIDE": "1933"})


@pytest.mark.gpu
def test_adapter_on_second_device(tmp_path):
    """ORBextractor::SetDevice(1): the adapter's handle, its streams, pinned buffers and TMA descriptors on a GPU that is not
    device 0 (needs a 2-GPU box: `gpurun --gpus 2`)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _adapter_vs_oracle(tmp_path, "stereo_euroc", 32, env={"ORBX_TEST_DEVICE": "1", "ORBX_TEST_STRIDE": "765"})


@pytest.mark.gpu
def test_adapter_stereo_matches_oracle(tmp_path):
    """Two adapter objects (one per eye, src/Frame.cc:78-81) + the GPU ComputeStereoMatches vs the stereo oracle."""
    from oracle import orb_oracle, stereo_oracle
    exe = build_adapter()
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["stereo_euroc"]
    left, right = fr.stereo_pair(w, h, 4242)
    mbf = 47.90639384423901
    mb = float(np.float32(mbf) / np.float32(435.2046959714599))
    rawl, rawr, outp, sout = tmp_path / "l.raw", tmp_path / "r.raw", tmp_path / "out.bin", tmp_path / "stereo.bin"
    left.tofile(rawl); right.tofile(rawr)
    subprocess.run([exe, str(w), str(h), str(nf), str(sf), str(nl), str(it), str(mt), str(rawl), str(outp), str(rawr), str(sout),
                    repr(mbf), repr(mb)], check=True)
    ex = orb_oracle.ORBextractor(nf, sf, nl, it, mt)
    rl, rr = ex(left), ex(right)
    u, d, _ = stereo_oracle.compute_stereo_matches(rl.keypoints, rl.descriptors, rl.pyramid, rr.keypoints, rr.descriptors, rr.pyramid,
                                                   ex.GetScaleFactors(), ex.GetInverseScaleFactors(), mbf, mb)
    b = sout.read_bytes()
    n = int(np.frombuffer(b, np.int32, 1)[0])
    assert n == rl.n
    ug, dg = np.frombuffer(b, np.float32, n, 4), np.frombuffer(b, np.float32, n, 4 + 4 * n)
    assert (ug >= 0).sum() > 100
    assert np.array_equal(ug.view(np.uint32), u.view(np.uint32)) and np.array_equal(dg.view(np.uint32), d.view(np.uint32))
