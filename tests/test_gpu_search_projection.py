"""GPU parity of orbx_search_by_projection (ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono), reference
src/ORBmatcher.cc:1328-1470 with Frame::GetFeaturesInArea src/Frame.cc:327-380) against oracle/match_oracle.py, which is
pinned against the reference's own lines.  Bar: nmatches and the map-point assignment of every current keypoint identical
(integer results; the float32 projection feeding them is bit-reproduced, tolerance 0)."""
import numpy as np
import pytest

from oracle import match_oracle
from orbslam2_with_quadrics_b200 import ORBextractor, OrbxError, _capi
from orbslam2_with_quadrics_b200 import frames as fr
from orbslam2_with_quadrics_b200 import match_cases as mc

pytestmark = pytest.mark.gpu
K_TUM1, D_TUM1 = (517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)
K_KITTI, D_RECT = (718.856, 718.856, 607.1928, 185.2157), (0.0, 0.0, 0.0, 0.0)


def build_query(rng, frame, kps, desc, grid, K4, n_last, nlevels, motion):
    xy, start, items, bounds = grid
    Tc = mc.pose(rng)
    Tl = mc.pose(rng, t=(0.0, 0.0, {"still": 0.0, "forward": 0.9, "backward": -0.9}[motion]))
    last = mc.make_last_frame(rng, xy, kps["octave"].astype(np.int32), kps["angle"].astype(np.float32), desc, K4, Tc, n_last, nlevels)
    return dict(cur_frame=frame, Tcw_cur=Tc, Tcw_last=Tl, **last)


def oracle_result(q, kps, desc, grid, K4, mbf, mb, sf, th, mono, check, u_right):
    xy, start, items, bounds = grid
    return match_oracle.search_by_projection(
        q["world"], q["mp_desc"], q["mp_obs"], q["outlier"], q["last_octave"], q["last_angle"], q["Tcw_cur"], q["Tcw_last"], xy,
        kps["octave"].astype(np.int32), kps["angle"].astype(np.float32), desc, u_right, start, items, bounds, K4, mbf, mb, sf, th,
        mono, check)


def test_monocular_batch_matches_oracle():
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=4)
    res = gx.extract_batch([fr.cluttered_scene(w, h, 900 + i) for i in range(3)] + [fr.flat_frame(w, h)])
    grids = gx.undistort_grid(K_TUM1, D_TUM1)
    rng = np.random.default_rng(11)
    qs = [build_query(rng, f, res[f][0], res[f][1], grids[f], K_TUM1, n, nl, "still") for f, n in ((0, 1000), (1, 350), (2, 1700))]
    for th, check in ((15.0, True), (30.0, True), (15.0, False)):
        out = gx.search_by_projection(qs, K_TUM1, 40.0, 0.08, th, True, check)
        for q, (n, m, rounds) in zip(qs, out):
            f = q["cur_frame"]
            n0, m0 = oracle_result(q, res[f][0], res[f][1], grids[f], K_TUM1, 40.0, 0.08, gx.GetScaleFactors(), th, True, check, None)
            assert n == n0 and np.array_equal(m, m0)
            assert n > 0.2 * len(q["mp_obs"]) and rounds >= 2          # contention really needed more than one round
    # queries in another order, one of them against the frame without keypoints, one with an empty LastFrame
    empty = dict(qs[0]); empty.update(cur_frame=3)
    none = {k: (v[:0] if isinstance(v, np.ndarray) and v.ndim and len(v) == len(qs[1]["mp_obs"]) else v) for k, v in qs[1].items()}
    out = gx.search_by_projection([qs[2], empty, none], K_TUM1, 40.0, 0.08, 15.0, True)
    n0, m0 = oracle_result(qs[2], res[2][0], res[2][1], grids[2], K_TUM1, 40.0, 0.08, gx.GetScaleFactors(), 15.0, True, True, None)
    assert out[0][0] == n0 and np.array_equal(out[0][1], m0)
    assert out[1][0] == 0 and len(out[1][1]) == 0
    assert out[2][0] == 0 and np.all(out[2][1] == -1) and len(out[2][1]) == len(res[1][0])
    gx.close()


@pytest.mark.parametrize("list_cap", ["1", "2"])
def test_full_search_path_of_the_resolve_kernel(list_cap, monkeypatch):
    """Candidate lists that do not fit (forced here by shrinking them to 1 or 2 entries) fall back to the full search
    inside the resolve kernel; the result is the same."""
    monkeypatch.setenv("ORBX_SP_LIST_CAP", list_cap)
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["stereo_euroc"]
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=2)
    res = gx.extract_batch([fr.cluttered_scene(w, h, 51), fr.cluttered_scene(w, h, 52)])
    K4, D = (458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)
    grids = gx.undistort_grid(K4, D)
    rng = np.random.default_rng(int(list_cap))
    # few flipped bits and a wide window: many candidates at distance <= TH_HIGH per point
    qs = []
    for f in (0, 1):
        xy, start, items, bounds = grids[f]
        Tc = mc.pose(rng)
        last = mc.make_last_frame(rng, xy, res[f][0]["octave"].astype(np.int32), res[f][0]["angle"].astype(np.float32), res[f][1], K4, Tc,
                                  1500, nl, dup_frac=0.6, flip_bits=6)
        qs.append(dict(cur_frame=f, Tcw_cur=Tc, Tcw_last=mc.pose(rng), **last))
    out = gx.search_by_projection(qs, K4, 47.9, 0.11, 40.0, True)
    for q, (n, m, rounds) in zip(qs, out):
        f = q["cur_frame"]
        n0, m0 = oracle_result(q, res[f][0], res[f][1], grids[f], K4, 47.9, 0.11, gx.GetScaleFactors(), 40.0, True, True, None)
        assert n == n0 and np.array_equal(m, m0) and n > 300
    gx.close()


@pytest.mark.parametrize("motion", ["still", "forward", "backward"])
def test_stereo_frame_with_device_uright_matches_oracle(motion):
    """The stereo Frame: mvuRight comes from orbx_stereo_match on the same handle and never leaves HBM."""
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["stereo_kitti"]
    left, right = fr.stereo_pair(w, h, 40)
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=2)
    res = gx.extract_batch([left, right])
    mbf, mb = 386.1448, 0.5371657
    (u_right, _), = gx.stereo_match(gx, mbf, mb, left_frames=[0], right_frames=[1])
    grids = gx.undistort_grid(K_KITTI, D_RECT)
    rng = np.random.default_rng({"still": 21, "forward": 22, "backward": 23}[motion])
    q = build_query(rng, 0, res[0][0], res[0][1], grids[0], K_KITTI, 1900, nl, motion)
    assert match_oracle.motion_flags(q["Tcw_cur"], q["Tcw_last"], mb, False) == (motion == "forward", motion == "backward")
    for th in (7.0, 14.0):
        (n, m, rounds), = gx.search_by_projection([q], K_KITTI, mbf, mb, th, False, True, use_stereo=True)
        n0, m0 = oracle_result(q, res[0][0], res[0][1], grids[0], K_KITTI, mbf, mb, gx.GetScaleFactors(), th, False, True, u_right)
        assert n == n0 and np.array_equal(m, m0) and n > 100
    assert (u_right > 0).sum() > 100                                     # the mvuRight test (:1405-1411) really ran
    gx.close()


def test_4k_and_bad_arguments():
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_4k"]
    K4, D = (3100.0, 3098.0, 1915.5, 1082.25), (0.12, -0.31, 0.0007, -0.0004, 0.09)
    gx = ORBextractor(nf, sf, nl, it, mt)
    kps, desc = gx(fr.cluttered_scene(w, h, 9))
    rng = np.random.default_rng(5)
    with pytest.raises(OrbxError):
        gx.search_by_projection([dict(cur_frame=0, world=np.zeros((1, 3)), mp_desc=np.zeros((1, 32)), mp_obs=[1], last_octave=[0],
                                      last_angle=[0.0], Tcw_cur=np.eye(4), Tcw_last=np.eye(4))], K4, 0.0, 0.0, 15.0, True)   # no grid yet
    (grid,) = gx.undistort_grid(K4, D)
    q = build_query(rng, 0, kps, desc, grid, K4, 4000, nl, "still")
    (n, m, rounds), = gx.search_by_projection([q], K4, 0.0, 0.0, 15.0, True)
    n0, m0 = oracle_result(q, kps, desc, grid, K4, 0.0, 0.0, gx.GetScaleFactors(), 15.0, True, True, None)
    assert n == n0 and np.array_equal(m, m0)
    bad = dict(q); bad.update(cur_frame=1)
    with pytest.raises(OrbxError):
        gx.search_by_projection([bad], K4, 0.0, 0.0, 15.0, True)
    bad = dict(q); bad["last_octave"] = np.full(len(q["mp_obs"]), nl, np.int32)
    with pytest.raises(OrbxError):
        gx.search_by_projection([bad], K4, 0.0, 0.0, 15.0, True)
    with pytest.raises(OrbxError):
        gx.search_by_projection([q], K4, 0.0, 0.0, 15.0, False, True, use_stereo=True)      # no stereo match on this handle
    gx.close()


def test_timing_report_beside_reference_lines():
    """Measurement of the row: device time of one launch over a batch of (LastFrame, CurrentFrame) queries (CUDA events on
    the handle's stream, staging H2D included), end-to-end wall time through the C ABI, and the reference's own lines
    (oracle/_ref, one host thread) on the same queries.  Written to gpurun_out/r01_search_projection.json when that
    directory exists; the assertions are parity only."""
    import json
    import os
    import time

    import torch
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["rgbd_1080p"]
    B = 16
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=B, download_pyramid=False)
    imgs = [fr.cluttered_scene(w, h, 3000 + i) for i in range(4)]
    res = gx.extract_batch([imgs[i % 4] for i in range(B)])
    K4, D = (1050.0, 1050.0, 959.5, 539.5), (0.05, -0.11, 0.0004, -0.0003, 0.02)
    grids = gx.undistort_grid(K4, D)
    rng = np.random.default_rng(77)
    qs = [build_query(rng, f, res[f][0], res[f][1], grids[f], K4, len(res[f][0]), nl, "still") for f in range(B)]
    out = gx.search_by_projection(qs, K4, 0.0, 0.0, 15.0, True)
    prepared = gx._projection_queries(qs)
    st = torch.cuda.ExternalStream(gx.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        gx.search_by_projection_device(prepared, K4, 0.0, 0.0, 15.0, True)
    gx.synchronize()
    K = 20
    t0 = time.perf_counter()
    e0.record(st)
    for _ in range(K):
        gx.search_by_projection_device(prepared, K4, 0.0, 0.0, 15.0, True)
    e1.record(st)
    gx.synchronize()
    wall_dev = (time.perf_counter() - t0) / K * 1e3
    dev_ms = e0.elapsed_time(e1) / K
    import ctypes as C
    from orbslam2_with_quadrics_b200 import _capi
    cres = (_capi.OrbxProjectionResult * B)()
    k4 = (C.c_float * 4)(*K4)
    t0 = time.perf_counter()
    for _ in range(K):                                     # the C ABI call itself (queries marshalled once, no numpy copies)
        _capi.check(gx._L.orbx_search_by_projection(gx._h, B, prepared[0], k4, 0.0, 0.0, 15.0, 1, 1, 0, cres), gx._h)
    e2e_ms = (time.perf_counter() - t0) / K * 1e3
    cpu_ms = None
    if match_oracle.ref_available():
        t0 = time.perf_counter()
        for q, (n, m, _) in zip(qs[:4], out[:4]):
            f = q["cur_frame"]
            xy, start, items, bounds = grids[f]
            n0, m0 = match_oracle.ref_search_by_projection(
                q["world"], q["mp_desc"], q["mp_obs"], q["outlier"], q["last_octave"], q["last_angle"], q["Tcw_cur"], q["Tcw_last"],
                xy, res[f][0]["octave"].astype(np.int32), res[f][0]["angle"].astype(np.float32), res[f][1], None, start, items,
                bounds, K4, 0.0, 0.0, gx.GetScaleFactors(), 15.0, True, True)
            assert n == n0 and np.array_equal(m, m0)
        cpu_ms = (time.perf_counter() - t0) / 4 * 1e3
    rep = {"workload": "SearchByProjection, %d queries of %d LastFrame map points against 1920x1080 frames of ~%d keypoints, th 15, mono"
                       % (B, len(qs[0]["mp_obs"]), len(res[0][0])),
           "device_ms_per_batch_incl_staging_h2d": dev_ms, "device_us_per_query": dev_ms / B * 1e3,
           "host_wall_ms_per_batch_enqueue_only": wall_dev, "e2e_ms_per_batch_with_d2h": e2e_ms, "e2e_us_per_query": e2e_ms / B * 1e3,
           "reference_lines_cpu_ms_per_query_1_thread": cpu_ms, "rounds": [r for _, _, r in out],
           "nmatches": [n for n, _, _ in out]}
    print(json.dumps(rep))
    if os.path.isdir("gpurun_out"):
        json.dump(rep, open("gpurun_out/r01_search_projection.json", "w"), indent=1)
    gx.close()


def test_against_golden_of_the_reference_lines():
    """tests/golden/match_golden.json was produced by the reference's own lines on the ORACLE's extraction of the same
    frame; the GPU extraction is bit-identical, so the whole chain extract -> undistort/grid -> search lands on the digests."""
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_match_golden", os.path.join(here, "golden", "make_match_golden.py"))
    mmg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mmg)
    gold = json.load(open(os.path.join(here, "golden", "match_golden.json")))
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    gx = ORBextractor(nf, sf, nl, it, mt)
    kps, desc = gx(fr.cluttered_scene(w, h, 77))
    (grid,) = gx.undistort_grid(mmg.K_TUM1, mmg.D_TUM1)
    xy, start, items, bounds = grid
    cf = dict(xy_un=xy, cur_octave=kps["octave"].astype(np.int32), cur_angle=kps["angle"].astype(np.float32), desc=desc,
              cell_start=start, cell_items=items, bounds=bounds, sf=np.asarray(gx.GetScaleFactors(), np.float32), nlevels=nl)
    done = 0
    for case in mmg.CASES:
        if case[3]:
            continue                                   # stereo Frames take mvuRight from orbx_stereo_match (test above)
        sc = mmg.scenario(cf, *case[:4])
        q = dict(cur_frame=0, world=sc["world"], mp_desc=sc["mp_desc"], mp_obs=sc["mp_obs"], outlier=sc["outlier"],
                 last_octave=sc["last_octave"], last_angle=sc["last_angle"], Tcw_cur=sc["Tcw_cur"], Tcw_last=sc["Tcw_last"])
        for check in (True, False):
            (n, m, _), = gx.search_by_projection([q], mmg.K_TUM1, sc["mbf"], sc["mb"], case[4], case[5], check)
            assert mmg.digest(sc, n, m) == gold[mmg.key(case, check)]
            done += 1
    assert done == 6
    gx.close()


# ----------------------------------------------------------------------------- SearchLocalPoints' matcher (:45-129)
def local_oracle(q, kps, desc, grid, sf, th, nnratio, u_right):
    xy, start, items, bounds = grid
    return match_oracle.search_local_points(q["in_view"], q["proj_x"], q["proj_y"], q["proj_xr"], q["scale_level"], q["view_cos"],
                                            q["mp_desc"], q["mp_obs"], xy, kps["octave"].astype(np.int32), desc, u_right, q["cur_obs"],
                                            start, items, bounds, sf, th, nnratio)


@pytest.mark.parametrize("list_cap", [None, "2"])
def test_local_points_batch_matches_oracle(list_cap, monkeypatch):
    if list_cap:
        monkeypatch.setenv("ORBX_SP_LIST_CAP", list_cap)
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=3)
    res = gx.extract_batch([fr.cluttered_scene(w, h, 700 + i) for i in range(3)])
    grids = gx.undistort_grid(K_TUM1, D_TUM1)
    rng = np.random.default_rng(31)
    qs = []
    for f, n in ((0, 1500), (1, 400), (2, 2500)):
        lp = mc.make_local_points(rng, grids[f][0], res[f][0]["octave"].astype(np.int32), res[f][1], n, nl)
        qs.append(dict(cur_frame=f, **lp))
    for th, ratio in ((1.0, 0.8), (3.0, 0.8), (5.0, 0.9), (3.0, 0.3)):
        out = gx.search_local_points(qs, th, ratio)
        for q, (n, m, rounds) in zip(qs, out):
            f = q["cur_frame"]
            n0, m0 = local_oracle(q, res[f][0], res[f][1], grids[f], gx.GetScaleFactors(), th, ratio, None)
            assert n == n0 and np.array_equal(m, m0)
            assert ratio < 0.5 or n > 0.15 * len(q["mp_obs"])
    # no held points at all (cur_obs NULL) == all -1
    q = dict(qs[0]); q["cur_obs"] = None
    (n, m, _), = gx.search_local_points([q], 3.0)
    q2 = dict(qs[0]); q2["cur_obs"] = np.full(len(res[0][0]), -1, np.int32)
    n0, m0 = local_oracle(q2, res[0][0], res[0][1], grids[0], gx.GetScaleFactors(), 3.0, 0.8, None)
    assert n == n0 and np.array_equal(m, m0)
    # Points Frame::isInFrustum never accepted carry an UNINITIALISED mnTrackScaleLevel in the reference (set only at
    # src/Frame.cc:321, neither MapPoint constructor touches it, src/MapPoint.cc:32-73): garbage levels on skipped points
    # must not fail the call nor change the result; a garbage level on a point that IS processed is still refused.
    q3 = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in qs[0].items()}
    skipped = np.flatnonzero(q3["in_view"] == 0)
    assert len(skipped) > 10
    q3["scale_level"][skipped[::2]] = np.iinfo(np.int32).min
    q3["scale_level"][skipped[1::2]] = 12345
    (n3, m3, _), = gx.search_local_points([q3], 3.0)
    n0, m0 = local_oracle(qs[0], res[0][0], res[0][1], grids[0], gx.GetScaleFactors(), 3.0, 0.8, None)
    assert n3 == n0 and np.array_equal(m3, m0)
    q4 = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in qs[0].items()}
    q4["scale_level"][np.flatnonzero(q4["in_view"] != 0)[0]] = nl
    with pytest.raises(OrbxError) as ei:
        gx.search_local_points([q4], 3.0)
    assert ei.value.status == _capi.ERR_BAD_ARGS
    # a new extraction invalidates mvKeysUn / mGrid: matching against the previous frame's grid is refused, not silent
    gx.extract_batch([fr.cluttered_scene(w, h, 900)])
    with pytest.raises(OrbxError) as ei:
        gx.search_local_points([dict(qs[0], cur_frame=0)], 3.0)
    assert ei.value.status == _capi.ERR_BAD_ARGS
    gx.close()


def test_local_points_stereo_frame_and_golden():
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_match_golden", os.path.join(here, "golden", "make_match_golden.py"))
    mmg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mmg)
    gold = json.load(open(os.path.join(here, "golden", "match_golden.json")))
    # golden digests (made by the reference's own lines) on the mono-frame cases
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    gx = ORBextractor(nf, sf, nl, it, mt)
    kps, desc = gx(fr.cluttered_scene(w, h, 77))
    (grid,) = gx.undistort_grid(mmg.K_TUM1, mmg.D_TUM1)
    cf = dict(xy_un=grid[0], cur_octave=kps["octave"].astype(np.int32), cur_angle=kps["angle"].astype(np.float32), desc=desc,
              cell_start=grid[1], cell_items=grid[2], bounds=grid[3], sf=np.asarray(gx.GetScaleFactors(), np.float32), nlevels=nl)
    done = 0
    for case in mmg.LOCAL_CASES:
        if case[2]:
            continue
        sc = mmg.local_scenario(cf, *case[:3])
        (n, m, _), = gx.search_local_points([dict(cur_frame=0, **{k: sc[k] for k in (
            "in_view", "proj_x", "proj_y", "proj_xr", "scale_level", "view_cos", "mp_desc", "mp_obs", "cur_obs")})], case[3])
        assert mmg.local_digest(sc, n, m) == gold[mmg.local_key(case)]
        done += 1
    assert done == 3
    gx.close()
    # stereo Frame: mvuRight from orbx_stereo_match stays in HBM
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["stereo_kitti"]
    left, right = fr.stereo_pair(w, h, 41)
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=2)
    res = gx.extract_batch([left, right])
    (u_right, _), = gx.stereo_match(gx, 386.1448, 0.5371657, left_frames=[0], right_frames=[1])
    grids = gx.undistort_grid(K_KITTI, D_RECT)
    rng = np.random.default_rng(8)
    lp = mc.make_local_points(rng, grids[0][0], res[0][0]["octave"].astype(np.int32), res[0][1], 3000, nl)
    # expected right coordinates near the measured ones for half of the points, so that both outcomes of :100-105 occur
    tgt_ur = u_right[np.argmin(np.abs(grids[0][0][:, 0][None, :] - lp["proj_x"][:, None]) + np.abs(grids[0][0][:, 1][None, :] - lp["proj_y"][:, None]), axis=1)]
    lp["proj_xr"] = np.where((rng.random(3000) < 0.5) & (tgt_ur > 0), tgt_ur + rng.normal(0, 1.0, 3000), lp["proj_xr"]).astype(np.float32)
    q = dict(cur_frame=0, **lp)
    for th in (1.0, 3.0):
        (n, m, _), = gx.search_local_points([q], th, 0.8, use_stereo=True)
        n0, m0 = local_oracle(q, res[0][0], res[0][1], grids[0], gx.GetScaleFactors(), th, 0.8, u_right)
        assert n == n0 and np.array_equal(m, m0) and n > 300
        n1, _ = local_oracle(q, res[0][0], res[0][1], grids[0], gx.GetScaleFactors(), th, 0.8, None)
        assert n1 != n0                                                    # the mvuRight test changed the outcome
    gx.close()


def test_4k_local_points_with_large_shared_memory():
    """More map points than fit the default 48 KB of dynamic shared memory ((points + keypoint slots) * 4 B): the opt-in
    attribute path of both template instantiations."""
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_4k"]
    K4, D = (3100.0, 3098.0, 1915.5, 1082.25), (0.12, -0.31, 0.0007, -0.0004, 0.09)
    gx = ORBextractor(nf, sf, nl, it, mt)
    kps, desc = gx(fr.cluttered_scene(w, h, 10))
    (grid,) = gx.undistort_grid(K4, D)
    rng = np.random.default_rng(6)
    octv = kps["octave"].astype(np.int32)
    lp = mc.make_local_points(rng, grid[0], octv, desc, 12000, nl)
    q = dict(cur_frame=0, **lp)
    (n, m, _), = gx.search_local_points([q], 3.0)
    n0, m0 = local_oracle(q, kps, desc, grid, gx.GetScaleFactors(), 3.0, 0.8, None)
    assert n == n0 and np.array_equal(m, m0) and n > 1500
    q2 = build_query(rng, 0, kps, desc, grid, K4, 11000, nl, "still")
    (n, m, _), = gx.search_by_projection([q2], K4, 0.0, 0.0, 15.0, True)
    n0, m0 = oracle_result(q2, kps, desc, grid, K4, 0.0, 0.0, gx.GetScaleFactors(), 15.0, True, True, None)
    assert n == n0 and np.array_equal(m, m0)
    gx.close()
