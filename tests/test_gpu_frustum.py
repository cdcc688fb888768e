"""GPU parity of orbx_is_in_frustum -- Frame::isInFrustum (reference src/Frame.cc:269-325) with MapPoint::PredictScale
(src/MapPoint.cc:402-417), the loop of Tracking::SearchLocalPoints (src/Tracking.cc:1165-1178) -- against
oracle/match_oracle.py (pinned against the reference's own lines) and against the golden digests those lines produced.
Bar: mbTrackInView and mnTrackScaleLevel identical, mTrackProjX / Y / XR and mTrackViewCos identical by their float bits."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle import match_oracle
from orbslam2_with_quadrics_b200 import ORBextractor, OrbxError
from orbslam2_with_quadrics_b200 import frames as fr
from orbslam2_with_quadrics_b200 import match_cases as mc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_match_golden", os.path.join(HERE, "golden", "make_match_golden.py"))
mmg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mmg)
GOLD = json.load(open(os.path.join(HERE, "golden", "match_golden.json")))
K_TUM1, D_TUM1 = mmg.K_TUM1, mmg.D_TUM1
BOUNDS_TUM1 = (-35.0, 690.0, -30.0, 520.0)          # an undistorted image rectangle larger than 640 x 480, as k1 > 0 produces


def same_bits(a, b):
    return all(np.array_equal(np.ascontiguousarray(x).view(np.uint8), np.ascontiguousarray(y).view(np.uint8)) for x, y in zip(a, b))


def query(sc):
    return dict(consider=sc["consider"], world=sc["world"], normal=sc["normal"], min_dist=sc["min_dist"], max_dist=sc["max_dist"], Tcw=sc["Tcw"])


def test_golden_of_the_reference_lines():
    """the digests the reference's own lines produced (tests/golden/match_golden.json), on the frame the goldens were made on"""
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    gx = ORBextractor(nf, sf, nl, it, mt)
    kps, desc = gx(fr.cluttered_scene(w, h, 77))
    (grid,) = gx.undistort_grid(K_TUM1, D_TUM1)
    cf = dict(bounds=grid[3], sf=np.asarray(gx.GetScaleFactors(), np.float32), nlevels=nl)
    for case in mmg.FRUSTUM_CASES:
        sc = mmg.frustum_scenario(cf, *case[:2])
        (r,) = gx.is_in_frustum([query(sc)], sc["K4"], case[3], sc["bounds"], sc["log_scale_factor"], case[2])
        assert mmg.frustum_digest(sc, *r) == GOLD[mmg.frustum_key(case)]
    gx.close()


def test_batch_of_poses_matches_oracle_without_an_extraction():
    """several Frames' local maps in one call, on a handle that has not extracted anything (the call reads no frame state);
    10 levels, a KITTI camera, one empty list, consider = NULL"""
    nl, sf = 10, 1.2
    gx = ORBextractor(4000, sf, nl, 20, 7, max_batch=2)
    K4, bounds = (718.856, 718.856, 607.1928, 185.2157), (0.0, 1241.0, 0.0, 376.0)
    sfs = np.asarray(gx.GetScaleFactors(), np.float32)
    lsf = float(np.float32(np.log(np.float32(sf))))
    rng = np.random.default_rng(8)
    scs = []
    for n in (2500, 0, 700, 4000):
        Tc = mc.pose(rng, scale_r=0.4, t=(1.0, 0.5, -2.0))
        sc = mc.make_frustum_points(rng, K4, bounds, Tc, n, sfs)
        sc["Tcw"] = Tc
        scs.append(sc)
    scs[2]["consider"] = None
    for mbf, cos_limit in ((386.1448, 0.5), (0.0, 0.9)):
        res = gx.is_in_frustum([query(sc) for sc in scs], K4, mbf, bounds, lsf, cos_limit)
        for sc, r in zip(scs, res):
            want = match_oracle.is_in_frustum(sc["consider"], sc["world"], sc["normal"], sc["min_dist"], sc["max_dist"], sc["Tcw"], K4, mbf,
                                              bounds, lsf, nl, cos_limit)
            assert same_bits(r, want)
        assert sum(int(r[0].sum()) for r in res) > 500
        assert set(np.concatenate([r[2][r[0] > 0] for r in res]).tolist()) == set(range(nl))
    gx.close()


@pytest.mark.parametrize("sf,nl", [(1.2, 8), (1.1, 12), (2.0, 5), (1.414, 6)])
def test_predict_scale_at_the_level_boundaries(sf, nl):
    """MapPoint::PredictScale goes through the host's logf; the device only compares the ratio with per-level thresholds found on the
    host.  Points on the optical axis of an identity pose (always in view) with mfMaxDistance / dist swept through every level
    boundary, a few ulps either side of sf^k, plus non-finite and non-positive ratios."""
    gx = ORBextractor(1000, sf, nl, 20, 7)
    lsf = float(np.float32(np.log(np.float32(sf))))
    rng = np.random.default_rng(int(sf * 1000) + nl)
    z = rng.uniform(0.5, 20.0, 6000).astype(np.float32)
    world = np.stack([np.zeros_like(z), np.zeros_like(z), z], 1)
    normal = np.tile(np.float32([0, 0, 1]), (len(z), 1))
    k = rng.integers(-2, nl + 2, len(z))
    target = (np.float32(sf) ** k.astype(np.float32)).astype(np.float32)              # the ratio aimed at
    ulps = rng.integers(-6, 7, len(z))
    ratio = (target.view(np.int32) + ulps.astype(np.int32)).view(np.float32)
    max_dist = (ratio * z).astype(np.float32)
    wide = rng.random(len(z)) < 0.3
    max_dist[wide] = (z[wide] * np.exp(rng.uniform(-2, 4, int(wide.sum())))).astype(np.float32)
    max_dist[:6] = np.float32([0.0, -1.0, np.inf, np.nan, 3.0e38, 1e-42])
    min_dist = np.zeros_like(z)
    q = dict(consider=None, world=world, normal=normal, min_dist=min_dist, max_dist=max_dist, Tcw=np.eye(4, dtype=np.float32))
    K4, bounds = (500.0, 500.0, 320.0, 240.0), (0.0, 640.0, 0.0, 480.0)
    (r,) = gx.is_in_frustum([q], K4, 40.0, bounds, lsf, 0.5)
    want = match_oracle.is_in_frustum(None, world, normal, min_dist, max_dist, q["Tcw"], K4, 40.0, bounds, lsf, nl, 0.5)
    assert same_bits(r, want)
    inside = r[0] > 0                                    # dist <= 1.2 * max_dist
    assert inside.sum() > 3000 and set(r[2][inside].tolist()) == set(range(nl))
    gx.close()


def test_feeds_search_local_points():
    """Tracking::SearchLocalPoints as a chain on the device results: extract -> undistort / grid -> orbx_is_in_frustum ->
    orbx_search_local_points, against the same chain of the two oracles"""
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    gx = ORBextractor(nf, sf, nl, it, mt)
    kps, desc = gx(fr.cluttered_scene(w, h, 91))
    (grid,) = gx.undistort_grid(K_TUM1, D_TUM1)
    xy, start, items, bounds = grid
    sfs = np.asarray(gx.GetScaleFactors(), np.float32)
    octave = kps["octave"].astype(np.int32)
    rng = np.random.default_rng(14)
    Tc = mc.pose(rng)
    kf = mc.make_reloc_keyframe(rng, xy, octave, kps["angle"].astype(np.float32), desc, K_TUM1, Tc, 1500, sfs)
    T64 = np.asarray(Tc, np.float64)
    Ow = -T64[:3, :3].T @ T64[:3, 3]
    ray = kf["world"].astype(np.float64) - Ow
    ray /= np.linalg.norm(ray, axis=1, keepdims=True)
    normal = (ray + rng.normal(0, 0.35, ray.shape)).astype(np.float32)
    normal /= np.linalg.norm(normal, axis=1, keepdims=True).astype(np.float32)
    consider = (kf["valid"] == 1).astype(np.uint8)
    lsf = mmg.kf_log_scale(dict(sf=sfs))
    fq = dict(consider=consider, world=kf["world"], normal=normal, min_dist=kf["min_dist"], max_dist=kf["max_dist"], Tcw=Tc)
    (r,) = gx.is_in_frustum([fq], K_TUM1, 40.0, bounds, lsf, 0.5)
    want = match_oracle.is_in_frustum(consider, kf["world"], normal, kf["min_dist"], kf["max_dist"], Tc, K_TUM1, 40.0, bounds, lsf, nl, 0.5)
    assert same_bits(r, want) and 300 < int(r[0].sum()) < 1400
    in_view, proj, level, vcos = r
    mp_obs = rng.integers(1, 6, len(in_view)).astype(np.int32)
    cur_obs = np.where(rng.random(len(desc)) < 0.3, rng.integers(0, 4, len(desc)), -1).astype(np.int32)
    lq = dict(cur_frame=0, in_view=in_view, proj_x=proj[:, 0], proj_y=proj[:, 1], proj_xr=proj[:, 2], scale_level=level, view_cos=vcos,
              mp_desc=kf["mp_desc"], mp_obs=mp_obs, cur_obs=cur_obs)
    (n, m, _), = gx.search_local_points([lq], 3.0, 0.8)
    n2, m2 = match_oracle.search_local_points(want[0], want[1][:, 0], want[1][:, 1], want[1][:, 2], want[2], want[3], kf["mp_desc"], mp_obs,
                                              xy, octave, desc, None, cur_obs, start, items, bounds, sfs, 3.0, 0.8)
    assert n == n2 > 100 and np.array_equal(m, m2)
    gx.close()


def test_bad_arguments():
    gx = ORBextractor(1000, 1.2, 8, 20, 7)
    sc = mc.make_frustum_points(np.random.default_rng(1), K_TUM1, BOUNDS_TUM1, np.eye(4, dtype=np.float32), 10, np.float32(1.2) ** np.arange(8, dtype=np.float32))
    sc["Tcw"] = np.eye(4, dtype=np.float32)
    for lsf in (0.0, -0.2, float("nan"), float("inf")):
        with pytest.raises(OrbxError):
            gx.is_in_frustum([query(sc)], K_TUM1, 40.0, BOUNDS_TUM1, lsf)
    (r,) = gx.is_in_frustum([dict(consider=None, world=np.zeros((0, 3), np.float32), normal=np.zeros((0, 3), np.float32),
                                  min_dist=np.zeros(0, np.float32), max_dist=np.zeros(0, np.float32), Tcw=np.eye(4, dtype=np.float32))],
                            K_TUM1, 40.0, BOUNDS_TUM1, 0.18)
    assert len(r[0]) == 0
    gx.close()
