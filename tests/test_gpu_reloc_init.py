"""GPU parity of the two remaining Frame-side matchers against oracle/match_oracle.py (pinned against the reference's own
lines) and against the golden digests those lines produced:
  orbx_search_by_projection_kf     ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist), reference
                                   src/ORBmatcher.cc:1472-1599 (Tracking::Relocalization)
  orbx_search_for_initialization   ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:405-520 (MonocularInitialization)
Bar: identical integer results (and, for the initialisation matcher, identical vbPrevMatched floats)."""
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


def frame_view(kps, desc, grid, sf, nl):
    xy, start, items, bounds = grid
    return dict(xy_un=xy, cur_octave=kps["octave"].astype(np.int32), cur_angle=kps["angle"].astype(np.float32), desc=desc,
                cell_start=start, cell_items=items, bounds=bounds, sf=np.asarray(sf, np.float32), nlevels=nl)


def kf_query(sc, cf, frame=0):
    lsf = mmg.kf_log_scale(cf)
    inr, pred = match_oracle.kf_prepare(sc["valid"], sc["world"], sc["min_dist"], sc["max_dist"], sc["Tcw_cur"], lsf, cf["nlevels"])
    q = dict(cur_frame=frame, search=((sc["valid"] == 1) & (inr > 0)).astype(np.uint8), world=sc["world"], pred_level=pred,
             mp_desc=sc["mp_desc"], kf_angle=sc["kf_angle"], Tcw_cur=sc["Tcw_cur"], cur_held=sc["cur_held"].astype(np.int32))
    return q, inr, pred


def kf_oracle(sc, inr, pred, th, orb_dist, check):
    args = {k: v for k, v in sc.items() if k not in ("min_dist", "max_dist")}
    return match_oracle.search_by_projection_kf(in_range=inr, pred_level=pred, th=th, orb_dist=orb_dist, check_orientation=check, **args)


def test_keyframe_projection_golden_of_the_reference_lines():
    """extract -> undistort/grid -> orbx_search_by_projection_kf lands on the digests the reference's own lines produced."""
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    gx = ORBextractor(nf, sf, nl, it, mt)
    kps, desc = gx(fr.cluttered_scene(w, h, 77))
    (grid,) = gx.undistort_grid(K_TUM1, D_TUM1)
    cf = frame_view(kps, desc, grid, gx.GetScaleFactors(), nl)
    for case in mmg.KF_CASES:
        sc = mmg.kf_scenario(cf, *case[:2])
        q, inr, pred = kf_query(sc, cf)
        (n, m, rounds), = gx.search_by_projection_kf([q], K_TUM1, case[2], case[3], case[4])
        assert mmg.kf_digest(sc, n, m, inr, pred) == GOLD[mmg.kf_key(case)]
        assert rounds >= 2 or case[1] < 100                # collisions needed more than one fixed-point round
    gx.close()


@pytest.mark.parametrize("list_cap", [None, "2"])
def test_keyframe_projection_batch_matches_oracle(list_cap, monkeypatch):
    if list_cap:
        monkeypatch.setenv("ORBX_SP_LIST_CAP", list_cap)       # lists that do not fit: the full search inside the resolve kernel
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["stereo_kitti"]
    K4, D = (718.856, 718.856, 607.1928, 185.2157), (0.0, 0.0, 0.0, 0.0)
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=4)
    res = gx.extract_batch([fr.cluttered_scene(w, h, 300 + i) for i in range(3)] + [fr.flat_frame(w, h)])
    grids = gx.undistort_grid(K4, D)
    rng = np.random.default_rng(21)
    scs, qs, staged = [], [], []
    for f, n in ((0, 1800), (1, 400), (2, 2500)):
        cf = frame_view(res[f][0], res[f][1], grids[f], gx.GetScaleFactors(), nl)
        Tc = mc.pose(rng)
        kf = mc.make_reloc_keyframe(rng, cf["xy_un"], cf["cur_octave"], cf["cur_angle"], cf["desc"], K4, Tc, n, cf["sf"])
        sc = dict(**kf, Tcw_cur=Tc, xy_un=cf["xy_un"], cur_octave=cf["cur_octave"], cur_angle=cf["cur_angle"], desc=cf["desc"],
                  cell_start=cf["cell_start"], cell_items=cf["cell_items"], bounds=cf["bounds"], K4=K4, sf=cf["sf"])
        q, inr, pred = kf_query(sc, cf, f)
        scs.append(sc); qs.append(q); staged.append((inr, pred))
    for th, dist, check in ((10.0, 100, True), (3.0, 64, True), (10.0, 100, False)):
        out = gx.search_by_projection_kf(qs, K4, th, dist, check)
        for sc, (inr, pred), (n, m, _) in zip(scs, staged, out):
            n0, m0 = kf_oracle(sc, inr, pred, th, dist, check)
            assert n == n0 and np.array_equal(m, m0)
            assert n > 0.1 * len(sc["valid"])
    # a query against the frame without keypoints, one with garbage levels on points that are not searched, cur_held = NULL
    empty = dict(qs[0]); empty.update(cur_frame=3)
    junk = dict(qs[1]); junk["pred_level"] = np.where(junk["search"] > 0, junk["pred_level"], -2 ** 31).astype(np.int32)
    free = dict(qs[1]); free["cur_held"] = None
    out = gx.search_by_projection_kf([empty, junk, free], K4, 10.0, 100, True)
    assert out[0][0] == 0 and len(out[0][1]) == 0
    n0, m0 = kf_oracle(scs[1], *staged[1], 10.0, 100, True)
    assert out[1][0] == n0 and np.array_equal(out[1][1], m0)
    sc_free = dict(scs[1]); sc_free["cur_held"] = np.zeros_like(scs[1]["cur_held"])
    n1, m1 = kf_oracle(sc_free, *staged[1], 10.0, 100, True)
    assert out[2][0] == n1 and np.array_equal(out[2][1], m1) and n1 > n0
    bad = dict(qs[0]); bad["pred_level"] = np.full_like(qs[0]["pred_level"], nl)
    with pytest.raises(OrbxError):
        gx.search_by_projection_kf([bad], K4, 10.0, 100, True)              # a searched point with a level outside the pyramid
    with pytest.raises(OrbxError):
        gx.search_by_projection_kf([qs[0]], K4, 10.0, 256, True)            # ORBdist >= 256 would index with bestIdx2 = -1 (:1532)
    gx.close()


def init_oracle(sc, nnratio, check, window):
    return match_oracle.search_for_initialization(nnratio=nnratio, check_orientation=check, window=window, **sc)


def init_query(sc, frame=0):
    return dict(cur_frame=frame, octave1=sc["octave1"], angle1=sc["angle1"], desc1=sc["desc1"], prev_matched=sc["prev_matched"])


def test_search_for_initialization_golden_of_the_reference_lines():
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    gx = ORBextractor(nf, sf, nl, it, mt)
    kps, desc = gx(fr.cluttered_scene(w, h, 77))
    (grid,) = gx.undistort_grid(K_TUM1, D_TUM1)
    cf = frame_view(kps, desc, grid, gx.GetScaleFactors(), nl)
    for case in mmg.INIT_CASES:
        sc = mmg.init_scenario(cf, *case[:2])
        (n, m, prev), = gx.search_for_initialization([init_query(sc)], case[2], case[3], case[4])
        assert mmg.init_digest(sc, n, m, prev) == GOLD[mmg.init_key(case)]
    gx.close()


@pytest.mark.parametrize("list_cap", [None, "8"])
def test_search_for_initialization_batch_matches_oracle(list_cap, monkeypatch):
    if list_cap:
        monkeypatch.setenv("ORBX_INIT_LIST_CAP", list_cap)     # lists that do not fit: the window is re-scanned by the resolve kernel
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    gx = ORBextractor(2 * nf, sf, nl, it, mt, max_batch=3)     # MonocularInitialization's extractor asks for 2 x nFeatures (src/Tracking.cc:127)
    res = gx.extract_batch([fr.cluttered_scene(w, h, 500 + i) for i in range(2)] + [fr.flat_frame(w, h)])
    grids = gx.undistort_grid(K_TUM1, D_TUM1)
    rng = np.random.default_rng(31)
    scs = []
    for f, n1 in ((0, 2000), (1, 700)):
        cf = frame_view(res[f][0], res[f][1], grids[f], gx.GetScaleFactors(), nl)
        f1 = mc.make_initial_frame(rng, cf["xy_un"], cf["cur_octave"], cf["cur_angle"], cf["desc"], n1)
        scs.append(dict(**f1, xy_un2=cf["xy_un"], octave2=cf["cur_octave"], angle2=cf["cur_angle"], desc2=cf["desc"],
                        cell_start=cf["cell_start"], cell_items=cf["cell_items"], bounds=cf["bounds"]))
    for ratio, check, window in ((0.9, True, 100), (0.9, False, 100), (0.7, True, 30)):
        out = gx.search_for_initialization([init_query(scs[0], 0), init_query(scs[1], 1)], ratio, check, window)
        for sc, (n, m, prev) in zip(scs, out):
            n0, m0, p0 = init_oracle(sc, ratio, check, window)
            assert n == n0 and np.array_equal(m, m0) and np.array_equal(prev, p0)
            assert n > 0.08 * len(sc["octave1"])
    # against the frame without keypoints; an F1 without keypoints
    none = {k: v[:0] for k, v in init_query(scs[1], 1).items() if isinstance(v, np.ndarray)}
    none["cur_frame"] = 1
    out = gx.search_for_initialization([init_query(scs[0], 2), none], 0.9, True, 100)
    assert out[0][0] == 0 and np.all(out[0][1] == -1) and np.array_equal(out[0][2], scs[0]["prev_matched"])
    assert out[1][0] == 0 and len(out[1][1]) == 0
    gx.close()
