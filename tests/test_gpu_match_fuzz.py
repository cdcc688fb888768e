"""Seeded GPU fuzz of the device-side descriptor consumers against their oracles: random image sizes, extractor settings,
cameras, numbers of map points, thresholds, ratios, motion, vocabulary shapes and candidate-list capacities.  Bar: every
integer result identical, BowVector values bit-exact."""
import numpy as np
import pytest

from oracle import bow_oracle, match_oracle
from orbslam2_with_quadrics_b200 import ORBextractor, Vocabulary
from orbslam2_with_quadrics_b200 import frames as fr
from orbslam2_with_quadrics_b200 import match_cases as mc
from orbslam2_with_quadrics_b200 import vocabulary as vc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", list(range(10)))
def test_consumers_fuzz(seed, monkeypatch):
    rng = np.random.default_rng(9000 + seed)
    w = int(rng.integers(200, 900))
    h = int(rng.integers(max(160, w // 2 + 8), min(700, 2 * w - 8)))
    nf = int(rng.integers(300, 1800))
    nl = int(rng.integers(3, 9))
    sf = float(rng.choice([1.2, 1.15, 1.3, 1.5]))
    while min(w, h) / sf ** (nl - 1) < 70:
        nl -= 1
    if rng.random() < 0.4:
        monkeypatch.setenv("ORBX_SP_LIST_CAP", str(int(rng.integers(1, 6))))
    K4 = (float(rng.uniform(0.5, 1.2) * w), float(rng.uniform(0.5, 1.2) * w), w / 2 + float(rng.normal(0, 5)), h / 2 + float(rng.normal(0, 5)))
    D = (0.0, 0.0, 0.0, 0.0) if rng.random() < 0.3 else (float(rng.normal(0, 0.1)), float(rng.normal(0, 0.1)), float(rng.normal(0, 1e-3)),
                                                           float(rng.normal(0, 1e-3)), float(rng.normal(0, 0.05)))
    gx = ORBextractor(nf, sf, nl, 20, 7, max_batch=2)
    res = gx.extract_batch([fr.cluttered_scene(w, h, 100 + seed), fr.cluttered_scene(w, h, 200 + seed)])
    grids = gx.undistort_grid(K4, D)
    sfs = gx.GetScaleFactors()
    for f in (0, 1):
        kps, desc = res[f]
        if len(kps) < 20:
            continue
        xy, start, items, bounds = grids[f]
        octv, ang = kps["octave"].astype(np.int32), kps["angle"].astype(np.float32)
        # --- SearchByProjection(CurrentFrame, LastFrame)
        mono = bool(rng.random() < 0.5)
        Tc = mc.pose(rng)
        Tl = mc.pose(rng, t=(0.0, 0.0, float(rng.choice([0.0, 0.7, -0.7]))))
        nL = int(rng.integers(1, 2 * len(kps)))
        last = mc.make_last_frame(rng, xy, octv, ang, desc, K4, Tc, nL, nl, dup_frac=float(rng.uniform(0, 0.7)), flip_bits=int(rng.integers(2, 40)))
        th = float(rng.choice([7.0, 15.0, 30.0, 60.0]))
        check = bool(rng.random() < 0.7)
        (n, m, _), = gx.search_by_projection([dict(cur_frame=f, Tcw_cur=Tc, Tcw_last=Tl, **last)], K4, 30.0, 0.1, th, mono, check)
        n0, m0 = match_oracle.search_by_projection(last["world"], last["mp_desc"], last["mp_obs"], last["outlier"], last["last_octave"],
                                                   last["last_angle"], Tc, Tl, xy, octv, ang, desc, None, start, items, bounds, K4, 30.0,
                                                   0.1, sfs, th, mono, check)
        assert n == n0 and np.array_equal(m, m0), ("projection", seed, f)
        # --- SearchByProjection(F, vpMapPoints, th)
        nP = int(rng.integers(1, 3 * len(kps)))
        lp = mc.make_local_points(rng, xy, octv, desc, nP, nl, dup_frac=float(rng.uniform(0, 0.7)), flip_bits=int(rng.integers(2, 40)),
                                  held_frac=float(rng.uniform(0, 0.8)))
        th2, ratio = float(rng.choice([1.0, 3.0, 5.0])), float(rng.choice([0.6, 0.8, 0.9]))
        (n, m, _), = gx.search_local_points([dict(cur_frame=f, **lp)], th2, ratio)
        n0, m0 = match_oracle.search_local_points(lp["in_view"], lp["proj_x"], lp["proj_y"], lp["proj_xr"], lp["scale_level"], lp["view_cos"],
                                                  lp["mp_desc"], lp["mp_obs"], xy, octv, desc, None, lp["cur_obs"], start, items, bounds,
                                                  sfs, th2, ratio)
        assert n == n0 and np.array_equal(m, m0), ("local", seed, f)
    # --- ComputeBoW + SearchByBoW
    k, L = int(rng.integers(2, 11)), int(rng.integers(2, 6))
    voc = vc.random_vocabulary(k, L, seed=seed, irregular=bool(rng.random() < 0.5))
    gv = Vocabulary(voc)
    levelsup = int(rng.integers(0, L + 2))
    bows = gx.compute_bow(gv, levelsup=levelsup)
    for f in (0, 1):
        kps, desc = res[f]
        want = bow_oracle.transform(voc, desc, levelsup)
        got = bows[f]
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint64), want[1].view(np.uint64)), ("bow", seed, f)
        assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3]), ("featvec", seed, f)
        if len(kps) < 20:
            continue
        ang = kps["angle"].astype(np.float32)
        kf = mc.make_keyframe(rng, desc, ang, int(rng.integers(1, 2 * len(kps))), flip_bits=int(rng.integers(2, 30)))
        _, _, kn, kfeat = bow_oracle.transform(voc, kf["kf_desc"], levelsup)
        ratio, check = float(rng.choice([0.6, 0.7, 0.75, 0.9])), bool(rng.random() < 0.7)
        (n, m), = gx.search_by_bow([dict(cur_frame=f, kf_fv_nodes=kn, kf_fv_features=kfeat, **kf)], ratio, check)
        n0, m0 = match_oracle.search_by_bow(kf["kf_desc"], kf["kf_valid"], kf["kf_angle"], kn, kfeat, desc, ang, want[2], want[3], ratio, check)
        assert n == n0 and np.array_equal(m, m0), ("search_by_bow", seed, f)
    gv.close()
    gx.close()
