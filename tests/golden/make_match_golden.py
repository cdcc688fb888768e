#!/usr/bin/env python3
"""Regenerates tests/golden/match_golden.json: nmatches and the SHA-256 of the per-keypoint assignment that the
REFERENCE'S OWN LINES of ORBmatcher::SearchByProjection(Frame&, const Frame&) (oracle/_ref/libstereoref.so, built from
/root/reference/src/ORBmatcher.cc:1328-1470 + src/Frame.cc:327-380 by oracle/build_stereo_ref.sh) produce on the seeded
scenarios below; the restatement (oracle/match_oracle.py) and the CUDA path are compared against these digests where the
reference sources are absent.

    python tests/golden/make_match_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import frame_oracle, match_oracle, orb_oracle      # noqa: E402
from orbslam2_with_quadrics_b200 import frames as fr           # noqa: E402
from orbslam2_with_quadrics_b200 import match_cases as mc      # noqa: E402

K_TUM1 = (517.306408, 516.469215, 318.643040, 255.313989)      # Examples/Monocular/TUM1.yaml:9-17
D_TUM1 = (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)
# (seed, LastFrame.N, motion, stereo Frame (mvuRight set), th, bMono)
CASES = [(1, 900, "still", False, 15.0, True), (2, 1500, "forward", True, 7.0, False), (3, 1500, "backward", True, 7.0, False),
         (4, 1200, "still", True, 14.0, False), (5, 600, "forward", False, 30.0, True), (6, 40, "still", False, 15.0, True)]


def current_frame():
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    ex = orb_oracle.ORBextractor(nf, sf, nl, it, mt)
    res = ex(fr.cluttered_scene(w, h, 77))
    xy, start, items, b = frame_oracle.undistort_and_grid(res.keypoints, K_TUM1, D_TUM1, w, h)
    return dict(xy_un=xy, cur_octave=res.keypoints["octave"].astype(np.int32), cur_angle=res.keypoints["angle"].astype(np.float32),
                desc=res.descriptors, cell_start=start, cell_items=items, bounds=b, sf=np.asarray(ex.GetScaleFactors(), np.float32),
                nlevels=nl)


def scenario(cf, seed, n_last, motion, stereo):
    rng = np.random.default_rng(seed)
    Tc = mc.pose(rng)
    tz = {"still": 0.0, "forward": 0.6, "backward": -0.6}[motion]
    Tl = mc.pose(rng, t=(0.0, 0.0, tz))
    last = mc.make_last_frame(rng, cf["xy_un"], cf["cur_octave"], cf["cur_angle"], cf["desc"], K_TUM1, Tc, n_last, cf["nlevels"])
    u_right = None
    if stereo:
        u_right = np.where(rng.random(len(cf["desc"])) < 0.6, cf["xy_un"][:, 0] - rng.uniform(2, 40, len(cf["desc"])), -1).astype(np.float32)
    return dict(**last, Tcw_cur=Tc, Tcw_last=Tl, xy_un=cf["xy_un"], cur_octave=cf["cur_octave"], cur_angle=cf["cur_angle"],
                desc=cf["desc"], u_right=u_right, cell_start=cf["cell_start"], cell_items=cf["cell_items"], bounds=cf["bounds"],
                K4=K_TUM1, mbf=40.0, mb=0.08, sf=cf["sf"])


# SearchLocalPoints' matcher: (seed, vpMapPoints.size(), stereo Frame, th)  -- th = 1 / 3 (RGB-D) / 5 (after relocalisation),
# src/Tracking.cc:1185-1191
LOCAL_CASES = [(1, 1200, False, 1.0), (2, 2000, True, 3.0), (3, 800, False, 5.0), (4, 1500, True, 1.0), (5, 30, False, 3.0)]


def local_scenario(cf, seed, n_points, stereo):
    rng = np.random.default_rng(1000 + seed)
    lp = mc.make_local_points(rng, cf["xy_un"], cf["cur_octave"], cf["desc"], n_points, cf["nlevels"])
    u_right = None
    if stereo:
        u_right = np.where(rng.random(len(cf["desc"])) < 0.6, cf["xy_un"][:, 0] - rng.uniform(2, 40, len(cf["desc"])), -1).astype(np.float32)
    return dict(**lp, xy_un=cf["xy_un"], cur_octave=cf["cur_octave"], desc=cf["desc"], u_right=u_right, cell_start=cf["cell_start"],
                cell_items=cf["cell_items"], bounds=cf["bounds"], sf=cf["sf"])


def local_digest(sc, n, m):
    ins = hashlib.sha256()
    for k in ("in_view", "proj_x", "proj_y", "proj_xr", "scale_level", "view_cos", "mp_desc", "mp_obs", "cur_obs"):
        ins.update(np.ascontiguousarray(sc[k]).tobytes())
    return {"nmatches": int(n), "set": int((m >= 0).sum()), "inputs_sha256": ins.hexdigest(),
            "match_sha256": hashlib.sha256(np.ascontiguousarray(m, np.int32).tobytes()).hexdigest()}


def local_key(case):
    return "local/seed%d/n%d/%s/th%g" % (case[0], case[1], "stereo" if case[2] else "mono-frame", case[3])


# SearchByBoW(KeyFrame*, Frame&): (vocabulary k, L, levelsup, seed, pKF->N, mfNNratio, mbCheckOrientation) -- ratio 0.7 in
# TrackReferenceKeyFrame (src/Tracking.cc:771), 0.75 in Relocalization (:1376)
BOW_CASES = [(10, 3, 1, 1, 900, 0.7, True), (10, 4, 2, 2, 1500, 0.75, True), (10, 4, 2, 3, 1200, 0.7, False), (4, 6, 4, 4, 700, 0.9, True),
             (10, 4, 6, 5, 500, 0.7, True)]


def bow_scenario(cf, case):
    from oracle import bow_oracle
    from orbslam2_with_quadrics_b200 import vocabulary as vc
    k, L, levelsup, seed, n_kf = case[:5]
    voc = vc.random_vocabulary(k, L, seed=seed)
    rng = np.random.default_rng(2000 + seed)
    kf = mc.make_keyframe(rng, cf["desc"], cf["cur_angle"], n_kf)
    _, _, kn, kfeat = bow_oracle.transform(voc, kf["kf_desc"], levelsup)
    _, _, fn, ff = bow_oracle.transform(voc, cf["desc"], levelsup)
    return voc, dict(**kf, kf_fv_nodes=kn, kf_fv_features=kfeat, f_desc=cf["desc"], f_angle=cf["cur_angle"], f_fv_nodes=fn,
                     f_fv_features=ff)


def bow_digest(sc, n, m):
    ins = hashlib.sha256()
    for k in ("kf_desc", "kf_valid", "kf_angle", "kf_fv_nodes", "kf_fv_features", "f_fv_nodes", "f_fv_features"):
        ins.update(np.ascontiguousarray(sc[k]).tobytes())
    return {"nmatches": int(n), "inputs_sha256": ins.hexdigest(),
            "match_sha256": hashlib.sha256(np.ascontiguousarray(m, np.int32).tobytes()).hexdigest()}


def bow_key(case):
    return "bow/k%d/L%d/up%d/seed%d/n%d/ratio%g/ori%d" % (case[0], case[1], case[2], case[3], case[4], case[5], int(case[6]))


# Relocalization's SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist): (seed, pKF->N, th, ORBdist, mbCheckOrientation)
# -- th 10 / ORBdist 100, then th 3 / ORBdist 64 (src/Tracking.cc:1452, :1466)
KF_CASES = [(1, 1200, 10.0, 100, True), (2, 1500, 3.0, 64, True), (3, 900, 10.0, 100, False), (4, 40, 3.0, 64, True),
            (5, 2000, 10.0, 100, True)]


def kf_scenario(cf, seed, n_points):
    rng = np.random.default_rng(3000 + seed)
    Tc = mc.pose(rng)
    kf = mc.make_reloc_keyframe(rng, cf["xy_un"], cf["cur_octave"], cf["cur_angle"], cf["desc"], K_TUM1, Tc, n_points, cf["sf"])
    return dict(**kf, Tcw_cur=Tc, xy_un=cf["xy_un"], cur_octave=cf["cur_octave"], cur_angle=cf["cur_angle"], desc=cf["desc"],
                cell_start=cf["cell_start"], cell_items=cf["cell_items"], bounds=cf["bounds"], K4=K_TUM1, sf=cf["sf"])


def kf_log_scale(cf):
    return float(np.float32(np.log(np.float32(cf["sf"][1]))))      # Frame::mfLogScaleFactor = log(mfScaleFactor) (src/Frame.cc:69)


def kf_digest(sc, n, m, in_range, pred):
    ins = hashlib.sha256()
    for k in ("valid", "world", "mp_desc", "min_dist", "max_dist", "kf_angle", "cur_held", "Tcw_cur"):
        ins.update(np.ascontiguousarray(sc[k]).tobytes())
    return {"nmatches": int(n), "set": int((m >= 0).sum()), "inputs_sha256": ins.hexdigest(),
            "match_sha256": hashlib.sha256(np.ascontiguousarray(m, np.int32).tobytes()).hexdigest(),
            "staged_sha256": hashlib.sha256(np.ascontiguousarray(in_range, np.uint8).tobytes()
                                            + np.ascontiguousarray(pred, np.int32).tobytes()).hexdigest()}


def kf_key(case):
    return "kf/seed%d/n%d/th%g/dist%d/ori%d" % (case[0], case[1], case[2], case[3], int(case[4]))


# MonocularInitialization's SearchForInitialization: (seed, F1.N, mfNNratio, mbCheckOrientation, windowSize) -- ORBmatcher(0.9, true),
# window 100 (src/Tracking.cc:597-600)
INIT_CASES = [(1, 1000, 0.9, True, 100), (2, 1000, 0.9, False, 100), (3, 600, 0.7, True, 50), (4, 30, 0.9, True, 100),
              (5, 1000, 0.95, True, 200)]


def init_scenario(cf, seed, n1):
    rng = np.random.default_rng(4000 + seed)
    f1 = mc.make_initial_frame(rng, cf["xy_un"], cf["cur_octave"], cf["cur_angle"], cf["desc"], n1)
    return dict(**f1, xy_un2=cf["xy_un"], octave2=cf["cur_octave"], angle2=cf["cur_angle"], desc2=cf["desc"],
                cell_start=cf["cell_start"], cell_items=cf["cell_items"], bounds=cf["bounds"])


def init_digest(sc, n, m, prev):
    ins = hashlib.sha256()
    for k in ("xy_un1", "octave1", "angle1", "desc1", "prev_matched"):
        ins.update(np.ascontiguousarray(sc[k]).tobytes())
    return {"nmatches": int(n), "inputs_sha256": ins.hexdigest(),
            "match_sha256": hashlib.sha256(np.ascontiguousarray(m, np.int32).tobytes()).hexdigest(),
            "prev_sha256": hashlib.sha256(np.ascontiguousarray(prev, np.float32).tobytes()).hexdigest()}


def init_key(case):
    return "init/seed%d/n%d/ratio%g/ori%d/win%d" % (case[0], case[1], case[2], int(case[3]), case[4])


def digest(sc, n, m):
    ins = hashlib.sha256()
    for k in ("world", "mp_desc", "mp_obs", "outlier", "last_octave", "last_angle", "Tcw_cur", "Tcw_last"):
        ins.update(np.ascontiguousarray(sc[k]).tobytes())
    return {"nmatches": int(n), "holders": int((m >= 0).sum()), "inputs_sha256": ins.hexdigest(),
            "match_sha256": hashlib.sha256(np.ascontiguousarray(m, np.int32).tobytes()).hexdigest()}


def key(case, check):
    return "seed%d/n%d/%s/%s/th%g/%s/ori%d" % (case[0], case[1], case[2], "stereo" if case[3] else "mono-frame", case[4],
                                              "bMono" if case[5] else "bStereo", int(check))


# Frame::isInFrustum over the local map (Tracking::SearchLocalPoints, src/Tracking.cc:1165-1178, viewingCosLimit 0.5):
# (seed, mvpLocalMapPoints.size(), viewingCosLimit, mbf)
FRUSTUM_CASES = [(1, 3000, 0.5, 40.0), (2, 5000, 0.5, 0.0), (3, 800, 0.8, 386.1448), (4, 50, 0.5, 40.0), (5, 6000, 0.2, 47.9)]


def frustum_scenario(cf, seed, n_points):
    rng = np.random.default_rng(5000 + seed)
    Tc = mc.pose(rng, scale_r=0.3, t=(0.3, -0.2, 0.5))
    pts = mc.make_frustum_points(rng, K_TUM1, cf["bounds"], Tc, n_points, cf["sf"])
    return dict(**pts, Tcw=Tc, K4=K_TUM1, bounds=cf["bounds"], log_scale_factor=kf_log_scale(cf), nlevels=cf["nlevels"])


def frustum_digest(sc, in_view, proj, level, vcos):
    ins = hashlib.sha256()
    for k in ("consider", "world", "normal", "min_dist", "max_dist", "Tcw"):
        ins.update(np.ascontiguousarray(sc[k]).tobytes())
    out = hashlib.sha256()
    out.update(np.ascontiguousarray(in_view, np.uint8).tobytes())
    out.update(np.ascontiguousarray(proj, np.float32).tobytes())
    out.update(np.ascontiguousarray(level, np.int32).tobytes())
    out.update(np.ascontiguousarray(vcos, np.float32).tobytes())
    return {"in_view": int(np.asarray(in_view).sum()), "inputs_sha256": ins.hexdigest(), "fields_sha256": out.hexdigest()}


def frustum_key(case):
    return "frustum/seed%d/n%d/cos%g/mbf%g" % case


if __name__ == "__main__":
    from oracle import stereo_oracle
    stereo_oracle.ref_build()
    assert match_oracle.ref_available(), "the reference's own lines are needed to make the golden file"
    cf = current_frame()
    out = {}
    for case in CASES:
        sc = scenario(cf, *case[:4])
        for check in (True, False):
            n, m = match_oracle.ref_search_by_projection(th=case[4], mono=case[5], check_orientation=check, **sc)
            out[key(case, check)] = digest(sc, n, m)
            print(key(case, check), out[key(case, check)]["nmatches"], out[key(case, check)]["holders"])
    for case in LOCAL_CASES:
        sc = local_scenario(cf, *case[:3])
        n, m = match_oracle.ref_search_local_points(th=case[3], **sc)
        out[local_key(case)] = local_digest(sc, n, m)
        print(local_key(case), n, int((m >= 0).sum()))
    for case in BOW_CASES:
        _, sc = bow_scenario(cf, case)
        n, m = match_oracle.ref_search_by_bow(nnratio=case[5], check_orientation=case[6], **sc)
        out[bow_key(case)] = bow_digest(sc, n, m)
        print(bow_key(case), n, len(set(sc["f_fv_nodes"].tolist())))
    for case in KF_CASES:
        sc = kf_scenario(cf, *case[:2])
        n, m, inr, pred = match_oracle.ref_search_by_projection_kf(log_scale_factor=kf_log_scale(cf), th=case[2], orb_dist=case[3],
                                                                   check_orientation=case[4], **sc)
        out[kf_key(case)] = kf_digest(sc, n, m, inr, pred)
        print(kf_key(case), n, int(inr.sum()))
    for case in INIT_CASES:
        sc = init_scenario(cf, *case[:2])
        n, m, prev = match_oracle.ref_search_for_initialization(sf=cf["sf"], nnratio=case[2], check_orientation=case[3], window=case[4], **sc)
        out[init_key(case)] = init_digest(sc, n, m, prev)
        print(init_key(case), n)
    for case in FRUSTUM_CASES:
        sc = frustum_scenario(cf, *case[:2])
        r = match_oracle.ref_is_in_frustum(cos_limit=case[2], mbf=case[3], **sc)
        out[frustum_key(case)] = frustum_digest(sc, *r)
        print(frustum_key(case), out[frustum_key(case)]["in_view"], sorted(set(r[2][r[0] > 0].tolist())))
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "match_golden.json"), "w"), indent=1, sort_keys=True)
