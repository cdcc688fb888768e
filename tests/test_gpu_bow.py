"""GPU parity of orbx_compute_bow (Frame::ComputeBoW, reference src/Frame.cc:395-402 -> DBoW2
TemplatedVocabulary::transform) against oracle/bow_oracle.py, which is pinned against the reference's own DBoW2 lines.
Bar: word ids, FeatureVector pairs and their order identical; BowVector values bit-exact as float64 (tolerance 0: the
additions and the division are done in the reference's order with IEEE double operations)."""
import json
import os
import time

import numpy as np
import pytest

from oracle import bow_oracle
from orbslam2_with_quadrics_b200 import ORBextractor, OrbxError, Vocabulary
from orbslam2_with_quadrics_b200 import frames as fr
from orbslam2_with_quadrics_b200 import vocabulary as vc

pytestmark = pytest.mark.gpu


def same(got, want):
    ids, vals, fn, ff = got
    ids0, vals0, fn0, ff0 = want
    assert np.array_equal(ids, ids0) and np.array_equal(vals.view(np.uint64), vals0.view(np.uint64))
    assert np.array_equal(fn, fn0) and np.array_equal(ff, ff0)


@pytest.mark.parametrize("k,L,irregular", [(10, 3, False), (10, 4, True), (4, 6, True)])
def test_batch_matches_oracle(k, L, irregular):
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=4)
    res = gx.extract_batch([fr.cluttered_scene(w, h, 800 + i) for i in range(3)] + [fr.flat_frame(w, h)])
    voc = vc.random_vocabulary(k, L, seed=10 * k + L, irregular=irregular)
    gv = Vocabulary(voc)
    for levelsup in (4, 1, L + 2):
        out = gx.compute_bow(gv, levelsup=levelsup)
        assert len(out) == 4
        for (kps, desc), got in zip(res, out):
            same(got, bow_oracle.transform(voc, desc, levelsup))
        assert len(out[3][0]) == 0 and len(out[3][2]) == 0                        # no keypoints: empty vectors
        assert len(out[0][0]) < len(out[0][2]) or L > 3                            # small vocabulary: words repeat
    out = gx.compute_bow(gv, frames=[2, 0])
    same(out[0], bow_oracle.transform(voc, res[2][1], 4))
    same(out[1], bow_oracle.transform(voc, res[0][1], 4))
    gv.close()
    gx.close()


def test_orbvoc_shape_1080p_4k_timing_and_bad_arguments():
    voc = vc.random_vocabulary(10, 6, seed=1)                                      # ORBvoc's shape: 10^6 words
    gv = Vocabulary(voc)
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["rgbd_1080p"]
    B = 16
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=B, download_pyramid=False)
    imgs = [fr.cluttered_scene(w, h, 3100 + i) for i in range(4)]
    res = gx.extract_batch([imgs[i % 4] for i in range(B)])
    out = gx.compute_bow(gv)
    for f in range(4):
        same(out[f], bow_oracle.transform(voc, res[f][1], 4))
    # measurement of the row: device time per batch, end to end, and the reference's own lines on one host thread
    import torch
    st = torch.cuda.ExternalStream(gx.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    frames = list(range(B))
    for _ in range(3):
        gx.compute_bow_device(gv, frames)
    gx.synchronize()
    K = 20
    e0.record(st)
    for _ in range(K):
        gx.compute_bow_device(gv, frames)
    e1.record(st)
    gx.synchronize()
    dev_ms = e0.elapsed_time(e1) / K
    import ctypes as C
    from orbslam2_with_quadrics_b200 import _capi
    cres = (_capi.OrbxBowResult * B)()
    t0 = time.perf_counter()
    for _ in range(K):                                                             # the C ABI call itself (no numpy copies)
        _capi.check(gx._L.orbx_compute_bow(gx._h, gv._v, B, None, 4, cres), gx._h)
    e2e_ms = (time.perf_counter() - t0) / K * 1e3
    cpu_ms = None
    if bow_oracle.ref_available():
        secs = []
        for f in range(4):
            same(out[f], bow_oracle.ref_transform(voc, res[f][1], 4))
            secs.append(bow_oracle.ref_last_transform_seconds())
        cpu_ms = float(np.median(secs)) * 1e3                                      # the transform() call alone
    rep = {"workload": "ComputeBoW, %d frames of ~%d ORB descriptors, vocabulary k=10 L=6 (1 111 111 nodes, 35.6 MB), levelsup 4" % (
               B, len(res[0][1])), "device_ms_per_batch": dev_ms, "device_us_per_frame": dev_ms / B * 1e3,
           "e2e_ms_per_batch_with_d2h": e2e_ms, "e2e_us_per_frame": e2e_ms / B * 1e3,
           "reference_lines_cpu_ms_per_frame_1_thread": cpu_ms,
           "words_per_frame": [len(o[0]) for o in out[:4]]}
    print(json.dumps(rep))
    if os.path.isdir("gpurun_out"):
        json.dump(rep, open("gpurun_out/r01_compute_bow.json", "w"), indent=1)
    gx.close()
    # 4K: more than 4096 keypoint slots per frame (8192-entry sort)
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_4k"]
    gx = ORBextractor(nf, sf, nl, it, mt)
    with pytest.raises(OrbxError):
        gx.compute_bow(gv, frames=[0])                                             # nothing extracted yet
    kps, desc = gx(fr.cluttered_scene(w, h, 9))
    (got,) = gx.compute_bow(gv)
    same(got, bow_oracle.transform(voc, desc, 4))
    with pytest.raises(OrbxError):
        gx.compute_bow(gv, frames=[1])
    gx.close()
    gv.close()
    bad = vc.random_vocabulary(3, 2, seed=1)
    bad["child_items"] = bad["child_items"].copy()
    bad["child_items"][0] = bad["child_items"][1]                                  # a node listed twice
    with pytest.raises(OrbxError):
        Vocabulary(bad)


# ----------------------------------------------------------------------------- SearchByBoW(KeyFrame*, Frame&) (:159-288)
def test_search_by_bow_matches_oracle_and_golden():
    import importlib.util
    from oracle import match_oracle
    from orbslam2_with_quadrics_b200 import match_cases as mc
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_match_golden", os.path.join(here, "golden", "make_match_golden.py"))
    mmg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mmg)
    gold = json.load(open(os.path.join(here, "golden", "match_golden.json")))
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=3)
    # frame 1 of the batch is the golden scenarios' frame (seed 77): the chain extract -> compute_bow -> search lands on
    # the digests made by the reference's own lines
    res = gx.extract_batch([fr.cluttered_scene(w, h, 5), fr.cluttered_scene(w, h, 77), fr.cluttered_scene(w, h, 6)])
    cf = dict(desc=res[1][1], cur_angle=res[1][0]["angle"].astype(np.float32))
    with pytest.raises(OrbxError):
        gx.search_by_bow([dict(cur_frame=1, kf_desc=np.zeros((4, 32)), kf_valid=[1] * 4, kf_angle=[0.0] * 4, kf_fv_nodes=[1],
                               kf_fv_features=[0])])                                     # no FeatureVector yet
    for case in mmg.BOW_CASES:
        voc, sc = mmg.bow_scenario(cf, case)
        gv = Vocabulary(voc)
        gx.compute_bow(gv, frames=[2, 1], levelsup=case[2])                               # frame 1 sits in slot 1 of this call
        q = dict(cur_frame=1, **{k: sc[k] for k in ("kf_desc", "kf_valid", "kf_angle", "kf_fv_nodes", "kf_fv_features")})
        (n, m), = gx.search_by_bow([q], case[5], case[6])
        assert mmg.bow_digest(sc, n, m) == gold[mmg.bow_key(case)]
        gv.close()
    # a batch of different KeyFrames against different frames, one vocabulary
    voc = vc.random_vocabulary(10, 4, seed=9)
    gv = Vocabulary(voc)
    bows = gx.compute_bow(gv, levelsup=2)
    rng = np.random.default_rng(3)
    qs, scs = [], []
    for f, n_kf in ((2, 1800), (0, 600), (1, 1100)):
        kf = mc.make_keyframe(rng, res[f][1], res[f][0]["angle"].astype(np.float32), n_kf)
        _, _, kn, kfeat = bow_oracle.transform(voc, kf["kf_desc"], 2)
        qs.append(dict(cur_frame=f, kf_fv_nodes=kn, kf_fv_features=kfeat, **kf))
    for ratio, ori in ((0.7, True), (0.75, False), (0.9, True)):
        out = gx.search_by_bow(qs, ratio, ori)
        for q, (n, m) in zip(qs, out):
            f = q["cur_frame"]
            n0, m0 = match_oracle.search_by_bow(q["kf_desc"], q["kf_valid"], q["kf_angle"], q["kf_fv_nodes"], q["kf_fv_features"],
                                                res[f][1], res[f][0]["angle"].astype(np.float32), bows[f][2], bows[f][3], ratio, ori)
            assert n == n0 and np.array_equal(m, m0) and n > 0.15 * len(q["kf_valid"])
    bad = dict(qs[0]); bad["kf_fv_nodes"] = bad["kf_fv_nodes"][::-1].copy()
    with pytest.raises(OrbxError):
        gx.search_by_bow([bad])                                                           # not in map order
    gv.close()
    gx.close()


def test_search_by_bow_timing_report():
    """Measurement of the row: 16 KeyFrame/Frame pairs at 1080p (2000 features each side, 10^6-word vocabulary, nodes at
    level 2), device time per batch (CUDA events, staging upload included), C-ABI end to end, and the reference's own
    lines on one host thread."""
    import ctypes as C
    import torch
    from oracle import match_oracle
    from orbslam2_with_quadrics_b200 import _capi
    from orbslam2_with_quadrics_b200 import match_cases as mc
    voc = vc.random_vocabulary(10, 6, seed=1)
    gv = Vocabulary(voc)
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["rgbd_1080p"]
    B = 16
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=B, download_pyramid=False)
    imgs = [fr.cluttered_scene(w, h, 3100 + i) for i in range(4)]
    res = gx.extract_batch([imgs[i % 4] for i in range(B)])
    bows = gx.compute_bow(gv)
    rng = np.random.default_rng(1)
    qs = []
    for f in range(B):
        kf = mc.make_keyframe(rng, res[f][1], res[f][0]["angle"].astype(np.float32), len(res[f][1]))
        if f < 4:
            _, _, kn, kfeat = bow_oracle.transform(voc, kf["kf_desc"], 4)
            qs.append(dict(cur_frame=f, kf_fv_nodes=kn, kf_fv_features=kfeat, **kf))
        else:
            qs.append(dict(qs[f % 4], cur_frame=f))
    out = gx.search_by_bow(qs, 0.7, True)
    cq, keep = gx._bow_queries(qs)
    cres = (_capi.OrbxProjectionResult * B)()
    st = torch.cuda.ExternalStream(gx.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        _capi.check(gx._L.orbx_search_by_bow_device(gx._h, B, cq, 0.7, 1), gx._h)
    gx.synchronize()
    K = 20
    e0.record(st)
    for _ in range(K):
        _capi.check(gx._L.orbx_search_by_bow_device(gx._h, B, cq, 0.7, 1), gx._h)
    e1.record(st)
    gx.synchronize()
    dev_ms = e0.elapsed_time(e1) / K
    t0 = time.perf_counter()
    for _ in range(K):
        _capi.check(gx._L.orbx_search_by_bow(gx._h, B, cq, 0.7, 1, cres), gx._h)
    e2e_ms = (time.perf_counter() - t0) / K * 1e3
    cpu_ms = None
    if match_oracle.ref_available():
        t0 = time.perf_counter()
        for q, (n, m) in zip(qs[:4], out[:4]):
            f = q["cur_frame"]
            n0, m0 = match_oracle.ref_search_by_bow(q["kf_desc"], q["kf_valid"], q["kf_angle"], q["kf_fv_nodes"], q["kf_fv_features"],
                                                    res[f][1], res[f][0]["angle"].astype(np.float32), bows[f][2], bows[f][3], 0.7, True)
            assert n == n0 and np.array_equal(m, m0)
        cpu_ms = (time.perf_counter() - t0) / 4 * 1e3
    rep = {"workload": "SearchByBoW, %d KeyFrame/Frame pairs, ~%d features per side, %d common nodes" % (
               B, len(res[0][1]), len(set(bows[0][2].tolist()))), "device_ms_per_batch_incl_staging_h2d": dev_ms,
           "device_us_per_pair": dev_ms / B * 1e3, "e2e_ms_per_batch_with_d2h": e2e_ms, "e2e_us_per_pair": e2e_ms / B * 1e3,
           "reference_lines_cpu_ms_per_pair_1_thread_incl_stub_setup": cpu_ms, "nmatches": [n for n, _ in out[:4]]}
    print(json.dumps(rep))
    if os.path.isdir("gpurun_out"):
        json.dump(rep, open("gpurun_out/r01_search_by_bow.json", "w"), indent=1)
    del keep
    gv.close()
    gx.close()
