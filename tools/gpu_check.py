#!/usr/bin/env python3
"""Stage-by-stage parity report of the CUDA path against the oracle (run on a GPU box).
Debug aid; the judged parity tests are tests/test_gpu_*.py."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orbslam2_with_quadrics_b200 import ORBextractor, _capi
from orbslam2_with_quadrics_b200.frames import CONFIGS, cluttered_scene, noise_frame, checker_frame, flat_frame
from oracle import orb_oracle


def compare(name, img, nf, sf, nl, it, mt, verbose=True):
    ora = orb_oracle.ORBextractor(nf, sf, nl, it, mt)
    ro = ora(img)
    gx = ORBextractor(nf, sf, nl, it, mt)
    t = time.time()
    kps, desc = gx(img)
    dt = time.time() - t
    ok = True
    rep = {"name": name, "n_gpu": len(kps), "n_ref": ro.n, "first_call_s": round(dt, 3)}
    for l in range(nl):
        pyr = gx.stage_dump(0, l, _capi.STAGE_PYRAMID)
        d = int((pyr != ro.pyramid[l]).sum())
        hp = gx.pyramid(0)[l]
        dh = int((hp != ro.pyramid[l]).sum())
        cand = gx.stage_dump(0, l, _capi.STAGE_CANDIDATES)
        rc = np.stack(ro.candidates[l], axis=1) if len(ro.candidates[l][0]) else np.zeros((0, 3), np.int32)
        c_ok = cand.shape == rc.shape and np.array_equal(cand, rc)
        kept = gx.stage_dump(0, l, _capi.STAGE_KEPT)
        rk = np.stack(ro.kept[l], axis=1) if len(ro.kept[l][0]) else np.zeros((0, 3), np.int32)
        k_ok = kept.shape == rk.shape and np.array_equal(kept, rk)
        ang = gx.stage_dump(0, l, _capi.STAGE_ANGLES)
        a_ok = len(ang) == len(ro.angles[l]) and np.array_equal(ang.view(np.uint32), ro.angles[l].view(np.uint32))
        bl = gx.stage_dump(0, l, _capi.STAGE_BLURRED)
        b_d = int((bl != ro.blurred[l]).sum()) if ro.blurred[l] is not None else -1
        line = "  L%d pyr_diff=%d host_pyr_diff=%d cand %d/%d %s kept %d/%d %s angles %s blur_diff=%d" % (
            l, d, dh, len(cand), len(rc), c_ok, len(kept), len(rk), k_ok, a_ok, b_d)
        if verbose:
            print(line)
        if not c_ok and verbose:
            sa = set(map(tuple, cand.tolist())); sb = set(map(tuple, rc.tolist()))
            print("     cand set equal:", sa == sb, "only_gpu", list(sa - sb)[:5], "only_ref", list(sb - sa)[:5])
            if sa == sb:
                idx = next(i for i in range(len(cand)) if tuple(cand[i]) != tuple(rc[i]))
                print("     first order diff at", idx, cand[idx], rc[idx])
        if not k_ok and c_ok and verbose:
            sa = set(map(tuple, kept.tolist())); sb = set(map(tuple, rk.tolist()))
            print("     kept set equal:", sa == sb, "only_gpu", list(sa - sb)[:5], "only_ref", list(sb - sa)[:5])
        if not a_ok and k_ok and verbose:
            bad = np.nonzero(ang.view(np.uint32) != ro.angles[l].view(np.uint32))[0]
            print("     angle diffs:", len(bad), [(float(ang[i]), float(ro.angles[l][i])) for i in bad[:5]])
        ok &= d == 0 and dh == 0 and c_ok and k_ok and a_ok and b_d <= 0
    same_kp = len(kps) == ro.n and all(np.array_equal(kps[f], ro.keypoints[f]) for f in kps.dtype.names)
    same_desc = len(kps) == ro.n and np.array_equal(desc, ro.descriptors)
    if len(kps) == ro.n and not same_desc:
        bad = np.nonzero((desc != ro.descriptors).any(axis=1))[0]
        print("  desc rows differing:", len(bad), "of", ro.n, bad[:10])
    rep.update(kp_equal=bool(same_kp), desc_equal=bool(same_desc), stages_ok=bool(ok))
    print(json.dumps(rep))
    gx.close()
    return ok and same_kp and same_desc


if __name__ == "__main__":
    names = sys.argv[1:] or ["mono_tum", "stereo_euroc", "stereo_kitti", "rgbd_1080p", "mono_4k"]
    allok = True
    for name in names:
        if name in CONFIGS:
            w, h, nf, sf, nl, it, mt, _ = CONFIGS[name]
            allok &= compare(name, cluttered_scene(w, h, 1234), nf, sf, nl, it, mt)
        elif name == "noise":
            allok &= compare(name, noise_frame(640, 480, 5), 1000, 1.2, 8, 20, 7)
        elif name == "checker":
            allok &= compare(name, checker_frame(640, 480, 5), 1000, 1.2, 8, 20, 7)
        elif name == "flat":
            allok &= compare(name, flat_frame(640, 480), 1000, 1.2, 8, 20, 7)
    print("ALL OK" if allok else "MISMATCH")
    sys.exit(0 if allok else 1)
