"""Parity tests proper: the CUDA path, called through the C ABI (ctypes -> liborbx.so), against the
oracle on the same seeded frames.  Bar: pyramid bytes, ordered (x, y, size, response, octave) and
descriptors bit-exact; angles compared bit-exactly too (the north_star tolerance is 1e-3 degrees --
ANGLE_TOL below -- and is asserted separately so a regression to 'within tolerance' is visible)."""
import hashlib
import importlib.util
import json
import os
import threading

import numpy as np
import pytest

from oracle import orb_oracle
from orbslam2_with_quadrics_b200 import ORBextractor, OrbxError, _capi
from orbslam2_with_quadrics_b200 import frames as fr

pytestmark = pytest.mark.gpu
ANGLE_TOL = 1e-3     # degrees (BASELINE.json north_star)

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)
GOLD = json.load(open(os.path.join(HERE, "golden", "orb_golden.json")))


def assert_same(kps, desc, ro):
    assert len(kps) == ro.n
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kps[f], ro.keypoints[f]), f
    if ro.n:
        assert float(np.abs(kps["angle"] - ro.keypoints["angle"]).max()) <= ANGLE_TOL
    assert np.array_equal(kps["angle"].view(np.uint32), ro.keypoints["angle"].view(np.uint32))
    mism = int((desc != ro.descriptors).any(axis=1).sum()) if ro.n else 0
    assert mism == 0, "descriptor rows differing: %d of %d" % (mism, ro.n)


def stage_parity(gx, ro, nl):
    for l in range(nl):
        assert np.array_equal(gx.stage_dump(0, l, _capi.STAGE_PYRAMID), ro.pyramid[l]), ("pyramid", l)
        assert np.array_equal(gx.pyramid(0)[l], ro.pyramid[l]), ("host pyramid", l)
        rc = np.stack(ro.candidates[l], axis=1).reshape(-1, 3)
        assert np.array_equal(gx.stage_dump(0, l, _capi.STAGE_CANDIDATES), rc), ("candidates", l)
        rk = np.stack(ro.kept[l], axis=1).reshape(-1, 3)
        assert np.array_equal(gx.stage_dump(0, l, _capi.STAGE_KEPT), rk), ("kept", l)
        assert np.array_equal(gx.stage_dump(0, l, _capi.STAGE_ANGLES).view(np.uint32), ro.angles[l].view(np.uint32)), ("angles", l)
        if ro.blurred[l] is not None:
            assert np.array_equal(gx.stage_dump(0, l, _capi.STAGE_BLURRED), ro.blurred[l]), ("blur", l)


@pytest.mark.parametrize("name", list(fr.CONFIGS))
@pytest.mark.parametrize("seed", [1234, 2234, 4321])
def test_configs_stage_by_stage(name, seed):
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS[name]
    if name == "mono_4k" and seed == 4321:
        pytest.skip("two 4K seeds are enough")
    img = fr.cluttered_scene(w, h, seed)
    ro = orb_oracle.ORBextractor(nf, sf, nl, it, mt)(img)
    gx = ORBextractor(nf, sf, nl, it, mt)
    kps, desc = gx(img)
    assert_same(kps, desc, ro)
    stage_parity(gx, ro, nl)
    gx.close()


@pytest.mark.parametrize("key", sorted(k for k in GOLD if not k.startswith("_")))
def test_golden_digests(key):
    cases = {k: (img, args) for k, img, args in mg.cases()}
    img, args = cases[key]
    gx = ORBextractor(*args)
    kps, desc = gx(img)
    d = mg.digest(kps, desc, gx.pyramid(0))
    for f, v in d.items():
        assert GOLD[key][f] == v, f
    gx.close()


@pytest.mark.parametrize("kind", ["noise", "checker", "flat"])
def test_adversarial_frames(kind):
    img = {"noise": fr.noise_frame(640, 480, 5), "checker": fr.checker_frame(640, 480, 5), "flat": fr.flat_frame(640, 480)}[kind]
    ro = orb_oracle.ORBextractor(1000, 1.2, 8, 20, 7)(img)
    gx = ORBextractor(1000, 1.2, 8, 20, 7)
    kps, desc = gx(img)
    assert_same(kps, desc, ro)
    stage_parity(gx, ro, 8)
    if kind == "flat":
        assert len(kps) == 0 and desc.shape == (0, 32)        # descriptors released (:1064-1065)
    gx.close()


@pytest.mark.parametrize("shape,args", [((333, 517), (500, 1.2, 6, 20, 7)), ((480, 640), (1500, 1.5, 4, 12, 5)),
                                         ((200, 900), (700, 1.1, 10, 20, 7)), ((131, 257), (200, 1.2, 3, 20, 7)),
                                         ((600, 400), (400, 1.3, 5, 30, 10)), ((480, 640), (500, 2.0, 3, 20, 7)),
                                         ((480, 752), (800, 1.8, 4, 20, 7))])
def test_odd_shapes_strides_and_parameters(shape, args):
    h, w = shape
    big = fr.cluttered_scene(w + 64, h + 32, 900 + h)
    view = big[7:7 + h, 13:13 + w]                              # non-contiguous rows: "any step"
    ro = orb_oracle.ORBextractor(*args)(np.ascontiguousarray(view))
    gx = ORBextractor(*args)
    kps, desc = gx(view)
    assert_same(kps, desc, ro)
    stage_parity(gx, ro, args[2])
    gx.close()


def test_batch_equals_single_and_is_deterministic():
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["stereo_euroc"]
    imgs = [fr.cluttered_scene(w, h, 50 + i) for i in range(5)] + [fr.flat_frame(w, h), fr.noise_frame(w, h, 3)]
    gb = ORBextractor(nf, sf, nl, it, mt, max_batch=len(imgs))
    g1 = ORBextractor(nf, sf, nl, it, mt)
    batch = gb.extract_batch(imgs)
    again = gb.extract_batch(imgs)
    for i, im in enumerate(imgs):
        k1, d1 = g1(im)
        for k, d in (batch[i], again[i]):
            assert np.array_equal(k.view(np.uint8), k1.view(np.uint8)) and np.array_equal(d, d1)
        ro = orb_oracle.ORBextractor(nf, sf, nl, it, mt)(im)
        assert_same(k1, d1, ro)
        assert np.array_equal(gb.pyramid(i)[3], ro.pyramid[3])
    gb.close(); g1.close()


def test_large_batch_tile_resize_path_equals_single_frames():
    """A 40 x 1080p batch runs the large pyramid levels through pyr_resize8_tile_kernel (TMA source tiles, full 16-row
    strips) in sub-batches that start at frame > 0; a single frame takes pyr_resize8_kernel's short strips.  Keypoints,
    descriptors and every pyramid plane (border included) must be the same bytes."""
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["rgbd_1080p"]
    base = [fr.cluttered_scene(w, h, 700 + i) for i in range(4)]
    imgs = [base[i % 4] for i in range(40)]
    gb = ORBextractor(nf, sf, nl, it, mt, max_batch=len(imgs), download_pyramid=True)
    g1 = ORBextractor(nf, sf, nl, it, mt, download_pyramid=True)
    batch = gb.extract_batch(imgs)
    single = []
    for im in base:
        k1, d1 = g1(im)
        single.append((k1.copy(), d1.copy(), [g1.pyramid(0)[l].copy() for l in range(nl)]))
    for i in (0, 1, 9, 18, 31, 32, 35, 39):                   # frames of every sub-batch
        k1, d1, p1 = single[i % 4]
        assert np.array_equal(batch[i][0].view(np.uint8), k1.view(np.uint8)) and np.array_equal(batch[i][1], d1)
        pb = gb.pyramid(i)
        for l in range(nl):
            assert np.array_equal(pb[l], p1[l]), (i, l)
    gb.close(); g1.close()


def test_stereo_two_handles_two_threads():
    """src/Frame.cc:78-81: left and right extractors run concurrently on two host threads."""
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["stereo_kitti"]
    left, right = fr.cluttered_scene(w, h, 1), fr.cluttered_scene(w, h, 2)
    exs = [ORBextractor(nf, sf, nl, it, mt), ORBextractor(nf, sf, nl, it, mt)]
    out = [None, None]

    def run(i, im):
        for _ in range(5):
            out[i] = exs[i](im)
    th = [threading.Thread(target=run, args=(i, im)) for i, im in enumerate((left, right))]
    [t.start() for t in th]
    [t.join() for t in th]
    for i, im in enumerate((left, right)):
        assert_same(out[i][0], out[i][1], orb_oracle.ORBextractor(nf, sf, nl, it, mt)(im))
    assert exs[0].stream != exs[1].stream
    [e.close() for e in exs]


def test_device_resident_path_matches_host_path():
    import torch
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    imgs = [fr.cluttered_scene(w, h, 70 + i) for i in range(4)]
    pitch = 640
    dev = torch.from_numpy(np.stack(imgs)).cuda()
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=4, download_pyramid=False)
    gx.extract_device(dev.data_ptr(), 4, w, h, pitch, h * pitch)
    res = gx.fetch_results(4)
    for im, (k, d) in zip(imgs, res):
        assert_same(k, d, orb_oracle.ORBextractor(nf, sf, nl, it, mt)(im))
    assert 5 <= gx.launch_count <= nl + 6          # pyramid (2 chained launches, or one per level) + FAST (levels 0-1 / 2+, strips / big cells: <= 4) + octree + describe
    with pytest.raises(OrbxError):
        gx.pyramid(0)                      # not downloaded: must fail, not return stale data
    gx.close()


def test_api_conformance():
    gx = ORBextractor(1000, 1.2, 8, 20, 7)
    p = orb_oracle.OrbParams(1000, 1.2, 8, 20, 7)
    assert gx.GetLevels() == 8 and np.float32(gx.GetScaleFactor()) == np.float32(1.2)
    assert np.array_equal(gx.GetScaleFactors().view(np.uint32), p.mvScaleFactor.view(np.uint32))
    assert np.array_equal(gx.GetInverseScaleFactors().view(np.uint32), p.mvInvScaleFactor.view(np.uint32))
    assert np.array_equal(gx.GetScaleSigmaSquares().view(np.uint32), p.mvLevelSigma2.view(np.uint32))
    assert np.array_equal(gx.GetInverseScaleSigmaSquares().view(np.uint32), p.mvInvLevelSigma2.view(np.uint32))
    q, u = gx.level_quotas()
    assert q == p.mnFeaturesPerLevel and u == p.umax
    assert gx.level_sizes(640, 480) == p.level_sizes(640, 480)
    assert gx.algorithmic_bytes(640, 480) == 6564354
    assert gx(np.zeros((0, 0), np.uint8)) is None                       # empty image: outputs untouched
    with pytest.raises(OrbxError) as e:
        gx(np.zeros((60, 60), np.uint8))                                 # level < 62 px (App. B-7b)
    assert e.value.status == _capi.ERR_BAD_GEOMETRY
    with pytest.raises(OrbxError) as e:
        gx(np.zeros((100, 5000), np.uint8))
    assert e.value.status == _capi.ERR_BAD_GEOMETRY
    with pytest.raises(OrbxError) as e:
        gx(np.zeros((600, 300), np.uint8))      # width/height < 0.5: nIni = 0, the reference divides by zero (App. B-7)
    assert e.value.status == _capi.ERR_BAD_GEOMETRY
    img = fr.cluttered_scene(640, 480, 5)
    k1, d1 = gx(img)
    lvl0 = gx.mvImagePyramid[0]
    assert lvl0.shape == (480, 640) and np.array_equal(lvl0, img)        # pyramid lifetime: valid after return
    k2, d2 = gx(fr.cluttered_scene(752, 480, 6))                        # geometry change on the same object
    assert gx.mvImagePyramid[0].shape == (480, 752)
    k3, d3 = gx(img)
    assert np.array_equal(k1.view(np.uint8), k3.view(np.uint8)) and np.array_equal(d1, d3)
    gx.close()


def test_candidate_overflow_is_reported_not_truncated():
    gx = ORBextractor(1000, 1.2, 8, 20, 7, candidate_divisor=100000)     # capacity ~1024 per level
    with pytest.raises(OrbxError) as e:
        gx(fr.noise_frame(640, 480, 5))
    assert e.value.status == _capi.ERR_CANDIDATE_OVERFLOW
    kps, desc = gx(fr.cluttered_scene(640, 480, 1234))                   # the handle stays usable
    assert len(kps) == 1007
    gx.close()


def test_full_size_properties():
    """Size-independent properties at BASELINE's largest configs: determinism, quota, bounds, octave order,
    translation of the scene by a multiple of the cell grid is not assumed -- only invariants are."""
    for name in ("rgbd_1080p", "mono_4k"):
        w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS[name]
        img = fr.cluttered_scene(w, h, 999)
        gx = ORBextractor(nf, sf, nl, it, mt, download_pyramid=False)
        k1, d1 = gx(img)
        k2, d2 = gx(img)
        assert hashlib.sha256(k1.tobytes() + d1.tobytes()).digest() == hashlib.sha256(k2.tobytes() + d2.tobytes()).digest()
        assert nf <= len(k1) <= nf + 3 * nl
        assert np.all(np.diff(k1["octave"]) >= 0)                        # levels concatenated in order
        sfs = gx.GetScaleFactors()
        lv = gx.level_sizes(w, h)
        for l in range(nl):
            m = k1["octave"] == l
            x, y = k1["x"][m] / sfs[l], k1["y"][m] / sfs[l]
            assert x.min() >= 18.99 and y.min() >= 18.99 and x.max() <= lv[l][0] - 19.99 and y.max() <= lv[l][1] - 19.99
            assert np.all(k1["size"][m] == np.float32(int(np.float32(31) * sfs[l])))
        assert np.all((k1["angle"] >= 0) & (k1["angle"] < 360)) and np.all(k1["response"] >= 6)
        assert len(np.unique(np.stack([k1["x"], k1["y"], k1["octave"].astype(np.float32)], 1), axis=0)) == len(k1)
        gx.close()


@pytest.mark.parametrize("w,h,pad", [(1241, 376, 0), (1241, 376, 7), (752, 480, 0), (640, 480, 32)])
def test_pinned_inputs_all_copy_paths(w, h, pad):
    """Host frames in pinned memory take the direct 1-D H2D path (one merged copy when contiguous), with the
    caller's own row stride -- including strides that are not a multiple of 16 (unaligned level-0 copy)."""
    import torch
    n, stride = 5, w + pad
    host = torch.zeros((n, h, stride), dtype=torch.uint8).pin_memory()
    imgs = [fr.cluttered_scene(w, h, 300 + i) for i in range(n)]
    hnp = host.numpy()
    hnp[:] = 255                                     # padding bytes must never leak into the result
    for i, im in enumerate(imgs):
        hnp[i, :, :w] = im
    gx = ORBextractor(800, 1.2, 8, 20, 7, max_batch=n)
    out = gx.extract_batch([hnp[i, :, :w] for i in range(n)])                 # contiguous frames: merged copy
    out_rev = gx.extract_batch([hnp[i, :, :w] for i in reversed(range(n))])   # non-contiguous order: one copy per frame
    for i, im in enumerate(imgs):
        ro = orb_oracle.ORBextractor(800, 1.2, 8, 20, 7)(im)
        assert_same(out[i][0], out[i][1], ro)
        assert_same(out_rev[n - 1 - i][0], out_rev[n - 1 - i][1], ro)
    assert np.array_equal(gx.pyramid(0)[0], orb_oracle.ORBextractor(800, 1.2, 8, 20, 7)(imgs[n - 1]).pyramid[0])
    gx.close()


def test_large_batch_through_both_paths_matches_single():
    """32 frames: orbx_extract_batch (4 sub-batches, 2 kernel streams) and orbx_extract_device (2 half-batches)
    give exactly what 32 single-frame calls give."""
    import torch
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    imgs = [fr.cluttered_scene(w, h, 500 + i) for i in range(32)]
    g1 = ORBextractor(nf, sf, nl, it, mt)
    single = [g1(im) for im in imgs]
    gb = ORBextractor(nf, sf, nl, it, mt, max_batch=32, download_pyramid=False)
    batch = gb.extract_batch(imgs)
    dev = torch.from_numpy(np.stack(imgs)).cuda()
    gb.extract_device(dev.data_ptr(), 32, w, h, w, w * h)
    devres = gb.fetch_results(32)
    for i in range(32):
        for k, d in (batch[i], devres[i]):
            assert np.array_equal(k.view(np.uint8), single[i][0].view(np.uint8)) and np.array_equal(d, single[i][1])
    g1.close(); gb.close()


def test_graph_replay_is_identical_and_survives_geometry_changes():
    """The launch sequence is captured into a CUDA graph the second time it is seen and replayed afterwards: calls 1
    (eager), 2 (capture) and 3+ (replay) must return the same bytes, for a batch and for single frames, and a change of
    image size in between must drop the graphs of the old geometry."""
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    imgs = [fr.cluttered_scene(w, h, 70 + i) for i in range(8)]
    small = [fr.cluttered_scene(333, 257, 90 + i) for i in range(8)]
    ref = [orb_oracle.ORBextractor(nf, sf, nl, it, mt)(im) for im in imgs]
    gx = ORBextractor(nf, sf, nl, it, mt, max_batch=8)
    for rep in range(4):
        for (kps, desc), ro in zip(gx.extract_batch(imgs), ref):
            assert_same(kps, desc, ro)
    first_small = gx.extract_batch(small)
    for rep in range(3):
        for (k0, d0), (k1, d1) in zip(first_small, gx.extract_batch(small)):
            assert np.array_equal(k0, k1) and np.array_equal(d0, d1)
    for rep in range(3):                                      # back to the first geometry, single frames this time
        for im, ro in zip(imgs[:3], ref[:3]):
            kps, desc = gx(im)
            assert_same(kps, desc, ro)
    gx.close()


@pytest.mark.parametrize("offset,pad", [(0, 0), (0, 5), (3, 0), (16, 1)])
def test_level0_border_split_all_width_residues(offset, pad):
    """pyr_level0_kernel copies interior 16-byte vectors in one set of CTAs and gives the vectors that touch the reflected
    border to another (lane = (row, border vector)); which vectors are which depends on width mod 16 and on the alignment of
    the caller's rows.  Every residue, with aligned (128-bit loads) and unaligned (funnel-shift) source rows, device-resident
    input: the padded level-0 plane -- border included -- and level 1 made from it are the oracle's bytes."""
    import torch
    h = 80
    for w in range(100, 116):
        img = fr.cluttered_scene(w, h, 500 + w)
        pitch = (w + 15) // 16 * 16 + pad
        buf = torch.zeros(offset + h * pitch + 64, dtype=torch.uint8)
        flat = buf.numpy()
        flat[:] = 255                                             # bytes between the rows must never reach the plane
        rows = flat[offset:offset + h * pitch].reshape(h, pitch)
        rows[:, :w] = img
        dev = buf.cuda()
        gx = ORBextractor(100, 1.2, 2, 20, 7, max_batch=1, download_pyramid=False)
        gx.extract_device(dev.data_ptr() + offset, 1, w, h, pitch, h * pitch)
        ro = orb_oracle.ORBextractor(100, 1.2, 2, 20, 7)(img)
        for l in range(2):
            assert np.array_equal(gx.stage_dump(0, l, _capi.STAGE_PYRAMID), ro.pyramid[l]), (w, l)
        (k, d), = gx.fetch_results(1)
        assert_same(k, d, ro)
        gx.close()
