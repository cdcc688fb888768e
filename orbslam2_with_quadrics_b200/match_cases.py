"""Seeded LastFrame / CurrentFrame scenarios for ORBmatcher::SearchByProjection(Frame&, const Frame&) parity tests.

A scenario takes the current frame's keypoints (positions, octaves, angles, descriptors, grid) as given and builds a
LastFrame whose map points re-project onto them: world point = back-projection of a current keypoint at a random depth
through the current pose, plus pixel-level noise; the map point's descriptor = the keypoint's with random bit flips.
Contention (several map points landing on one keypoint, with and without observations), outliers, missing map points,
points behind the camera / outside the image and unrelated descriptors are mixed in."""
import numpy as np

f32 = np.float32


def rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def pose(rng, scale_r=0.05, t=(0.0, 0.0, 0.0)):
    T = np.eye(4)
    T[:3, :3] = rot(*(rng.normal(0, scale_r, 3)))
    T[:3, 3] = np.asarray(t) + rng.normal(0, 0.05, 3)
    return T.astype(f32)


def make_last_frame(rng, xy_un, cur_octave, cur_angle, desc, K4, Tcw_cur, n_last, nlevels, dup_frac=0.3, flip_bits=25,
                    angle_noise=8.0, angle_outlier_frac=0.15):
    nC = len(desc)
    fx, fy, cx, cy = [float(v) for v in K4]
    Tc = np.asarray(Tcw_cur, np.float64).reshape(4, 4)
    Rwc, twc = Tc[:3, :3].T, -Tc[:3, :3].T @ Tc[:3, 3]
    # which current keypoint every last-frame point is aimed at: a mix of distinct targets and deliberate duplicates
    tgt = rng.integers(0, nC, n_last)
    ndup = int(dup_frac * n_last)
    tgt[rng.integers(0, n_last, ndup)] = tgt[rng.integers(0, n_last, ndup)]
    z = rng.uniform(0.8, 12.0, n_last)
    px = xy_un[tgt, 0].astype(np.float64) + rng.normal(0, 2.0, n_last)
    py = xy_un[tgt, 1].astype(np.float64) + rng.normal(0, 2.0, n_last)
    Xc = np.stack([(px - cx) / fx * z, (py - cy) / fy * z, z], 1)
    behind = rng.random(n_last) < 0.03
    Xc[behind] *= -1.0
    far = rng.random(n_last) < 0.03
    Xc[far, 0] += 50.0 * z[far]
    world = (Xc @ Rwc.T + twc).astype(f32)
    mp_desc = desc[tgt].copy()
    flips = rng.integers(0, 256, (n_last, flip_bits))
    for k in range(flip_bits):
        on = rng.random(n_last) < 0.6
        mp_desc[np.arange(n_last)[on], flips[on, k] // 8] ^= (1 << (flips[on, k] % 8)).astype(np.uint8)
    unrelated = rng.random(n_last) < 0.1
    mp_desc[unrelated] = rng.integers(0, 256, (int(unrelated.sum()), 32), dtype=np.uint8)
    mp_obs = rng.choice([-1, 0, 0, 1, 2, 5], n_last).astype(np.int32)
    outlier = (rng.random(n_last) < 0.05).astype(np.uint8)
    last_octave = np.clip(cur_octave[tgt] + rng.choice([-2, -1, 0, 0, 0, 1, 2], n_last), 0, nlevels - 1).astype(np.int32)
    last_angle = (cur_angle[tgt] + rng.normal(0, angle_noise, n_last)).astype(f32)
    wild = rng.random(n_last) < angle_outlier_frac
    last_angle[wild] = rng.uniform(0, 360, int(wild.sum())).astype(f32)
    last_angle = np.mod(last_angle, f32(360.0)).astype(f32)
    return dict(world=world, mp_desc=mp_desc, mp_obs=mp_obs, outlier=outlier, last_octave=last_octave, last_angle=last_angle)


def grid_from_points(xy_un, bounds):
    """Frame::AssignFeaturesToGrid / PosInGrid (src/Frame.cc:230-245, :382-392) for a synthetic keypoint set (tests that
    do not go through orbx_undistort_grid)."""
    mnx, mxx, mny, mxy = [f32(b) for b in bounds]
    winv = f32(f32(64) / f32(mxx - mnx))
    hinv = f32(f32(48) / f32(mxy - mny))
    cells = [[] for _ in range(64 * 48)]
    for i in range(len(xy_un)):
        gx = int(np.round(f32(f32(xy_un[i, 0] - mnx) * winv)))      # roundf on non-ties; ties do not occur for these inputs
        gy = int(np.round(f32(f32(xy_un[i, 1] - mny) * hinv)))
        if gx < 0 or gx >= 64 or gy < 0 or gy >= 48:
            continue
        cells[gx * 48 + gy].append(i)
    start = np.zeros(64 * 48 + 1, np.int32)
    start[1:] = np.cumsum([len(c) for c in cells])
    items = np.array([i for c in cells for i in c], np.int32)
    return start, items


def make_local_points(rng, xy_un, cur_octave, desc, n_points, nlevels, dup_frac=0.3, flip_bits=25, held_frac=0.35):
    """Map points of the local map as Frame::isInFrustum leaves them for ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th)
    (src/ORBmatcher.cc:45-129): projections near current keypoints (with duplicates that collide), predicted levels around the
    keypoint's octave, viewing cosines on both sides of 0.998, descriptors = the keypoint's with flipped bits (near-duplicates
    make the second-best ratio test bite), and the map points the frame already holds from the motion-model step."""
    nC = len(desc)
    tgt = rng.integers(0, nC, n_points)
    ndup = int(dup_frac * n_points)
    tgt[rng.integers(0, n_points, ndup)] = tgt[rng.integers(0, n_points, ndup)]
    proj_x = (xy_un[tgt, 0] + rng.normal(0, 1.5, n_points)).astype(f32)
    proj_y = (xy_un[tgt, 1] + rng.normal(0, 1.5, n_points)).astype(f32)
    proj_xr = (proj_x - rng.uniform(1, 40, n_points)).astype(f32)
    scale_level = np.clip(cur_octave[tgt] + rng.choice([0, 0, 0, 1, 1, -1], n_points), 0, nlevels - 1).astype(np.int32)
    view_cos = np.where(rng.random(n_points) < 0.5, rng.uniform(0.9981, 1.0, n_points), rng.uniform(0.5, 0.998, n_points)).astype(f32)
    mp_desc = desc[tgt].copy()
    flips = rng.integers(0, 256, (n_points, flip_bits))
    for k in range(flip_bits):
        on = rng.random(n_points) < 0.6
        mp_desc[np.arange(n_points)[on], flips[on, k] // 8] ^= (1 << (flips[on, k] % 8)).astype(np.uint8)
    unrelated = rng.random(n_points) < 0.1
    mp_desc[unrelated] = rng.integers(0, 256, (int(unrelated.sum()), 32), dtype=np.uint8)
    mp_obs = rng.choice([0, 1, 2, 3, 7], n_points).astype(np.int32)
    in_view = (rng.random(n_points) < 0.9).astype(np.uint8)
    cur_obs = np.where(rng.random(nC) < held_frac, rng.choice([0, 1, 4], nC), -1).astype(np.int32)
    return dict(in_view=in_view, proj_x=proj_x, proj_y=proj_y, proj_xr=proj_xr, scale_level=scale_level, view_cos=view_cos,
                mp_desc=mp_desc, mp_obs=mp_obs, cur_obs=cur_obs)


def make_keyframe(rng, f_desc, f_angle, n_kf, flip_bits=12, angle_noise=6.0, angle_outlier_frac=0.15):
    """A reference KeyFrame for ORBmatcher::SearchByBoW(KeyFrame*, Frame&) (src/ORBmatcher.cc:159-288) that really matches the
    frame: its descriptors are the frame's with a few flipped bits (so that most land in the same vocabulary node and under
    TH_LOW), in shuffled order, with duplicates (two KeyFrame features competing for one frame feature), unrelated ones,
    features without / with bad map points, and angles that agree up to noise except for a fraction of outliers."""
    nF = len(f_desc)
    src = rng.integers(0, nF, n_kf)
    ndup = n_kf // 5
    src[rng.integers(0, n_kf, ndup)] = src[rng.integers(0, n_kf, ndup)]
    kf_desc = f_desc[src].copy()
    flips = rng.integers(0, 256, (n_kf, flip_bits))
    for k in range(flip_bits):
        on = rng.random(n_kf) < 0.5
        kf_desc[np.arange(n_kf)[on], flips[on, k] // 8] ^= (1 << (flips[on, k] % 8)).astype(np.uint8)
    unrelated = rng.random(n_kf) < 0.1
    kf_desc[unrelated] = rng.integers(0, 256, (int(unrelated.sum()), 32), dtype=np.uint8)
    kf_valid = rng.choice([0, 1, 1, 1, 1, 1, 2], n_kf).astype(np.uint8)
    kf_angle = (f_angle[src] + rng.normal(0, angle_noise, n_kf)).astype(f32)
    wild = rng.random(n_kf) < angle_outlier_frac
    kf_angle[wild] = rng.uniform(0, 360, int(wild.sum())).astype(f32)
    kf_angle = np.mod(kf_angle, f32(360.0)).astype(f32)
    return dict(kf_desc=kf_desc, kf_valid=kf_valid, kf_angle=kf_angle)


def make_reloc_keyframe(rng, xy_un, cur_octave, cur_angle, desc, K4, Tcw_cur, n_points, sf, dup_frac=0.3, flip_bits=25,
                        angle_noise=8.0, angle_outlier_frac=0.15, held_frac=0.3):
    """A candidate KeyFrame for ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist)
    (src/ORBmatcher.cc:1472-1599, Tracking::Relocalization): map points back-projected from current keypoints (with collisions,
    points behind the camera / outside the image), distance-invariance ranges built so that the predicted level lands around the
    keypoint's octave (some out of range), bad / already-found / missing points, and the keypoints the frame already holds."""
    nC = len(desc)
    fx, fy, cx, cy = [float(v) for v in K4]
    Tc = np.asarray(Tcw_cur, np.float64).reshape(4, 4)
    Rwc, twc = Tc[:3, :3].T, -Tc[:3, :3].T @ Tc[:3, 3]
    tgt = rng.integers(0, nC, n_points)
    ndup = int(dup_frac * n_points)
    tgt[rng.integers(0, n_points, ndup)] = tgt[rng.integers(0, n_points, ndup)]
    z = rng.uniform(0.8, 12.0, n_points)
    px = xy_un[tgt, 0].astype(np.float64) + rng.normal(0, 2.0, n_points)
    py = xy_un[tgt, 1].astype(np.float64) + rng.normal(0, 2.0, n_points)
    Xc = np.stack([(px - cx) / fx * z, (py - cy) / fy * z, z], 1)
    behind = rng.random(n_points) < 0.03
    Xc[behind] *= -1.0
    far = rng.random(n_points) < 0.03
    Xc[far, 0] += 50.0 * z[far]
    world = (Xc @ Rwc.T + twc).astype(f32)
    dist = np.linalg.norm(world.astype(np.float64) - twc, axis=1)
    # mfMaxDistance = dist * scale^(level + jitter): PredictScale = ceil(log(max / dist) / log(scale)) ~ the target level
    lvl = np.clip(cur_octave[tgt] + rng.choice([-1, 0, 0, 0, 1], n_points), 0, len(sf) - 1)
    jitter = rng.uniform(-0.9, 0.0, n_points)
    max_dist = (dist * np.power(float(sf[1]), lvl + jitter)).astype(f32)
    min_dist = (max_dist / f32(sf[-1])).astype(f32)
    out_of_range = rng.random(n_points) < 0.05
    max_dist[out_of_range] = (dist[out_of_range] * 0.5).astype(f32)
    mp_desc = desc[tgt].copy()
    flips = rng.integers(0, 256, (n_points, flip_bits))
    for k in range(flip_bits):
        on = rng.random(n_points) < 0.6
        mp_desc[np.arange(n_points)[on], flips[on, k] // 8] ^= (1 << (flips[on, k] % 8)).astype(np.uint8)
    unrelated = rng.random(n_points) < 0.1
    mp_desc[unrelated] = rng.integers(0, 256, (int(unrelated.sum()), 32), dtype=np.uint8)
    valid = rng.choice([0, 1, 1, 1, 1, 1, 1, 2, 3], n_points).astype(np.uint8)
    kf_angle = (cur_angle[tgt] + rng.normal(0, angle_noise, n_points)).astype(f32)
    wild = rng.random(n_points) < angle_outlier_frac
    kf_angle[wild] = rng.uniform(0, 360, int(wild.sum())).astype(f32)
    kf_angle = np.mod(kf_angle, f32(360.0)).astype(f32)
    cur_held = (rng.random(nC) < held_frac).astype(np.uint8)
    return dict(valid=valid, world=world, mp_desc=mp_desc, min_dist=min_dist, max_dist=max_dist, kf_angle=kf_angle, cur_held=cur_held)


def make_initial_frame(rng, xy_un2, octave2, angle2, desc2, n1, flip_bits=10, shift=(6.0, -4.0), angle_noise=6.0,
                       angle_outlier_frac=0.15, dup_frac=0.25):
    """An initial frame F1 for ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:405-520): its level-0 keypoints are the
    second frame's moved by a small image motion (so the 100-px window holds the true partner and many distractors), descriptors
    with a few flipped bits, duplicates (two F1 keypoints competing for one F2 keypoint, the later one stealing it when it is
    strictly closer), unrelated descriptors, keypoints of higher octaves (skipped) and angles with outliers."""
    n2 = len(desc2)
    lvl0 = np.nonzero(np.asarray(octave2) == 0)[0]
    src = lvl0[rng.integers(0, len(lvl0), n1)] if len(lvl0) else rng.integers(0, n2, n1)
    ndup = int(dup_frac * n1)
    src[rng.integers(0, n1, ndup)] = src[rng.integers(0, n1, ndup)]
    xy1 = (xy_un2[src] + np.asarray(shift, f32) + rng.normal(0, 3.0, (n1, 2))).astype(f32)
    octave1 = np.where(rng.random(n1) < 0.8, 0, rng.integers(1, 4, n1)).astype(np.int32)
    desc1 = desc2[src].copy()
    flips = rng.integers(0, 256, (n1, flip_bits))
    for k in range(flip_bits):
        on = rng.random(n1) < 0.5
        desc1[np.arange(n1)[on], flips[on, k] // 8] ^= (1 << (flips[on, k] % 8)).astype(np.uint8)
    unrelated = rng.random(n1) < 0.1
    desc1[unrelated] = rng.integers(0, 256, (int(unrelated.sum()), 32), dtype=np.uint8)
    angle1 = (angle2[src] + rng.normal(0, angle_noise, n1)).astype(f32)
    wild = rng.random(n1) < angle_outlier_frac
    angle1[wild] = rng.uniform(0, 360, int(wild.sum())).astype(f32)
    angle1 = np.mod(angle1, f32(360.0)).astype(f32)
    return dict(xy_un1=xy1, octave1=octave1, angle1=angle1, desc1=desc1, prev_matched=xy1.copy())


def make_frustum_points(rng, K4, bounds, Tcw, n_points, sf, extreme_frac=0.02):
    """Local map points for Frame::isInFrustum (src/Frame.cc:269-325, Tracking::SearchLocalPoints): positions spread well beyond
    the image (so the bounds test rejects some), a few behind the camera, mean viewing directions from head-on to grazing (the
    0.5 cosine limit cuts through them), scale-invariance ranges placed so that MapPoint::PredictScale lands on every level and
    the distance test rejects some, plus a few extreme values (ranges of zero width, huge / tiny distances, points already seen)."""
    fx, fy, cx, cy = [float(v) for v in K4]
    minx, maxx, miny, maxy = [float(v) for v in bounds]
    Tc = np.asarray(Tcw, np.float64).reshape(4, 4)
    Rwc, twc = Tc[:3, :3].T, -Tc[:3, :3].T @ Tc[:3, 3]
    z = rng.uniform(0.5, 15.0, n_points)
    px = rng.uniform(minx - 0.15 * (maxx - minx), maxx + 0.15 * (maxx - minx), n_points)
    py = rng.uniform(miny - 0.15 * (maxy - miny), maxy + 0.15 * (maxy - miny), n_points)
    Xc = np.stack([(px - cx) / fx * z, (py - cy) / fy * z, z], 1)
    behind = rng.random(n_points) < 0.04
    Xc[behind] *= -1.0
    world = (Xc @ Rwc.T + twc).astype(f32)
    PO = world.astype(np.float64) - twc
    dist = np.linalg.norm(PO, axis=1)
    # mean viewing direction: the ray rotated by a random angle in [0, 80] degrees about a random axis
    ray = PO / np.maximum(dist, 1e-9)[:, None]
    axis = rng.normal(0, 1, (n_points, 3))
    axis -= (axis * ray).sum(1, keepdims=True) * ray
    axis /= np.maximum(np.linalg.norm(axis, axis=1, keepdims=True), 1e-9)
    ang = np.deg2rad(rng.uniform(0, 80, n_points))
    normal = (ray * np.cos(ang)[:, None] + axis * np.sin(ang)[:, None]).astype(f32)
    # mfMaxDistance = dist * scale^(level + jitter): PredictScale = ceil(log(max / dist) / log(scale)) spreads over all levels (+- 2 beyond)
    nl = len(sf)
    lvl = rng.integers(-2, nl + 2, n_points)
    jitter = rng.uniform(-0.999, 0.0, n_points)
    s1 = float(sf[1])
    max_dist = (dist * np.power(s1, lvl + jitter)).astype(f32)
    min_dist = (max_dist / f32(sf[-1]) * rng.uniform(0.6, 1.1, n_points)).astype(f32)
    on_boundary = rng.random(n_points) < 0.05                       # ratios that sit exactly on a power of the scale factor
    max_dist[on_boundary] = (dist[on_boundary].astype(f32) * np.power(f32(s1), rng.integers(0, nl, int(on_boundary.sum())).astype(f32))).astype(f32)
    ext = rng.random(n_points) < extreme_frac
    k = int(ext.sum())
    max_dist[ext] = rng.choice(np.array([0.0, 1e-30, 1e30, 3e38], f32), k)
    min_dist[ext] = rng.choice(np.array([0.0, 1e-30, 1e-3], f32), k)
    consider = (rng.random(n_points) >= 0.1).astype(np.uint8)       # 0: already seen in this frame / bad (src/Tracking.cc:1169-1172)
    return dict(consider=consider, world=world, normal=normal, min_dist=min_dist, max_dist=max_dist)
