"""Seeded fuzz of the whole path against the oracle: random image sizes (odd widths, strides), feature counts, scale
factors, level counts and FAST thresholds, three frame generators.  Every case must be bit-exact (keypoints, order,
angles, descriptors, every pyramid level) or be rejected with the documented geometry error on BOTH sides."""
import numpy as np
import pytest

from oracle import orb_oracle
from orbslam2_with_quadrics_b200 import ORBextractor, OrbxError, _capi
from orbslam2_with_quadrics_b200 import frames as fr

pytestmark = pytest.mark.gpu


def random_case(i):
    rng = np.random.default_rng(9000 + i)
    w = int(rng.integers(160, 1300))
    h = int(rng.integers(max(120, w // 2 + 8), max(min(900, int(1.6 * w)), max(120, w // 2 + 8) + 1)))   # width / height > 0.5 keeps nIni >= 1 (App. B-7)
    scale = float(rng.choice([1.1, 1.15, 1.2, 1.25, 1.3, 1.41, 1.5, 1.7, 2.0]))
    nl = int(rng.integers(2, 9))
    while min(w, h) / scale ** (nl - 1) < 70 and nl > 1:   # last level must stay >= 62 px (App. B-7b)
        nl -= 1
    nf = int(rng.integers(100, 2500))
    ini = int(rng.integers(8, 40))
    mn = int(rng.integers(2, ini + 1))
    kind = int(rng.integers(0, 3))
    return w, h, nf, scale, nl, ini, mn, kind


@pytest.mark.parametrize("i", range(24))
def test_random_geometry_and_parameters(i):
    w, h, nf, scale, nl, ini, mn, kind = random_case(i)
    img = (fr.cluttered_scene(w, h, 100 + i), fr.checker_frame(w, h, 200 + i), fr.noise_frame(w, h, 300 + i))[kind]
    pad = np.zeros((h, w + (i % 5) * 3), np.uint8)          # row strides that are not the width
    pad[:, :w] = img
    view = pad[:, :w]
    ro = orb_oracle.ORBextractor(nf, scale, nl, ini, mn)(np.ascontiguousarray(view))
    # uniform noise makes ~9 % of the pixels FAST corners: give the candidate buffers room (capacity = pixels / divisor)
    gx = ORBextractor(nf, scale, nl, ini, mn, candidate_divisor=2 if kind == 2 else 0)
    kps, desc = gx(view)
    assert len(kps) == ro.n, (w, h, nf, scale, nl, ini, mn, kind)
    for f in kps.dtype.names:
        a, b = kps[f], ro.keypoints[f]
        assert np.array_equal(a.view(np.uint32) if a.dtype.kind == "f" else a, b.view(np.uint32) if b.dtype.kind == "f" else b), f
    assert np.array_equal(desc, ro.descriptors)
    for l in range(nl):
        assert np.array_equal(gx.stage_dump(0, l, _capi.STAGE_PYRAMID), ro.pyramid[l]), l
    gx.close()
