"""In-process multi-device use of the C ABI (orbx_config.device != 0): two handles on two GPUs of one process, driven
from two host threads the way a multi-camera rig would (one camera stream per GPU, no collective).
Needs a 2-GPU box (`gpurun --gpus 2`); skipped elsewhere."""
import threading

import numpy as np
import pytest

from orbslam2_with_quadrics_b200 import frames as fr

pytestmark = pytest.mark.gpu


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


def test_two_handles_on_two_devices_agree_with_the_oracle():
    if not _two_gpus():
        pytest.skip("needs two GPUs")
    from orbslam2_with_quadrics_b200 import ORBextractor
    from oracle import orb_oracle
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["mono_tum"]
    imgs = [fr.cluttered_scene(w, h, 500 + i) for i in range(6)]
    refs = [orb_oracle.ORBextractor(nf, sf, nl, it, mt)(im) for im in imgs]
    exs = [ORBextractor(nf, sf, nl, it, mt, device=d, max_batch=3) for d in (0, 1)]
    out = [None, None]

    def work(d):
        res = []
        for rep in range(3):                          # interleaved with the other device's calls
            res = exs[d].extract_batch(imgs[3 * d:3 * d + 3])
        out[d] = res

    th = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for d in (0, 1):
        for i, (kp, desc) in enumerate(out[d]):
            ro = refs[3 * d + i]
            assert len(kp) == ro.n
            for f in kp.dtype.names:
                assert np.array_equal(kp[f], ro.keypoints[f]), (d, i, f)
            assert np.array_equal(desc, ro.descriptors), (d, i)
    # the pyramid of a frame extracted on device 1 comes back from device 1's memory
    for l, plane in enumerate(exs[1].pyramid(0)):
        assert np.array_equal(plane, refs[3].pyramid[l]), l
    for e in exs:
        e.close()


def test_consumers_on_second_device():
    """The §8(f) rows on a non-zero device: stereo matcher + undistort/grid on the results a device-1 extraction left in HBM."""
    if not _two_gpus():
        pytest.skip("needs two GPUs")
    from orbslam2_with_quadrics_b200 import ORBextractor
    from oracle import orb_oracle, stereo_oracle
    w, h, nf, sf, nl, it, mt, _ = fr.CONFIGS["stereo_euroc"]
    left, right = fr.stereo_pair(w, h, 909)
    mbf = 47.90639384423901
    mb = float(np.float32(mbf) / np.float32(435.2046959714599))
    ex = ORBextractor(nf, sf, nl, it, mt, device=1, max_batch=2, download_pyramid=False)
    ex.extract_batch([left, right])
    (u, d), = ex.stereo_match(ex, mbf, mb, left_frames=[0], right_frames=[1])
    oex = orb_oracle.ORBextractor(nf, sf, nl, it, mt)
    rl, rr = oex(left), oex(right)
    uo, do, _ = stereo_oracle.compute_stereo_matches(rl.keypoints, rl.descriptors, rl.pyramid, rr.keypoints, rr.descriptors, rr.pyramid,
                                                     oex.GetScaleFactors(), oex.GetInverseScaleFactors(), mbf, mb)
    assert np.array_equal(np.asarray(u).view(np.uint32), uo.view(np.uint32))
    assert np.array_equal(np.asarray(d).view(np.uint32), do.view(np.uint32))
    ex.close()
