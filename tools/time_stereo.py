#!/usr/bin/env python3
"""Device time of orbx_stereo_match_device (stereo_match_kernel + stereo_filter_kernel) for a batch of pairs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orbslam2_with_quadrics_b200 import ORBextractor
from orbslam2_with_quadrics_b200.frames import CONFIGS, stereo_pair
name = sys.argv[1] if len(sys.argv) > 1 else "stereo_euroc"
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 16
w, h, nf, sf, nl, it, mt, _ = CONFIGS[name]
pairs = [stereo_pair(w, h, 1234 + i) for i in range(4)]
imgs = [im for i in range(NP) for im in pairs[i % 4]]
ex = ORBextractor(nf, sf, nl, it, mt, max_batch=2 * NP, download_pyramid=False)
res = ex.extract_batch(imgs)
lf, rf = list(range(0, 2 * NP, 2)), list(range(1, 2 * NP, 2))
mbf, mb = (47.90639384423901, 0.11007784) if name != "stereo_kitti" else (386.1448, 0.53716572)
out = ex.stereo_match(ex, mbf, mb, lf, rf)
print("matched per pair:", [int((u >= 0).sum()) for u, _ in out[:4]], "of", [len(u) for u, _ in out[:4]])
st = torch.cuda.ExternalStream(ex.stream)
K = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(K): ex.stereo_match_device(ex, mbf, mb, lf, rf)
e1.record(st); ex.synchronize()
print(name, "pairs=%d: stereo match %.4f ms per batch" % (NP, e0.elapsed_time(e1) / K))
