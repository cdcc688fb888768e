#!/usr/bin/env python3
"""Per-stage device time of the launch sequence, no parity checks (for A/B builds via ORBX_LIB)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orbslam2_with_quadrics_b200 import ORBextractor
from orbslam2_with_quadrics_b200.frames import CONFIGS, cluttered_scene
name = sys.argv[1] if len(sys.argv) > 1 else "rgbd_1080p"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
w, h, nf, sf, nl, it, mt, _ = CONFIGS[name]
base = [cluttered_scene(w, h, 1234 + i) for i in range(8)]
pitch = (w + 15) // 16 * 16
host = np.zeros((B, h, pitch), np.uint8)
for i in range(B): host[i, :, :w] = base[i % 8]
dev = torch.from_numpy(host).cuda()
ex = ORBextractor(nf, sf, nl, it, mt, max_batch=B, download_pyramid=False)
for _ in range(3): ex.extract_device(dev.data_ptr(), B, w, h, pitch, h * pitch)
ex.synchronize()
ex.stage_timing(True)
K = 20
for _ in range(K): ex.extract_device(dev.data_ptr(), B, w, h, pitch, h * pitch)
ex.synchronize()
st = ex.stage_times()
print(os.environ.get("ORBX_LIB", "default"), name, "B=%d" % B, " ".join("%s=%.4f" % (n, ms / K) for n, ms, _ in st), "total=%.4f" % (sum(ms for _, ms, _ in st) / K))
