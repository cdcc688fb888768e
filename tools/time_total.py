#!/usr/bin/env python3
"""Whole-step device time (overlapped schedule, no stage events)."""
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
st = torch.cuda.ExternalStream(ex.stream)
for _ in range(5): ex.extract_device(dev.data_ptr(), B, w, h, pitch, h * pitch)
ex.synchronize()
K = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(K): ex.extract_device(dev.data_ptr(), B, w, h, pitch, h * pitch)
e1.record(st); ex.synchronize(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(name, "B=%d chunks=%s: %.4f ms/step  %.0f frames/s" % (B, os.environ.get("ORBX_DEVICE_CHUNKS", "2"), ms, B / ms * 1e3))
