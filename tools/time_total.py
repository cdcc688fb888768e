#!/usr/bin/env python3
"""Whole-step device time (overlapped schedule, no stage events).  ORBX_TT_HANDLES=H: steps alternate over H handles
(double buffering: step k + 1 starts while step k's tail is still running)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orbslam2_with_quadrics_b200 import ORBextractor
from orbslam2_with_quadrics_b200.frames import CONFIGS, cluttered_scene
name = sys.argv[1] if len(sys.argv) > 1 else "rgbd_1080p"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
H = int(os.environ.get("ORBX_TT_HANDLES", "1"))
w, h, nf, sf, nl, it, mt, _ = CONFIGS[name]
base = [cluttered_scene(w, h, 1234 + i) for i in range(8)]
pitch = (w + 15) // 16 * 16
host = np.zeros((B, h, pitch), np.uint8)
for i in range(B): host[i, :, :w] = base[i % 8]
dev = torch.from_numpy(host).cuda()
exs = [ORBextractor(nf, sf, nl, it, mt, max_batch=B, download_pyramid=False) for _ in range(H)]
sts = [torch.cuda.ExternalStream(ex.stream) for ex in exs]
for k in range(5 * H): exs[k % H].extract_device(dev.data_ptr(), B, w, h, pitch, h * pitch)
for ex in exs: ex.synchronize()
torch.cuda.synchronize()
K = 30 * H
e0 = torch.cuda.Event(enable_timing=True)
ends = [torch.cuda.Event(enable_timing=True) for _ in range(H)]
e0.record(sts[0])
for s in sts[1:]: s.wait_event(e0)
for k in range(K): exs[k % H].extract_device(dev.data_ptr(), B, w, h, pitch, h * pitch)
for e, s in zip(ends, sts): e.record(s)
for ex in exs: ex.synchronize()
torch.cuda.synchronize()
ms = max(e0.elapsed_time(e) for e in ends) / K
print(name, "B=%d chunks=%s handles=%d: %.4f ms/step  %.0f frames/s" % (B, os.environ.get("ORBX_DEVICE_CHUNKS", "2"), H, ms, B / ms * 1e3))
