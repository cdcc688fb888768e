#!/usr/bin/env python3
"""e2e probe for the stereo configs (debug aid): effect of the pyramid D2H and of threads."""
import sys, os, time, threading, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orbslam2_with_quadrics_b200 import ORBextractor, _capi
from orbslam2_with_quadrics_b200.frames import cluttered_scene, CONFIGS
name = sys.argv[1] if len(sys.argv) > 1 else "stereo_euroc"
w, h, nf, sf, nl, it, mt, nimg = CONFIGS[name]
base = [cluttered_scene(w, h, 1234 + i) for i in range(4)]
L = _capi.lib()
pitch = (w + 15) // 16 * 16
for T, n, pyr in [(1, 32, 0), (1, 32, 1), (2, 16, 1), (2, 16, 0), (4, 8, 1), (1, 2, 1), (2, 2, 1), (4, 2, 1), (8, 2, 1)]:
    host = torch.zeros((T * n, h, pitch), dtype=torch.uint8).pin_memory()
    for i in range(T * n): host[i, :, :w] = torch.from_numpy(base[i % 4])
    exs = [ORBextractor(nf, sf, nl, it, mt, max_batch=n, download_pyramid=bool(pyr)) for _ in range(T)]
    args = []
    for t in range(T):
        ptrs = (C.c_void_p * n)(*[host[t * n + i].data_ptr() for i in range(n)]); strides = (C.c_size_t * n)(*[pitch] * n)
        args.append((ptrs, strides, (_capi.OrbxResult * n)()))
    K = 30
    def run(t, k):
        for _ in range(k): L.orbx_extract_batch(exs[t]._h, n, args[t][0], w, h, args[t][1], args[t][2])
    for t in range(T): run(t, 3)
    th = [threading.Thread(target=run, args=(t, K)) for t in range(T)]
    t0 = time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]; dt = time.perf_counter() - t0
    print("%s threads=%d imgs/call=%d pyramid_d2h=%d: %.0f images/s (%.3f ms per call)" % (name, T, n, pyr, T * n * K / dt, dt / K * 1e3))
    for e in exs: e.close()
