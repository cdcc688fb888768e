#!/usr/bin/env python3
"""e2e throughput with T host threads, each with its own handle (debug aid)."""
import sys, os, time, threading, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orbslam2_with_quadrics_b200 import ORBextractor, _capi
from orbslam2_with_quadrics_b200.frames import cluttered_scene
w, h = 1920, 1080
base = [cluttered_scene(w, h, 1234 + i) for i in range(4)]
L = _capi.lib()
for T, n in [(1, 32), (2, 16), (2, 32), (3, 16), (4, 8), (4, 16)]:
    host = torch.zeros((T * n, h, w), dtype=torch.uint8).pin_memory()
    for i in range(T * n): host[i] = torch.from_numpy(base[i % 4])
    exs = [ORBextractor(2000, 1.2, 8, 20, 7, max_batch=n, download_pyramid=False) for _ in range(T)]
    args = []
    for t in range(T):
        ptrs = (C.c_void_p * n)(*[host[t * n + i].data_ptr() for i in range(n)]); strides = (C.c_size_t * n)(*[w] * n)
        args.append((ptrs, strides, (_capi.OrbxResult * n)()))
    K = 20
    def run(t, k):
        for _ in range(k): L.orbx_extract_batch(exs[t]._h, n, args[t][0], w, h, args[t][1], args[t][2])
    for t in range(T): run(t, 3)
    th = [threading.Thread(target=run, args=(t, K)) for t in range(T)]
    t0 = time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]; dt = time.perf_counter() - t0
    print("threads=%d batch=%d: %.0f frames/s (%.3f ms per %d frames)" % (T, n, T * n * K / dt, dt / K * 1e3, T * n))
    for e in exs: e.close()
