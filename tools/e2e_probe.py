#!/usr/bin/env python3
"""Where does end-to-end time go? (debug aid)"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orbslam2_with_quadrics_b200 import ORBextractor, _capi
from orbslam2_with_quadrics_b200.frames import cluttered_scene
w, h, n = 1920, 1080, 32
host = torch.zeros((n, h, w), dtype=torch.uint8).pin_memory()
base = [cluttered_scene(w, h, 1234 + i) for i in range(4)]
for i in range(n): host[i] = torch.from_numpy(base[i % 4])
dev = torch.empty_like(host, device="cuda")
torch.cuda.synchronize()
for _ in range(3): dev.copy_(host, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): dev.copy_(host, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
print("H2D pinned %.1f MB in %.3f ms = %.1f GB/s" % (host.numel() / 1e6, dt * 1e3, host.numel() / dt / 1e9))
back = torch.empty_like(host).pin_memory()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): back.copy_(dev, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
print("D2H pinned %.3f ms = %.1f GB/s" % (dt * 1e3, host.numel() / dt / 1e9))
ex = ORBextractor(2000, 1.2, 8, 20, 7, max_batch=n, download_pyramid=False)
views = [host[i].numpy() for i in range(n)]
for _ in range(3): ex.extract_batch(views)
t = time.perf_counter()
for _ in range(10): ex.extract_batch(views)
print("python extract_batch %.3f ms/step" % ((time.perf_counter() - t) / 10 * 1e3))
# raw C call
L = _capi.lib()
ptrs = (C.c_void_p * n)(*[v.ctypes.data for v in views]); strides = (C.c_size_t * n)(*[w] * n)
res = (_capi.OrbxResult * n)()
t = time.perf_counter()
for _ in range(10): L.orbx_extract_batch(ex._h, n, ptrs, w, h, strides, res)
print("C orbx_extract_batch %.3f ms/step" % ((time.perf_counter() - t) / 10 * 1e3))
for b in (1, 4, 8):
    t = time.perf_counter()
    for _ in range(20): L.orbx_extract_batch(ex._h, b, ptrs, w, h, strides, res)
    print("C orbx_extract_batch n=%d %.3f ms/call" % (b, (time.perf_counter() - t) / 20 * 1e3))
