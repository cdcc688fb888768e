#!/usr/bin/env python3
"""Benchmark of the ORB feature front end (BASELINE.json metric: ORB extraction frames/s + latency).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl orbx|reference] [--config NAME] [--batch B]

A step = one pass of ORBextractor::operator() over one batch of synthetic frames per GPU.
`value`   : frames/s with the batch already resident in HBM (CUDA events on the library's stream).
`e2e`     : frames/s through the C ABI with pinned HOST buffers, H2D + D2H inside the timed region.
`roofline`: algorithmic bytes / device time against the measured HBM peak.
`--impl reference`: the reference's own unmodified ORBextractor TU (oracle/_ref) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from orbslam2_with_quadrics_b200 import frames as fr          # noqa: E402
from orbslam2_with_quadrics_b200 import geometry as geo       # noqa: E402

METRIC = "orb_extraction_frames_per_sec"
UNIT = "frames/s"
DISTINCT_FRAMES = 8


_REAL_STDOUT = None


def claim_stdout():
    """Rank 0 prints exactly ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner
    on the first collective whatever NCCL_DEBUG says), so file descriptor 1 is pointed at stderr for the rest of the
    process and the JSON line goes to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def workload_desc(name):
    w, h, nf, sf, nl, it, mt, nimg = fr.CONFIGS[name]
    return "%s: %dx%d 8-bit gray%s, nFeatures=%d, %d levels, scale %.1f, FAST %d/%d" % (
        name, w, h, " x2 (stereo pair = 1 frame)" if nimg == 2 else "", nf, nl, sf, it, mt)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_frames(name, stream, count):
    w, h, *_rest, nimg = fr.CONFIGS[name]
    base = [fr.cluttered_scene(w, h, fr.stream_seed(stream, i)) for i in range(min(count, DISTINCT_FRAMES))]
    return [base[i % len(base)] for i in range(count)]


# ------------------------------------------------------------------------------------------- reference arm
def run_reference_cpu(name, frames_per_step, steps, warmup, cores):
    """Times the reference's own code (unmodified TU on the OpenCV shim), one thread per frame."""
    from oracle import ref_lib                 # the one place bench.py executes oracle/: the CPU arm
    if not ref_lib.available():
        ref_lib.build()
    w, h, nf, sf, nl, it, mt, nimg = fr.CONFIGS[name]
    imgs = make_frames(name, 0, DISTINCT_FRAMES)
    nthreads = max(1, min(cores, frames_per_step))
    exs = [ref_lib.RefORBextractor(nf, sf, nl, it, mt) for _ in range(nthreads)]
    lat = []

    def work(tid, nsteps, record):
        for s in range(nsteps):
            for f in range(tid, frames_per_step, nthreads):
                for k in range(nimg):
                    t0 = time.perf_counter()
                    exs[tid](imgs[(f + k) % len(imgs)])
                    if record:
                        lat.append(time.perf_counter() - t0)

    def run(nsteps, record):
        th = [threading.Thread(target=work, args=(t, nsteps, record)) for t in range(nthreads)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    run(warmup, False)
    el = run(steps, True)
    fps = frames_per_step * steps / el
    # single-thread per-image latency (the other host threads idle), SURVEY.md §8(d)
    solo = []
    for i in range(4):
        t0 = time.perf_counter()
        exs[0](imgs[i % len(imgs)])
        solo.append((time.perf_counter() - t0) * 1e3)
    lat_ms = np.asarray(lat) * 1e3
    return {"value": fps, "unit": UNIT, "cores": nthreads, "kind": "reference",
            "sample": "%d frames/step x %d steps of %s, one thread per frame on %d host threads; "
                      "unmodified reference ORBextractor.cc on the scalar OpenCV shim (oracle/_ref)" % (
                          frames_per_step, steps, name, nthreads),
            "image_latency_ms_p50": float(np.percentile(lat_ms, 50)), "image_latency_ms_p99": float(np.percentile(lat_ms, 99)),
            "single_thread_image_latency_ms_p50": float(np.percentile(solo, 50)),
            "elapsed_s": el}


def opencv_primitive_times(name, reps=3):
    """Cross-check for the CPU baseline: single-thread time of the real OpenCV (cv2, SIMD) primitives
    the reference calls -- pyramid, FAST on whole levels (a lower bound on the per-cell calls) and the blur -- on
    one frame of the workload.  The reference's own code (cell loop, quadtree, orientation, descriptors) comes on top."""
    try:
        import cv2
    except Exception:
        return None
    cv2.setNumThreads(1)
    w, h, nf, sf, nl, it, mt, nimg = fr.CONFIGS[name]
    img = fr.cluttered_scene(w, h, fr.stream_seed(0, 0))
    sizes = geo.level_sizes(w, h, sf, nl)
    det = cv2.FastFeatureDetector_create(threshold=it, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    best = {"pyramid": 1e9, "fast_whole_levels": 1e9, "blur": 1e9}
    for _ in range(reps):
        t0 = time.perf_counter()
        lv = [cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)]
        for lw, lh in sizes[1:]:
            r = cv2.resize(lv[-1][19:-19, 19:-19], (lw, lh), interpolation=cv2.INTER_LINEAR)
            lv.append(cv2.copyMakeBorder(r, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
        t1 = time.perf_counter()
        for p in lv:
            det.detect(np.ascontiguousarray(p[19:-19, 19:-19]), None)
        t2 = time.perf_counter()
        for p in lv:
            cv2.GaussianBlur(np.ascontiguousarray(p[19:-19, 19:-19]), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        t3 = time.perf_counter()
        best["pyramid"] = min(best["pyramid"], (t1 - t0) * 1e3)
        best["fast_whole_levels"] = min(best["fast_whole_levels"], (t2 - t1) * 1e3)
        best["blur"] = min(best["blur"], (t3 - t2) * 1e3)
    best["sum"] = best["pyramid"] + best["fast_whole_levels"] + best["blur"]
    best["note"] = "cv2 %s, 1 thread, ms per image, best of %d" % (cv2.__version__, reps)
    return best


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    fps_step = max(8, 2 * cores) if args.config in ("rgbd_1080p", "mono_4k") else max(16, 4 * cores)
    r = run_reference_cpu(args.config, fps_step, args.steps, args.warmup, cores)
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["elapsed_s"] / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_desc(args.config), "frames_per_step": fps_step, "host_threads": r["cores"]},
            "cpu_baseline": r,
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------- own arm
def source_hash():
    """sha256 over the sources the library is built from: profiles/dominant_kernel_traffic.json records the hash of the
    tree its ncu capture was taken from, and a capture of another tree is not reported (roofline.traffic = null)."""
    import hashlib
    hsh = hashlib.sha256()
    d = os.path.join(ROOT, "orbslam2_with_quadrics_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".h", ".inc")):
            hsh.update(f.encode()); hsh.update(open(os.path.join(d, f), "rb").read())
    return hsh.hexdigest()[:16]


def h2d_ceiling(host, dev, reps, chunks, sharding, torch):
    """Pinned host -> device copies of one step's input, no kernels, all ranks at once: what the box's host side can feed."""
    n = host.numel()
    hv, dv = host.view(-1), dev.view(-1)
    step = n // chunks
    for _ in range(2):
        dv.copy_(hv, non_blocking=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sharding.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        for c in range(chunks):
            dv[c * step:(c + 1) * step].copy_(hv[c * step:(c + 1) * step], non_blocking=True)
    e1.record()
    torch.cuda.synchronize(); sharding.barrier()
    return sharding.max_over_ranks(e0.elapsed_time(e1)) / reps          # ms per step, slowest rank


def measure_config(args, name, B, K, W, rank, world, local, full, sampler):
    """One BASELINE config on this rank's GPU: device-resident value, per-stage times, e2e through the C ABI.
    full = the headline config (adds sustained, latency, drop-in defaults); otherwise a short pass for `extra_configs`."""
    import ctypes as C
    import torch
    from orbslam2_with_quadrics_b200 import ORBextractor, sharding, _capi

    w, h, nf, sf, nl, it, mt, nimg = fr.CONFIGS[name]
    nimgs = B * nimg                     # images per step per GPU
    balg = geo.algorithmic_bytes(w, h, nf, sf, nl) * nimg           # per frame
    stage_bytes = geo.stage_algorithmic_bytes(w, h, nf, sf, nl)

    # ---- synthetic input: this rank serves camera stream `rank` (no cross-rank data dependency)
    if nimg == 1:
        imgs = make_frames(name, rank, nimgs)
    else:              # stereo: rectified synthetic pairs (right = left shifted by band-wise disparities), pairs kept on this GPU
        base = [fr.stereo_pair(w, h, fr.stream_seed(rank, i)) for i in range(min(B, DISTINCT_FRAMES))]
        imgs = [x for i in range(B) for x in base[i % len(base)]]
    pitch = (w + 15) // 16 * 16
    host = torch.zeros((nimgs, h, pitch), dtype=torch.uint8).pin_memory()
    hnp = host.numpy()
    for i, im in enumerate(imgs):
        hnp[i, :, :w] = im
    dev = host.cuda(non_blocking=False)

    # only ComputeStereoMatches reads mvImagePyramid (src/Frame.cc:563-580): with the matcher on the device
    # (orbx_stereo_match, the default) the pyramids stay in HBM; --stereo-match host ships them to the CPU consumer
    stereo_dev = nimg == 2 and args.stereo_match == "device"
    need_pyr = nimg == 2 and not stereo_dev
    mbf, mb = (386.1448, 0.5371657) if name == "stereo_kitti" else (47.90639384423901, 0.11007784)   # KITTI00-02.yaml / EuRoC.yaml
    # Steps alternate over `--handles` extractor handles (default 2: double buffering -- the results of step k stay in HBM
    # while step k + 1 runs, and the latency-bound tail of one step overlaps the head of the next; each handle then takes
    # its 64-frame batch as ONE launch sequence: 1.373 -> 1.316 ms per 64 x 1080p).  --handles 1: one handle, two
    # half-batches on two streams, every step joined before the next one starts.
    NH = max(1, args.handles)
    exs = [ORBextractor(nf, sf, nl, it, mt, device=local, max_batch=nimgs, download_pyramid=False,
                        device_chunks=1 if NH > 1 else 0) for _ in range(NH)]
    ex = exs[0]
    tdev = torch.device("cuda", local)
    streams = [torch.cuda.ExternalStream(e.stream, device=tdev) for e in exs]
    stream = streams[0]
    step_no = [0]

    def step_device():
        exs[step_no[0] % NH].extract_device(dev.data_ptr(), nimgs, w, h, pitch, h * pitch)
        step_no[0] += 1

    def timed_steps(n):
        """n steps back to back; device time from the first handle's stream at the start to the last stream to finish"""
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(NH)]
        step_no[0] = 0
        sharding.barrier(); torch.cuda.synchronize()
        e0.record(stream)
        for s in streams[1:]:
            s.wait_event(e0)
        for _ in range(n):
            step_device()
        for e, s in zip(ends, streams):
            e.record(s)
        for e in exs:
            e.synchronize()
        torch.cuda.synchronize(); sharding.barrier()
        return max(e0.elapsed_time(e) for e in ends)

    for _ in range(W * NH):
        step_device()
    for e in exs:
        e.synchronize()
    res = ex.fetch_results(nimgs)
    kp_counts = [len(k) for k, _ in res]
    cand_per_level, retries_per_level = ex.fast_stats(0)     # FAST corners handed to the quadtree / cells re-run at minThFAST, frame 0
    if min(kp_counts) == 0:
        raise SystemExit("bench: a frame produced no keypoints; refusing to time a degenerate run")

    l0 = sum(e.launch_count for e in exs)
    e0 = torch.cuda.Event(enable_timing=True)
    dev_ms = timed_steps(K)
    launches = sum(e.launch_count for e in exs) - l0
    dev_ms_max = sharding.max_over_ranks(dev_ms)
    value = world * B * K / (dev_ms_max / 1e3)

    # ---- sustained: >= 2 s of back-to-back steps, its own clock sample beside it (the K-step window above is a burst)
    sustained = None
    if full and args.sustained_s > 0:
        Ks = max(K, int(np.ceil(args.sustained_s * 1e3 / (dev_ms / K))))
        s2 = ClockSampler(sampler.index) if rank == 0 else None
        if s2:
            s2.start()
            time.sleep(0.3)
        sus_ms = sharding.max_over_ranks(timed_steps(Ks))
        sustained = {"value": world * B * Ks / (sus_ms / 1e3), "unit": UNIT, "steps": Ks, "seconds": sus_ms / 1e3,
                     "ms_per_step": sus_ms / Ks, "clocks": s2.stop() if s2 else None}

    # second pass of K steps with per-stage CUDA events (each kernel then runs alone on the stream, i.e. without the
    # FAST / pyramid overlap of the timed pass): the per-kernel durations behind the roofline table
    ex.stage_timing(True)
    for _ in range(K):
        ex.extract_device(dev.data_ptr(), nimgs, w, h, pitch, h * pitch)
    ex.synchronize()
    stages = ex.stage_times()
    ex.stage_timing(False)
    for e in exs[1:]:
        e.close()

    # ---- what the host side of the box can feed: pinned -> device copies of the step's input, no kernels, all ranks at once
    chunks = 4 if B >= 16 else (2 if B >= 4 else 1)
    ceil_ms = h2d_ceiling(host, dev, 10, chunks, sharding, torch)
    in_bytes = nimgs * h * pitch
    ceil_gbs = world * in_bytes / (ceil_ms / 1e3) / 1e9

    # ---- e2e through the C ABI (orbx_extract_batch, the call the C++ adapter makes): pinned HOST frames in,
    #      keypoints + descriptors (+ pyramid for stereo) back in host memory when the call returns
    capi = _capi.lib()
    views = [hnp[i, :, :w] for i in range(nimgs)]
    T = max(1, min(args.e2e_threads if args.e2e_threads > 0 else auto_threads(world), B))   # host threads, one handle each (src/Frame.cc:78-81)
    per = [list(range(t * B // T * nimg, (t + 1) * B // T * nimg)) for t in range(T)]        # image indices per thread
    exs2 = [ORBextractor(nf, sf, nl, it, mt, device=local, max_batch=len(per[t]), download_pyramid=need_pyr) for t in range(T)]
    call = []
    for t in range(T):
        n_t = len(per[t])
        ptrs_t = (C.c_void_p * n_t)(*[views[i].__array_interface__["data"][0] for i in per[t]])
        strides_t = (C.c_size_t * n_t)(*[pitch] * n_t)
        call.append((exs2[t]._h, n_t, ptrs_t, strides_t, (_capi.OrbxResult * n_t)()))
    ptrs, strides = call[0][2], call[0][3]
    st_args = []
    for t in range(T):
        np_t = call[t][1] // 2
        st_args.append(((C.c_int * max(np_t, 1))(*range(0, 2 * np_t, 2)), (C.c_int * max(np_t, 1))(*range(1, 2 * np_t, 2)),
                        (_capi.OrbxStereoResult * max(np_t, 1))(), np_t))

    def worker(t, steps):
        if args.numa_bind:
            sharding.bind_to_gpu_numa(local)
        hd, n_t, p_t, s_t, r_t = call[t]
        lf, rf, sres, np_t = st_args[t]
        for _ in range(steps):
            _capi.check(capi.orbx_extract_batch(hd, n_t, p_t, w, h, s_t, r_t), hd)
            if stereo_dev:       # Frame::ComputeStereoMatches on the device-resident results; mvuRight / mvDepth come back
                _capi.check(capi.orbx_stereo_match(hd, hd, np_t, lf, rf, mbf, mb, sres), hd)

    def run_threads(steps):
        if T == 1:
            worker(0, steps)
            return
        th = [threading.Thread(target=worker, args=(t, steps)) for t in range(T)]
        for x in th:
            x.start()
        for x in th:
            x.join()

    run_threads(max(2, W // 2))
    Ke = max(3, min(K, 20))
    sharding.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_threads(Ke)
    for e in exs2:
        e.synchronize()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    sharding.barrier()
    e2e_ms_max = sharding.max_over_ranks(wall * 1e3)   # blocking host calls: wall clock between synchronised points, slowest rank
    e2e_value = world * B * Ke / (e2e_ms_max / 1e3)
    out = [kd for t in range(T) for kd in exs2[t]._copy_results(call[t][4], call[t][1])]
    n_out = int(np.mean([len(k) for k, _ in out]))
    slab = sum(((32 + lw + 19 + 63) // 64 * 64) * (lh + 38) for lw, lh in geo.level_sizes(w, h, sf, nl))
    kept_cap = sum(q + 4 * int(np.floor(float(np.float32(lw - 32) / np.float32(lh - 32)) + 0.5)) + 8
                   for q, (lw, lh) in zip(geo.level_quotas(nf, sf, nl), geo.level_sizes(w, h, sf, nl)))
    d2h = nimgs * kept_cap * 60 + 4 * (nimgs * (4 * nl + 1) + 16) + (nimgs * slab if need_pyr else 0)   # what the library copies
    if stereo_dev:
        d2h += nimgs * kept_cap * 8 + 4 * (nimgs * (4 * nl + 1) + 16)     # mvuRight + mvDepth of the batch, counters again
    e2e_in_gbs = world * nimgs * w * h * Ke / (e2e_ms_max / 1e3) / 1e9
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nimgs * w * h, "d2h_bytes_per_step": d2h,
           "steps": Ke, "pyramid_d2h": need_pyr, "host_threads": T,
           "stereo_match": ("device (orbx_stereo_match inside the timed region)" if stereo_dev else "host consumer (pyramid D2H)") if nimg == 2 else None,
           "call": "orbx_extract_batch (C ABI), %d frames per call per thread, pinned host frames" % (B // T),
           "h2d_ceiling_gbs": ceil_gbs, "h2d_achieved_gbs": e2e_in_gbs, "frac_of_h2d_ceiling": e2e_in_gbs / ceil_gbs,
           "h2d_ceiling_note": "pinned->device copies of the same %d bytes per rank per step (%d copies), no kernels, all %d ranks at once; "
                               "aggregate over ranks, slowest rank's time" % (in_bytes, chunks, world)}
    for e in exs2:
        e.close()

    # ---- per-frame latency through the C ABI, batch = 1 frame (rank 0 reports)
    lat = dropin = None
    if full and rank == 0 and args.latency_frames > 0:
        ex1 = ORBextractor(nf, sf, nl, it, mt, device=local, max_batch=nimg, download_pyramid=need_pyr)
        r1 = (_capi.OrbxResult * nimg)()
        psz = C.sizeof(C.c_void_p)
        for i in range(5):
            _capi.check(capi.orbx_extract_batch(ex1._h, nimg, ptrs, w, h, strides, r1), ex1._h)
        ts = []
        lf1, rf1, sres1 = (C.c_int * 1)(0), (C.c_int * 1)(1), (_capi.OrbxStereoResult * 1)()
        for i in range(args.latency_frames):
            k0 = (i * nimg) % call[0][1]
            pp = C.cast(C.byref(ptrs, k0 * psz), C.POINTER(C.c_void_p))
            t0 = time.perf_counter()
            rc = capi.orbx_extract_batch(ex1._h, nimg, pp, w, h, strides, r1)
            if stereo_dev and rc == 0:
                rc = capi.orbx_stereo_match(ex1._h, ex1._h, 1, lf1, rf1, mbf, mb, sres1)
            ts.append(time.perf_counter() - t0)
            _capi.check(rc, ex1._h)
        ts = np.asarray(ts) * 1e3
        lat = {"p50": float(np.percentile(ts, 50)), "p99": float(np.percentile(ts, 99)), "frames": len(ts),
               "batch": 1, "path": "C ABI, pinned host in, keypoints+descriptors out" + (", pyramid D2H" if need_pyr else "") +
                       (", stereo match on device, mvuRight+mvDepth out" if stereo_dev else "")}
        ex1.close()
        # ---- the drop-in's defaults (cpp/ORBextractor.cc: max_batch = 1, download_pyramid = 1, the caller's pageable
        #      cv::Mat): what Frame::ExtractORB (src/Frame.cc:247-253) gets without touching a single setting
        exd = ORBextractor(nf, sf, nl, it, mt, device=local, max_batch=1, download_pyramid=True)
        pageable = [np.array(imgs[i], copy=True, order="C") for i in range(min(len(imgs), DISTINCT_FRAMES))]
        rd = (_capi.OrbxResult * 1)()
        for i in range(5):
            _capi.check(capi.orbx_extract(exd._h, pageable[i % len(pageable)].ctypes.data, w, h, w, rd), exd._h)
        nd = max(50, min(args.latency_frames, 300))
        td = []
        tA = time.perf_counter()
        for i in range(nd):
            im = pageable[i % len(pageable)]
            t0 = time.perf_counter()
            rc = capi.orbx_extract(exd._h, im.ctypes.data, w, h, w, rd)
            td.append(time.perf_counter() - t0)
            _capi.check(rc, exd._h)
        tB = time.perf_counter()
        td = np.asarray(td) * 1e3
        dropin = {"p50_ms": float(np.percentile(td, 50)), "p99_ms": float(np.percentile(td, 99)), "images_per_s": nd / (tB - tA),
                  "images": nd, "path": "orbx_extract, max_batch 1, download_pyramid 1 (%.1f MB D2H per image), pageable input, "
                                        "one host thread: the C++ adapter's defaults" % (slab / 1e6)}
        exd.close()

    ex.close()
    del dev, host
    torch.cuda.empty_cache()
    # ---- roofline of the whole step and of each stage kernel
    stage_rows = []
    for nm, ms, ln in stages:
        sb = stage_bytes.get(nm, 0) * nimgs * K
        stage_rows.append({"kernel": nm, "ms_per_step": ms / K, "launches_per_step": ln / K,
                           "share": ms / max(sum(m for _, m, _ in stages), 1e-9),
                           "alg_gbs": (sb / (ms / 1e3) / 1e9) if ms > 0 else None})
    return {"name": name, "B": B, "nimgs": nimgs, "balg": balg, "value": value, "dev_ms": dev_ms, "dev_ms_max": dev_ms_max,
            "launches": int(launches), "stages": stage_rows, "e2e": e2e, "latency": lat, "dropin": dropin, "sustained": sustained,
            "n_out": n_out, "cand": cand_per_level, "retries": retries_per_level, "in_mb": nimgs * h * pitch / 1e6, "pyr_mb": nimgs * slab / 1e6}


def auto_threads(world):
    """Host threads per rank in the e2e leg: 4 measured best on one GPU (640x480: 2 threads 76.8k, 3-4 threads 99.9k
    frames/s); with many ranks on one host the threads of all ranks share the cores (plus one nvidia-smi sampler)."""
    return max(1, min(4, host_cores() // max(1, 2 * world)))


def roofline_of(m, K, peak, peak_src, name):
    achieved = m["balg"] * m["B"] * K / (m["dev_ms"] / 1e3) / 1e9
    dom = max(m["stages"], key=lambda r: r["ms_per_step"])
    traffic, traffic_note = None, "no ncu capture on file for this config"
    tf = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tf):
        try:
            doc = json.load(open(tf))
            t = doc.get(name)
            if t and t.get("source_hash") == source_hash():
                traffic = t["bytes_per_frame"] * m["B"]      # per launch sequence, like `achieved`
                traffic_note = t.get("source")
            elif t:
                traffic_note = "stale: the capture on file is of source tree %s, this tree is %s" % (t.get("source_hash"), source_hash())
        except Exception:
            pass
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
            "scope": "whole step (all stage kernels of the batch); algorithmic bytes %d per frame x %d frames per launch sequence" % (m["balg"], m["B"]),
            "dominant_kernel": dom["kernel"], "stages": m["stages"]}


def own_arm(args):
    import torch
    from orbslam2_with_quadrics_b200 import sharding

    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    rank, world, local = sharding.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    numa = (-1, 0)
    if args.numa_bind:                  # this rank's threads and pinned buffers on the CPUs next to its GPU
        numa = sharding.bind_to_gpu_numa(local)
    name = args.config
    K, W = args.steps, args.warmup
    sampler = ClockSampler(local if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local])
    if rank == 0:
        sampler.start()           # nvidia-smi needs a moment to start: it covers warm-up, the timed region and e2e
    m = measure_config(args, name, args.batch, K, W, rank, world, local, True, sampler)
    clocks = sampler.stop() if rank == 0 else None

    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        peak, peak_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    # ---- multi-GPU runs also cover the other multi-GPU configs of BASELINE.json with a short pass each:
    #      mono_4k in 64-frame batches split over the GPUs (config 5) and stereo_kitti batched per GPU (config 3)
    extras = []
    if world > 1 and not args.no_extra_configs:
        for xname, xB in (("mono_4k", max(1, 64 // world)), ("stereo_kitti", 32)):
            if xname == name:
                continue
            xm = measure_config(args, xname, xB, 5, 3, rank, world, local, False, sampler)
            if rank == 0:
                xr = roofline_of(xm, 5, peak, peak_src, xname)
                extras.append({"config": {"workload": workload_desc(xname), "batch_per_gpu": xB, "frames_per_step": world * xB},
                               "value": xm["value"], "unit": UNIT, "steps": 5, "warmup": 3, "ms_per_step": xm["dev_ms_max"] / 5,
                               "roofline": {k: xr[k] for k in ("achieved", "peak", "frac", "unit")},
                               "e2e": {k: xm["e2e"][k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "host_threads",
                                                                 "h2d_ceiling_gbs", "frac_of_h2d_ceiling", "stereo_match")}})
    if world > 1:
        sharding.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return 0
    roofline = roofline_of(m, K, peak, peak_src, name)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        per = max(8, cores)
        cpu = run_reference_cpu(name, per, 2, 1, cores)
        cpu["opencv_primitives_ms"] = opencv_primitive_times(name)
    B = args.batch
    line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": m["dev_ms_max"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_desc(name), "batch_per_gpu": B, "frames_per_step": world * B,
                       "distinct_frames": DISTINCT_FRAMES, "parallelism": "one camera stream per GPU, no collective",
                       "device_steps": ("steps alternate over %d handles (double buffering: a step's tail overlaps the next step's head); every step is one "
                                        "orbx_extract_device call on the whole batch" % args.handles) if args.handles > 1 else
                                       "one handle, two half-batches on two streams, every step joined before the next starts",
                       "l2_policy": "per-step working set (inputs %.0f MB + pyramids %.0f MB per GPU) exceeds the 126 MB L2" % (
                           m["in_mb"], m["pyr_mb"]),
                       "keypoints_per_image": m["n_out"], "fast_candidates_per_level_frame0": m["cand"],
                       "fast_retried_cells_per_level_frame0": m["retries"],
                       "host_placement": {"numa_bind": bool(args.numa_bind), "gpu_numa_node_rank0": numa[0], "cpus_bound_rank0": numa[1],
                                          "host_cores_visible": host_cores()}},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": m["e2e"], "e2e_dropin": m["dropin"], "sustained": m["sustained"],
            "latency_ms": m["latency"], "gpu_launches": m["launches"], "clocks": clocks}
    if extras:
        line["extra_configs"] = extras
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="orbx", choices=["orbx", "reference"])
    ap.add_argument("--config", default="rgbd_1080p", choices=list(fr.CONFIGS))
    ap.add_argument("--batch", type=int, default=64, help="frames per step per GPU (1080p: 32 -> 37.2k, 64 -> 39.4k, 128 -> 40.2k, 256 -> 40.7k frames/s)")
    ap.add_argument("--handles", type=int, default=2, help="extractor handles the device-resident steps alternate over (2 = double buffering; 1 = one handle, every step joined before the next)")
    ap.add_argument("--latency-frames", type=int, default=1000, help="single-frame calls timed for p50/p99 (SURVEY §8d: >= 1000)")
    ap.add_argument("--e2e-threads", type=int, default=0, help="host threads (one handle each) in the e2e measurement; 0 = auto: min(4, cores / (2 * ranks))")
    ap.add_argument("--no-numa-bind", dest="numa_bind", action="store_false", help="leave the rank's threads where the scheduler puts them")
    ap.add_argument("--sustained-s", type=float, default=2.0, help="seconds of back-to-back steps for the `sustained` record (0 = skip)")
    ap.add_argument("--no-extra-configs", action="store_true", help="multi-GPU runs: skip the short mono_4k / stereo_kitti passes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stereo-match", default="device", choices=["device", "host"],
                    help="stereo configs: run Frame::ComputeStereoMatches on the GPU (pyramids stay in HBM) or leave it to a "
                         "host consumer (pyramid D2H inside the timed region)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    claim_stdout()
    if args.impl == "reference":
        return reference_arm(args)
    return own_arm(args)


if __name__ == "__main__":
    sys.exit(main())
