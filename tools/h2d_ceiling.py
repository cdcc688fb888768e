#!/usr/bin/env python3
"""Host->device ceiling of the box for the bench's e2e leg: N ranks (one per GPU, torchrun) copy
the bench's per-step input (default 64 x 1920x1080 bytes) from pinned host memory to their GPU, no kernels, all ranks at
the same time.  Prints one JSON line: per-rank and aggregate GB/s, with and without binding each rank to its GPU's
NUMA node (orbx_bind_thread_to_device), plus the box topology the numbers depend on.

    python tools/h2d_ceiling.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/h2d_ceiling.py --gpus 8
"""
import argparse
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                                      # noqa: E402
from orbslam2_with_quadrics_b200 import sharding                  # noqa: E402


def measure(local, nbytes, reps, chunks):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(7)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    step = nbytes // chunks
    for _ in range(3):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sharding.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        for c in range(chunks):
            dev[c * step:(c + 1) * step].copy_(host[c * step:(c + 1) * step], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    sharding.barrier()
    del host, dev
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--bytes", type=int, default=64 * 1920 * 1080)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--chunks", type=int, default=4, help="copies per step (the bench sends 4 sub-batches per step)")
    args = ap.parse_args()
    real_stdout = os.fdopen(os.dup(1), "w")       # NCCL prints its banner on fd 1: keep the JSON line on a private copy
    os.dup2(2, 1)
    rank, world, local = sharding.init_from_env()
    torch.cuda.set_device(local)
    cpus_before = sorted(os.sched_getaffinity(0))
    out = {}
    for mode in ("unbound", "numa_bound"):
        node, nb = -1, 0
        if mode == "numa_bound":
            node, nb = sharding.bind_to_gpu_numa(local)
        ms = measure(local, args.bytes, args.reps, args.chunks)
        gbs = args.bytes * args.reps / (ms / 1e3) / 1e9
        ms_max = sharding.max_over_ranks(ms)
        per_rank = [None] * world
        if world > 1:
            t = torch.zeros(world, dtype=torch.float64, device="cuda")
            t[rank] = gbs
            torch.distributed.all_reduce(t)
            per_rank = [round(float(x), 2) for x in t.tolist()]
            nodes = torch.zeros(world, dtype=torch.float64, device="cuda")
            nodes[rank] = node
            torch.distributed.all_reduce(nodes)
            node_list = [int(x) for x in nodes.tolist()]
        else:
            per_rank = [round(gbs, 2)]
            node_list = [node]
        out[mode] = {"aggregate_gbs": round(world * args.bytes * args.reps / (ms_max / 1e3) / 1e9, 2),
                     "per_rank_gbs": per_rank, "gpu_numa_node": node_list if mode == "numa_bound" else None,
                     "cpus_bound_rank0": nb if mode == "numa_bound" else None}
    if rank == 0:
        topo = None
        try:
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        except Exception:
            pass
        numa = {}
        base = "/sys/devices/system/node"
        if os.path.isdir(base):
            for d in sorted(os.listdir(base)):
                if d.startswith("node") and d[4:].isdigit():
                    try:
                        numa[d] = open(os.path.join(base, d, "cpulist")).read().strip()
                    except OSError:
                        pass
        line = {"tool": "h2d_ceiling", "n_gpus": world, "bytes_per_step": args.bytes, "chunks_per_step": args.chunks,
                "reps": args.reps, "frames_per_s_if_1080p": None, "result": out,
                "cpus_allowed": "%d (%d..%d)" % (len(cpus_before), cpus_before[0], cpus_before[-1]), "numa_cpulists": numa,
                "topo": topo}
        best = max(out[m]["aggregate_gbs"] for m in out)
        line["frames_per_s_if_1080p"] = round(best * 1e9 / (1920 * 1080), 0)
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
