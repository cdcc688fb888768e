"""Multi-GPU host logic: camera streams are independent, so they are partitioned over ranks
(one process per GPU) with NO data-path collective; torch.distributed is used
only to align the timing window (barrier) and to take the max elapsed time over ranks."""
from __future__ import annotations

import os
from typing import List

import torch
import torch.distributed as dist


def rank_streams(n_streams: int, world_size: int, rank: int) -> List[int]:
    """Stream s is served by rank s mod world_size (a stereo pair is one stream: both images stay on one GPU)."""
    return [s for s in range(n_streams) if s % world_size == rank]


def init_from_env(backend: str | None = None):
    """Initialises torch.distributed from RANK/WORLD_SIZE/MASTER_* when launched by torchrun.
    Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29517")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value: float) -> float:
    """Slowest rank's time: the only number multi-GPU throughput may be computed from."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank: float, elapsed_s_this_rank: float) -> float:
    """Whole-job units/s: all ranks' units divided by the slowest rank's time."""
    return sum_over_ranks(units_this_rank) / max_over_ranks(elapsed_s_this_rank)


def bind_to_gpu_numa(local_rank: int):
    """Binds the calling thread (call it first thing in a rank: threads and pinned allocations made afterwards inherit
    it) to the CPUs of the NUMA node GPU `local_rank` hangs off, through the library's own orbx_bind_thread_to_device
    (host code stays behind the C ABI).  Returns (numa_node, cpus_bound); cpus_bound == 0 means the affinity was left
    alone (single-node box, or the container exposes no CPU of that node)."""
    import ctypes as C
    from . import _capi
    node, ncpus = C.c_int(-1), C.c_int(0)
    _capi.check(_capi.lib().orbx_bind_thread_to_device(int(local_rank), C.byref(node), C.byref(ncpus)))
    return node.value, ncpus.value
