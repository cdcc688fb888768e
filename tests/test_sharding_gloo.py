"""Multi-GPU host logic on CPU: two processes, gloo backend.  Streams are partitioned over
ranks with no data-path collective; only the timing window is reduced (max over ranks)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orbslam2_with_quadrics_b200 import sharding


def test_rank_streams_partition():
    for n, w in [(8, 1), (8, 2), (7, 4), (3, 8), (64, 8)]:
        seen = []
        for r in range(w):
            s = sharding.rank_streams(n, w, r)
            assert all(x % w == r for x in s)
            seen += s
        assert sorted(seen) == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, l = sharding.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    sharding.barrier()
    mine = sharding.rank_streams(5, world, rank)
    # each rank "processes" its own streams; rank 1 is slower
    frames = 10.0 * len(mine)
    elapsed = 1.0 + rank
    fps = sharding.aggregate_throughput(frames, elapsed)
    q.put((rank, mine, sharding.max_over_ranks(elapsed), fps))
    sharding.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_timing_reduction_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=100) for _ in ps)
    for p in ps:
        p.join(30)
        assert p.exitcode == 0
    assert out[0][1] == [0, 2, 4] and out[1][1] == [1, 3]
    for _, _, tmax, fps in out:
        assert tmax == 2.0                         # slowest rank
        assert abs(fps - 50.0 / 2.0) < 1e-9        # all ranks' frames / slowest rank's time


def test_single_process_is_identity():
    assert sharding.max_over_ranks(3.5) == 3.5 and sharding.aggregate_throughput(10, 2) == 5.0
