"""N > 1 path of bench.py on CPU: the world -> rank partition and the job's only collective (MAX of times, SUM of work), run
with world_size 2 over gloo (the GPU job uses the same code over NCCL)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("total,world_size", [(4096, 1), (4096, 2), (4096, 8), (10, 4), (3, 8), (1, 2)])
def test_shard_worlds_partitions_every_world_once(total, world_size):
    covered = []
    counts = []
    for rank in range(world_size):
        first, n = bench.shard_worlds(total, rank, world_size)
        covered.extend(range(first, first + n))
        counts.append(n)
    assert covered == list(range(total)), "contiguous blocks in rank order, every world exactly once"
    assert max(counts) - min(counts) <= 1, "balanced to within one world"


def _worker(rank, world_size, port, total_worlds, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    first, n = bench.shard_worlds(total_worlds, rank, world_size)
    # a rank's "measurement": its time grows with its share, its work is its share
    times = (0.5 + 0.001 * n, 0.6 + rank, 0.7)
    work = (n * 1240, 100 + rank, n)
    t, w = bench.reduce_job(dist, world_size, times, work, "cpu")
    out.put((rank, first, n, t, w))
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_job_world_size_2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    total = 101
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, out)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (r0, f0, n0, t0, w0), (r1, f1, n1, t1, w1) = res
    assert (f0, n0, f1, n1) == (0, 51, 51, 50)
    assert t0 == t1 and w0 == w1, "every rank sees the same reduced job"
    assert t0 == pytest.approx([0.5 + 0.001 * 51, 1.6, 0.7]), "times: max over ranks"
    assert w0 == pytest.approx([101 * 1240, 201, 101]), "work: sum over ranks"
