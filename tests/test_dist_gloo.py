"""Multi-rank host logic on CPU: world_size-2 gloo.  The path shards by batch with no data-path
collective; the only exchange is the scalar loss / count all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dmhomo_b200 import dist as ddist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    r, _, w = ddist.init("gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(230)
    per_sample = torch.rand(B, generator=g)                 # the same "per-sample losses" on every rank
    lo, hi = ddist.shard_range(B, rank, world)
    mine = ddist.shard(per_sample, rank, world)
    assert mine.numel() == hi - lo
    glob = ddist.global_mean_loss(mine.mean(), mine.numel())
    v = torch.tensor([mine.sum().double(), float(mine.numel())], dtype=torch.float64)
    ddist.all_reduce_sums(v)
    q.put((rank, lo, hi, float(glob), float(v[0] / v[1]), float(per_sample.mean())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_loss_matches_single_process_mean():
    world, B = 2, 13                                         # uneven shards: 7 + 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 7), (7, 13)]  # contiguous slices, first ranks take the remainder
    for r in res:
        assert abs(r[3] - r[5]) < 1e-6 and abs(r[4] - r[5]) < 1e-6


def test_shard_range_partitions_exactly():
    for n in (1, 7, 64, 4096, 8191):
        for world in (1, 2, 4, 8):
            spans = [ddist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
