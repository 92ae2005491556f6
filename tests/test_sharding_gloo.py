"""CPU, world_size 2 over gloo: the multi-rank plumbing used by bench.py / multi-GPU inference
(shard bounds, MAX-over-ranks timing, whole-job throughput, result gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from act3d_chained_diffuser_b200 import sharding


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 16, 33):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_bounds(5, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float32)[:, None] * torch.ones(1, 3)
        full = sharding.gather_rows(local, 5)
        slowest = sharding.max_over_ranks(1.0 + rank)
        thr = sharding.aggregate_throughput(units_this_rank=hi - lo, seconds_this_rank=1.0 + rank)
        dist.barrier()
        out.put((rank, full[:, 0].tolist(), slowest, thr))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, rows, slowest, thr in res:
        assert rows == [0.0, 1.0, 2.0, 3.0, 4.0]
        assert slowest == 2.0
        assert abs(thr - 5 / 2.0) < 1e-12
