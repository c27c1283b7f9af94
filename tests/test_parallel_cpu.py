"""world_size-2 gloo tests (CPU) of the data-parallel host logic used by bench.py --gpus N."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mode_diffusion_policy_b200 import parallel as P


def test_shard_bounds_cover_batch_exactly():
    for total in (1, 2, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [P.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank holds the same global tensor (seeded), takes its shard, "samples" (a per-trajectory function), gathers
        g = torch.Generator().manual_seed(0)
        x = torch.randn(total, 10, 7, generator=g)
        local = P.shard(x, rank, world)
        b, e = P.shard_bounds(total, rank, world)
        assert torch.equal(local, x[b:e])
        out = P.gather_actions(local * 2.0 + 1.0, total)
        assert torch.equal(out, x * 2.0 + 1.0)  # sharded result == single-process result, in global order
        # the job is as slow as its slowest rank
        ms = [10.0 + rank, 5.0 - rank]
        assert P.max_over_ranks(ms, "cpu") == [10.0 + world - 1, 5.0]
        value, ms_max = P.aggregate_throughput(units_per_rank=200.0, ms_local=100.0 * (rank + 1), device="cpu")
        assert ms_max == 100.0 * world
        assert abs(value - world * 200.0 / (0.1 * world)) < 1e-9
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [256, 7])
def test_two_rank_sharding_and_timing_reduction(total):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total), nprocs=2, join=True)
