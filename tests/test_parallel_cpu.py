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


def _layout(L=3, d=8):
    """A flat buffer laid out by tensor kind (all layers of one kind contiguous), like the engine's gradient buffer."""
    kinds = [("w_big", 40 * d), ("bias", d), ("w_mid", 10 * d), ("gain", d)]
    off, per_layer = 0, [[] for _ in range(L)]
    for _, n in kinds:
        for l in range(L):
            per_layer[l].append((off + l * n, n))
        off += L * n
        off = (off + 31) // 32 * 32  # sections are 128-byte aligned
    top_level = off  # non-block parameters live after the block sections
    return per_layer, off + 100, top_level


def test_grad_bucket_plan_tiles_the_flat_buffer_exactly_once():
    per_layer, total, _ = _layout()
    buckets, tail = P.plan_grad_buckets(per_layer, total, min_bucket=64)
    assert all(len(b) == 2 for b in buckets)  # w_big and w_mid of each layer; biases and gains ride in the tail
    seen = torch.zeros(total, dtype=torch.int32)
    for off, n in [sp for b in buckets for sp in b] + tail:
        seen[off: off + n] += 1
    assert bool((seen == 1).all())
    # adjacent tensors of one layer merge into a single bucket (q, k, v weights are contiguous in the engine)
    merged, _ = P.plan_grad_buckets([[(0, 50), (50, 50), (100, 28)]], 128, min_bucket=100)
    assert merged == [[(0, 128)]]
    with pytest.raises(ValueError):
        P.plan_grad_buckets([[(0, 100)], [(50, 100)]], 200, min_bucket=10)


def _bucket_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        per_layer, total, _ = _layout()
        buckets, tail = P.plan_grad_buckets(per_layer, total, min_bucket=64)
        grads = [torch.randn(total, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        flat = grads[rank].clone()
        for layer in range(len(buckets) - 1, -1, -1):  # same order GradAllReduce issues them in
            for off, n in buckets[layer]:
                dist.all_reduce(flat[off: off + n])
        for off, n in tail:
            dist.all_reduce(flat[off: off + n])
        flat /= world  # gloo has no AVG
        want = torch.stack(grads).sum(0) / world
        assert torch.allclose(flat, want, rtol=0, atol=1e-6)
    finally:
        dist.destroy_process_group()


def test_two_rank_bucketed_gradient_average_equals_mean():
    mp.spawn(_bucket_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_shard_plan_covers_the_flat_buffer_and_splits_evenly():
    per_layer, total, _ = _layout()
    big = [[sp for sp in lay if sp[1] >= 64] for lay in per_layer]  # what the engine would list as sharded tensors
    layers, tail = P.plan_shards(big, total)
    seen = torch.zeros(total, dtype=torch.int32)
    for off, n in [sp for lay in layers for sp in lay] + tail:
        seen[off: off + n] += 1
    assert bool((seen == 1).all())
    for world in (2, 4, 8):
        for off, n in layers[0]:
            parts = [P.shard_span(off, n, r, world) for r in range(world)]
            assert parts[0][0] == off and sum(m for _, m in parts) == n
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(world - 1))
    with pytest.raises(ValueError):
        P.shard_span(0, 10, 0, 4)
    with pytest.raises(ValueError):
        P.plan_shards([[(0, 100)], [(50, 100)]], 200)


def _adam_like(p, g, m, v):
    """Any element-wise optimizer arithmetic (what matters: every element sees the same inputs in both schedules)."""
    m.mul_(0.9).add_(g, alpha=0.1)
    v.mul_(0.95).addcmul_(g, g, value=0.05)
    p.mul_(1 - 1e-3).addcdiv_(m, v.sqrt() + 1e-8, value=-1e-2)


def _sharded_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        per_layer, total, _ = _layout()
        layers, tail = P.plan_shards([[sp for sp in lay if sp[1] >= 64] for lay in per_layer], total)
        gen = torch.Generator().manual_seed(3)
        p0 = torch.randn(total, generator=gen)
        # replicated schedule: all-reduce(mean) everything, every rank updates everything
        pa, ma, va = p0.clone(), torch.zeros(total), torch.zeros(total)
        # sharded schedule: reduce-scatter, update the own 1/world of every sharded span, all-gather the new values
        pb, mb, vb = p0.clone(), torch.zeros(total), torch.zeros(total)
        for step in range(3):
            g_local = torch.randn(total, generator=torch.Generator().manual_seed(100 * step + rank))
            ga = g_local.clone()
            dist.all_reduce(ga)
            ga /= world
            _adam_like(pa, ga, ma, va)
            gb = g_local.clone()
            for layer in range(len(layers) - 1, -1, -1):
                for off, n in layers[layer]:
                    mine = P.reduce_scatter_mean(gb[off: off + n], rank, world)
                    so, sn = P.shard_span(off, n, rank, world)
                    assert mine.data_ptr() == gb[so: so + sn].data_ptr()
                    _adam_like(pb[so: so + sn], mine, mb[so: so + sn], vb[so: so + sn])
                    P.all_gather_in_place(pb[off: off + n], rank, world)
            for off, n in tail:
                dist.all_reduce(gb[off: off + n])
                gb[off: off + n] /= world
                _adam_like(pb[off: off + n], gb[off: off + n], mb[off: off + n], vb[off: off + n])
        assert torch.equal(pa, pb)  # identical weights on every rank, bit for bit (each element updated once, by one rank)
        for buf_a, buf_b in ((ma, mb), (va, vb)):  # moments: current on the owner until gathered
            for lay in layers:
                for off, n in lay:
                    P.all_gather_in_place(buf_b[off: off + n], rank, world)
            assert torch.equal(buf_a, buf_b)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_update_equals_replicated_update():
    mp.spawn(_sharded_worker, args=(2, _free_port()), nprocs=2, join=True)
