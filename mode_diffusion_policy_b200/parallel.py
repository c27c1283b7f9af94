"""Data-parallel host logic.

Inference: trajectories are independent, so the batch is sharded across ranks with NO data-path collective;
torch.distributed is only used for the timing barrier, the max-over-ranks reduction of device times and (optionally)
gathering the (B, 10, 7) action chunks for a single caller.

Training: the one exchange step is the all-reduce (mean) of the engine's flat gradient buffer. `GradAllReduce` issues it
bucketed by layer on a side stream, gated by the engine's per-layer "gradients final" events, so NCCL moves layer l's
weight gradients over NVLink while the backward of layers l-1..0 is still running (what DDP's reducer hooks do in the
reference, mode/training_calvin.py:97)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) slice of `total` trajectories for `rank`: sizes differ by at most one."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    b, e = shard_bounds(t.shape[0], rank, world)
    return t[b:e]


def max_over_ranks(values, device) -> list[float]:
    """Element-wise MAX over ranks of a list of floats (device times): the job is as slow as its slowest rank."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def gather_actions(local: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather ragged shards of the sampled actions back into global batch order (optional convenience)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    width = max(shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world))
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([out[r][: shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0]] for r in range(world)])


def aggregate_throughput(units_per_rank: float, ms_local: float, device) -> tuple[float, float]:
    """Whole-job units/s: all ranks' units divided by the slowest rank's time. Returns (value, ms_max)."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    (ms,) = max_over_ranks([ms_local], device)
    return world * units_per_rank / (ms * 1e-3), ms


def plan_grad_buckets(ranges_per_layer: list[list[tuple[int, int]]], total: int, min_bucket: int = 1 << 20):
    """Bucket plan for the overlapped gradient all-reduce.

    ranges_per_layer[l] = (offset, numel) of every gradient tensor of block l inside a flat buffer of `total` elements.
    Adjacent tensors are merged; merged spans of at least `min_bucket` elements become per-layer buckets (reduced as
    soon as that layer's backward is done); everything else — small tensors and the non-block parameters — is covered
    by `tail` spans reduced once at the end. Returns (layer_buckets, tail); together they tile [0, total) exactly once
    (alignment gaps between sections ride along with the tail: they are never written and stay zero)."""
    def merge(spans):
        out = []
        for off, n in sorted(spans):
            if n == 0:
                continue
            if out and out[-1][0] + out[-1][1] == off:
                out[-1] = (out[-1][0], out[-1][1] + n)
            else:
                out.append((off, n))
        return out

    layer_buckets = [[sp for sp in merge(r) if sp[1] >= min_bucket] for r in ranges_per_layer]
    big = sorted(sp for lb in layer_buckets for sp in lb)
    tail, cur = [], 0
    for off, n in big:
        if off < cur:
            raise ValueError("overlapping gradient ranges")
        if off > cur:
            tail.append((cur, off - cur))
        cur = off + n
    if cur < total:
        tail.append((cur, total - cur))
    return layer_buckets, tail


class GradAllReduce:
    """Overlapped data-parallel gradient averaging for the engine's training step.

        reducer = GradAllReduce(engine, [n for n, _ in inner.named_parameters()], n_layers)
        loss, _ = model.loss(...)      # enqueues forward + backward (no host sync)
        reducer.run()                  # per-layer buckets on the side stream, then the tail; main stream waits
        loss.backward(); opt.step()
    """

    def __init__(self, engine, param_names, n_layers: int, group=None, min_bucket: int = 1 << 20):
        self.engine, self.group, self.n_layers = engine, group, n_layers
        per_layer = [[] for _ in range(n_layers)]
        for name in param_names:
            if name.startswith("blocks."):
                per_layer[int(name.split(".")[1])].append(engine.grad_range(name))
        self.flat = engine.flat_grads()
        self.layer_buckets, self.tail = plan_grad_buckets(per_layer, self.flat.numel(), min_bucket)
        self.stream = torch.cuda.Stream(device=self.flat.device)
        import os
        self.coalesce = os.environ.get("MODE_ALLREDUCE_COALESCE", "1") == "1" and hasattr(dist, "_coalescing_manager")

    def active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _reduce_spans(self, spans) -> None:
        """All-reduce (mean) a list of (offset, numel) spans of the flat buffer as ONE NCCL group launch when the
        backend can coalesce (a block's buckets are ~5 spans of 4-128 MB: one launch instead of five)."""
        views = [self.flat[off: off + n] for off, n in spans]
        if not views:
            return
        with torch.cuda.stream(self.stream):
            if self.coalesce and len(views) > 1:
                with dist._coalescing_manager(group=self.group, device=self.flat.device, async_ops=False):
                    for v in views:
                        dist.all_reduce(v, op=dist.ReduceOp.AVG, group=self.group)
            else:
                for v in views:
                    dist.all_reduce(v, op=dist.ReduceOp.AVG, group=self.group)

    def reduce_layer(self, layer: int) -> None:
        """Enqueue block `layer`'s large buckets on the side stream, behind that block's gradient-ready event."""
        self.engine.wait_grads(layer, self.stream)
        self._reduce_spans(self.layer_buckets[layer])

    def reduce_tail(self) -> None:
        """Enqueue everything the per-layer buckets do not cover (small tensors, non-block parameters)."""
        self.engine.wait_grads(-1, self.stream)
        self._reduce_spans(self.tail)

    def run(self) -> None:
        if not self.active():
            return
        main = torch.cuda.current_stream(self.flat.device)
        for layer in range(self.n_layers - 1, -1, -1):  # the backward finishes the last block first
            self.reduce_layer(layer)
        self.reduce_tail()
        main.wait_stream(self.stream)
