"""Data-parallel host logic.

Inference: trajectories are independent, so the batch is sharded across ranks with NO data-path collective;
torch.distributed is only used for the timing barrier, the max-over-ranks reduction of device times and (optionally)
gathering the (B, 10, 7) action chunks for a single caller.

Training: the one exchange step is the all-reduce (mean) of the engine's flat gradient buffer. `GradAllReduce` issues it
bucketed by layer on a side stream, gated by the engine's per-layer "gradients final" events, so NCCL moves layer l's
weight gradients over NVLink while the backward of layers l-1..0 is still running (what DDP's reducer hooks do in the
reference, mode/training_calvin.py:97). `ShardedGradExchange` splits that all-reduce around the optimizer:
reduce-scatter -> each rank updates 1/world of every large tensor -> all-gather of the bf16 weights (half the bytes of
the fp32 gradients), which also divides the HBM-bound AdamW pass by the number of ranks."""
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


def shard_span(off: int, numel: int, rank: int, world: int) -> tuple[int, int]:
    """Rank `rank`'s contiguous 1/world of the span (offset, numel); numel must divide evenly."""
    if numel % world:
        raise ValueError(f"span of {numel} elements does not split over {world} ranks")
    n = numel // world
    return off + rank * n, n


def plan_shards(tensors_per_layer: list[list[tuple[int, int]]], total: int):
    """Exchange plan of the sharded optimizer: tensors_per_layer[l] = (offset, numel) of block l's sharded tensors (what
    `mode_optimizer_shard_tensors` lists). Returns (per-layer spans sorted by offset, tail) where `tail` are the spans of
    [0, total) no sharded tensor covers (replicated parameters: all-reduced as before)."""
    layers = [sorted(t) for t in tensors_per_layer]
    tail, cur = [], 0
    for off, n in sorted(sp for lay in layers for sp in lay):
        if off < cur:
            raise ValueError("overlapping sharded tensors")
        if off > cur:
            tail.append((cur, off - cur))
        cur = off + n
    if cur < total:
        tail.append((cur, total - cur))
    return layers, tail


def reduce_scatter_mean(span: torch.Tensor, rank: int, world: int, group=None) -> torch.Tensor:
    """In place: afterwards this rank's 1/world slice of `span` holds the mean over ranks (the other slices are
    unspecified). NCCL: one reduce_scatter(AVG) whose output aliases its slot of the input; other backends (the CPU
    tests run gloo, which has neither reduce_scatter nor AVG): all_reduce + divide."""
    off, n = shard_span(0, span.numel(), rank, world)
    mine = span[off: off + n]
    if span.is_cuda:
        dist.reduce_scatter_tensor(mine, span, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(span, group=group)
        mine /= world
    return mine


def all_gather_in_place(span: torch.Tensor, rank: int, world: int, group=None) -> None:
    """Every rank contributes its 1/world slice of `span`; afterwards all of `span` is current everywhere."""
    off, n = shard_span(0, span.numel(), rank, world)
    if span.is_cuda:
        dist.all_gather_into_tensor(span, span[off: off + n], group=group)
    else:
        parts = [torch.empty(n, dtype=span.dtype) for _ in range(world)]
        dist.all_gather(parts, span[off: off + n].clone(), group=group)
        for r, part in enumerate(parts):
            span[r * n: (r + 1) * n] = part


class ShardedGradExchange(GradAllReduce):
    """Exchange half of the sharded data-parallel step (`optim.EngineAdamW.step_sharded`).

    Per block, on the side stream: reduce-scatter(mean) of the block's large gradient tensors, so that rank r holds the
    averaged gradient of elements [r, r+1) * numel / world of each; after that rank's optimizer launch wrote its part
    of the new weights (bf16, staging buffer), all-gather of the staging spans. Everything else (small tensors,
    non-block parameters, tensors that do not split evenly) is all-reduced at the end like `GradAllReduce.reduce_tail`.
    Compared with all-reduce + a replicated update, the wire carries 4 + 2 instead of 4 + 4 bytes per parameter and every
    rank runs 1/world of the HBM-bound optimizer pass."""

    def __init__(self, engine, param_names, n_layers: int, group=None):
        self.engine, self.group, self.n_layers = engine, group, n_layers
        self.param_names = list(param_names)
        self.flat = engine.flat_grads()
        self.stream = torch.cuda.Stream(device=self.flat.device)
        self.coalesce = hasattr(dist, "_coalescing_manager")
        self.rank = dist.get_rank(group) if self.active() else 0
        self.world = dist.get_world_size(group) if self.active() else 1
        self.layers = None  # planned by prepare(), once the optimizer has bound its parameters
        self.fallback = None  # a GradAllReduce when no tensor can be sharded over this world size
        import os
        # Replicated per-block tensors of at least this many elements would get their own all-reduce during the backward.
        # Off by default: on 8 GPUs the 12 extra collectives (the router's first Linear, 4 MB per block) cost more than
        # the shorter tail saves (14.1 vs 13.7 ms/step, profiles/r02_train_sharded.log); they ride in the tail instead.
        self.min_replicated_bucket = int(os.environ.get("MODE_SHARD_MIN_REPLICATED_BUCKET", 1 << 62))

    def prepare(self) -> None:
        """Switch the engine's optimizer to sharded groups and read back which tensors it shards (idempotent)."""
        if self.layers is not None:
            return
        eng = self.engine
        eng.set_optimizer_sharding(self.rank, self.world)
        self.layers, _ = plan_shards([eng.optimizer_shard_tensors(l) for l in range(self.n_layers)], self.flat.numel())
        if not any(self.layers):
            # nothing splits evenly over this many ranks (the 1024-element blocks of the large tensors do not divide by the
            # world size): keep the replicated, pipelined all-reduce step
            eng.set_optimizer_sharding(0, 1)
            self.fallback = GradAllReduce(eng, self.param_names, self.n_layers, group=self.group)
            self.fallback.stream = self.stream
            self.tail = self.fallback.tail
            return
        self.staging = eng.optimizer_staging()
        by_offset = {eng.grad_range(n)[0]: n for n in self.param_names}
        self.names = [[by_offset[off] for off, _ in lay] for lay in self.layers]  # parameter of every sharded span
        # replicated per-block tensors that get their own all-reduce during the backward instead of riding in the tail
        # after it (none by default, see min_replicated_bucket)
        sharded = {off for lay in self.layers for off, _ in lay}
        self.layer_replicated = [[] for _ in range(self.n_layers)]
        for name in self.param_names:
            if name.startswith("blocks."):
                off, n = eng.grad_range(name)
                if off not in sharded and n >= self.min_replicated_bucket:
                    self.layer_replicated[int(name.split(".")[1])].append((off, n))
        covered = [sorted(self.layers[l] + self.layer_replicated[l]) for l in range(self.n_layers)]
        _, self.tail = plan_shards(covered, self.flat.numel())

    def _spans(self, buf, layer):
        return [buf[off: off + n] for off, n in self.layers[layer]]

    def _grouped(self, fn, views) -> None:
        """One NCCL group launch for a block's ~12 spans."""
        if not views:
            return
        with torch.cuda.stream(self.stream):
            if self.coalesce and len(views) > 1:
                with dist._coalescing_manager(group=self.group, device=self.flat.device, async_ops=False):
                    for v in views:
                        fn(v)
            else:
                for v in views:
                    fn(v)

    def reduce_scatter_layer(self, layer: int) -> None:
        self.engine.wait_grads(layer, self.stream)
        self._grouped(lambda v: reduce_scatter_mean(v, self.rank, self.world, self.group), self._spans(self.flat, layer))
        self._reduce_spans(self.layer_replicated[layer])

    def gather_weights_layer(self, layer: int) -> None:
        """All-gather block `layer`'s bf16 staging spans (enqueue after that block's optimizer launch on this stream)."""
        self._grouped(lambda v: all_gather_in_place(v, self.rank, self.world, self.group), self._spans(self.staging, layer))

    def gather_flat(self, buf: torch.Tensor) -> None:
        """All-gather every sharded span of a flat buffer with the gradient layout (moments, EMA) on the side stream."""
        for layer in range(self.n_layers):
            self._grouped(lambda v: all_gather_in_place(v, self.rank, self.world, self.group), self._spans(buf, layer))

    def gather_parameters(self, params: dict) -> None:
        """All-gather the fp32 masters of the sharded tensors (each rank only updated its part) on the side stream."""
        for layer in range(self.n_layers):
            views = [params[n].data.view(-1) for n in self.names[layer]]
            self._grouped(lambda v: all_gather_in_place(v, self.rank, self.world, self.group), views)

    def run(self) -> None:
        raise RuntimeError("ShardedGradExchange is driven by optim.EngineAdamW.step_sharded")
