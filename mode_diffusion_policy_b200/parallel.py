"""Data-parallel host logic for inference: trajectories are independent, so the batch is sharded
across ranks with NO data-path collective; torch.distributed is only used for the timing barrier, the max-over-ranks
reduction of device times and (optionally) gathering the (B, 10, 7) action chunks for a single caller."""
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
