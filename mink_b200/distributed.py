"""Data-parallel sharding of independent IK instances over ranks (one process per GPU).

The path has no exchange step: every instance (q, targets) is independent, so a step
needs NO collective.  The only communication offered is an optional all-gather of `dq` (and status)
for callers that want the whole batch on every rank -- `torch.distributed` over NCCL on GPUs
(NVLink/NVSwitch), gloo on CPU for tests.
"""

from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `total` instances for `rank` (first `total % world` ranks
    carry one extra instance)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, rank: int = None, world: int = None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def all_gather_rows(local: torch.Tensor, total: int) -> torch.Tensor:
    """Concatenate per-rank row blocks (ragged allowed) into the full [total, ...] tensor on every rank."""
    world = dist.get_world_size()
    if world == 1:
        return local
    sizes = [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]
    width = max(sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)
