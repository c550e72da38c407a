"""Partitioning of the independent (prompt x view-group) units across the ranks of one node.

Each (prompt, CFG-pair) group of Nv*F latent images never interacts with another group -- cross-view attention regroups
inside a group (attention_processor.py:340) and temporal attention / GroupNorm-over-frames are per (b, n) sample -- so the
batch axis shards with NO data-path collective.  torch.distributed is used only for the barrier, the max-over-ranks timing
and an optional gather of the finished latents."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def shard_units(num_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of `num_units` independent units owned by `rank` (balanced to within one unit)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(num_units, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_latents(local: torch.Tensor, world: int, group=None) -> List[torch.Tensor]:
    """All-gather the per-rank result latents [units_r, Nv, 4, F, h, w] (1 MB per prompt) -- control path only."""
    import torch.distributed as dist
    if world == 1:
        return [local]
    sizes = [torch.zeros(1, dtype=torch.long, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.long, device=local.device), group=group)
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros(mx, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return [o[: int(s)] for o, s in zip(outs, sizes)]


def max_over_ranks(value_ms: float, device, world: int) -> float:
    import torch.distributed as dist
    if world == 1:
        return value_ms
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
