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


# ---------------------------------------------------------------------------------------------------------------------
# Rasterizer: cameras are independent given the gaussians (SURVEY 8e row "Rasterizer"): rank r renders cameras r::N of the
# batch, the gaussians + k-planes + MLPs are replicated, and ONE all-reduce per step sums the gradients of the learnable
# deformation field (k-plane grids + MLPs, ~2.8 MB fp32) -- what DDP does for the reference's `devices=-1` run (launch.py:115),
# here as a single flat bucket so the collective is launch-latency sized, not per-parameter.
# ---------------------------------------------------------------------------------------------------------------------
def shard_cameras(batch: dict, rank: int, world: int) -> dict:
    """Sub-batch of the cameras `rank::world` of a renderer batch (c2w / fovy / timestamps are per camera, the rest is
    shared).  The strided assignment balances the per-frame work (neighbouring cameras share a timestamp)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    bs = batch["c2w"].shape[0]
    idx = torch.arange(rank, bs, world, device=batch["c2w"].device)
    out = dict(batch)
    for k in ("c2w", "fovy", "timestamps", "elevation", "azimuth", "camera_distances"):
        if k in batch and torch.is_tensor(batch[k]) and batch[k].shape[:1] == (bs,):
            out[k] = batch[k][idx]
    out["camera_index"] = idx
    return out


def allreduce_gradients(params: Sequence[torch.nn.Parameter], world: int, group=None, average: bool = False) -> int:
    """Sum (or average) the .grad of `params` over the ranks with ONE all-reduce of a flat fp32 bucket; parameters without
    a gradient on this rank contribute zeros.  Returns the bucket size in bytes (0 when world == 1)."""
    if world == 1:
        return 0
    import torch.distributed as dist
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    dev = params[0].device
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        if p.grad is not None:
            flat[off:off + p.numel()] = p.grad.reshape(-1).float()
        off += p.numel()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
    off = 0
    for p in params:
        g = flat[off:off + p.numel()].reshape(p.shape).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += p.numel()
    return flat.numel() * 4


def gather_renders(local: torch.Tensor, rank: int, world: int, total: int, group=None) -> torch.Tensor:
    """Re-assemble [total, ...] images from the per-rank `rank::world` slices (forward all-gather of SURVEY 8e: the ranks
    running the UNet need every rendered view).  Not differentiable: the SDS gradient comes back through
    `scatter_render_grads`."""
    if world == 1:
        return local
    import torch.distributed as dist
    per = (total + world - 1) // world
    pad = torch.zeros(per, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local.detach()
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    full = torch.empty(total, *local.shape[1:], dtype=local.dtype, device=local.device)
    for r in range(world):
        n = len(range(r, total, world))
        full[r::world] = outs[r][:n]
    return full


def scatter_render_grads(full_grad: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Gradient w.r.t. this rank's renders out of the gradient w.r.t. the gathered batch (every rank computed the same
    SDS gradient on the full batch, or it was broadcast)."""
    return full_grad[rank::world]
