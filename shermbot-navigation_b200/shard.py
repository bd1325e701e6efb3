"""Sharding of independent filters / scans over the GPUs of one box.

Filters (configs 2, 4) and scans (config 3) are independent units with no exchange step: rank r of W owns the contiguous block
``shard_range(B, r, W)``; the ONLY communication of a run is the gather of final states and the all-reduce of error statistics
at the end (NCCL over NVLink on GPUs; the same code runs over gloo on CPU tensors for the tests). One process per GPU,
launched by torchrun; nothing here touches the data path.
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `total` units owned by `rank`; sizes differ by at most one, earlier ranks take the extras."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_states(x_local, total: int, group=None):
    """All-gather of the per-rank state blocks (shape [n_local, len]) into the full [total, len] tensor, rank order = filter order.
    Blocks may differ by one row (shard_range): they are padded to the largest block for the collective and trimmed after."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x_local
    world = dist.get_world_size(group)
    nmax = (total + world - 1) // world
    pad = torch.zeros((nmax,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    pad[: x_local.shape[0]] = x_local
    out = torch.empty((world * nmax,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        parts.append(out[r * nmax: r * nmax + (hi - lo)])
    return torch.cat(parts, dim=0)


def allreduce_stats(stats, group=None):
    """Sum of a small fp64 statistics vector (squared errors, NEES, association mismatches, status counters) over the ranks."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats
