"""Multi-GPU plumbing: independent parameter sets are sharded over ranks (one process per GPU), each
rank evaluates its contiguous block, and ONE collective gathers the spectra.  There
is no other exchange on this path -- every set reads the shared immutable table and writes its own
spectrum (reference Q18: all lookups are pure)."""
from __future__ import annotations

from typing import Callable, Tuple


def shard_bounds(n_sets: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`: ceil(B/R) sets per rank, the tail ranks may be short/empty."""
    per = -(-n_sets // world_size)
    lo = min(rank * per, n_sets)
    return lo, min(lo + per, n_sets)


def evaluate_sharded(evaluate: Callable, params, n_flux: int, group=None, out=None):
    """Evaluate `params` ([B, P] torch tensor, identical on every rank) as shards and all-gather.

    `evaluate(shard_params, shard_out)` fills shard_out ([b, n_flux], same device) for its block.
    Returns the full [B, n_flux] tensor on every rank (padded rows of short shards are dropped).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B = params.shape[0]
    per = -(-B // world)
    lo, hi = shard_bounds(B, world, rank)
    local = torch.zeros((per, n_flux), dtype=torch.float64, device=params.device)
    if hi > lo:
        evaluate(params[lo:hi].contiguous(), local[: hi - lo])
    full = torch.empty((world * per, n_flux), dtype=torch.float64, device=params.device)
    dist.all_gather_into_tensor(full, local, group=group)
    res = full[:B]
    if out is not None:
        out.copy_(res)
        return out
    return res
