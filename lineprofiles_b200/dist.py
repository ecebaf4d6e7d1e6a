"""Multi-GPU plumbing: independent parameter sets are sharded over ranks, each rank evaluates its contiguous
block, and the spectra are gathered.  There is no other exchange on this path -- every set reads
the shared immutable table and writes its own spectrum (reference Q18: all lookups are pure).

The product path is the C ABI's klp_group (include/klineprofiles_b200.h): `rank_group` forms it from an initialised
torch.distributed job (one process per GPU; torch only carries the 128-byte NCCL id to the other ranks) and
`evaluate_sharded_device` runs the chunked evaluation whose NCCL gather overlaps the next chunk's kernels.
`evaluate_sharded` is the generic host-side logic (shard bounds, padding, one all-gather) with a pluggable evaluator;
the CPU tests drive it over gloo."""
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


def rank_group(table, group=None):
    """The klp_group of a torch.distributed job with one process per GPU: rank 0 creates the NCCL id, torch broadcasts
    the 128 bytes, every rank joins with its table replica."""
    import torch.distributed as dist

    from . import api

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    box = [api.nccl_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return api.Group.from_rank(table, box[0], rank, world)


def evaluate_sharded_device(kgroup, model, params, d_energy, n_flux: int, precision: str = "f32", d_spectrum=None,
                            stream=None):
    """`params` ([B, P] CUDA tensor, identical on every rank) evaluated as shards through the C ABI's group entry:
    rank r integrates sets [r * per, (r+1) * per), per = ceil(B / nranks) (short shards are padded with copies of the last
    set so that every rank moves the same amount), chunk by chunk with the NCCL gather of one chunk overlapping the
    kernels of the next.  Returns the [B, n_flux] spectra on every rank."""
    import torch

    world, rank = kgroup.nranks, kgroup.rank
    B = params.shape[0]
    per = -(-B // world)
    lo, hi = shard_bounds(B, world, rank)
    shard = params[lo:hi]
    if hi - lo < per:   # pad: the collective needs equal shards
        pad = params[-1:].expand(per - (hi - lo), -1)
        shard = torch.cat([shard, pad], 0)
    shard = shard.contiguous()
    full = torch.empty((world * per, n_flux), dtype=torch.float64, device=params.device)
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    kgroup.eval_allgather_device(model, [shard], [d_energy], n_flux, [full], per,
                                 d_spectrum=[d_spectrum] if d_spectrum is not None else None, precision=precision, streams=[st])
    if world * per == B:
        return full
    # drop the padded rows of the short shards
    keep = torch.cat([torch.arange(r * per, r * per + (shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0]),
                                   device=params.device) for r in range(world)])
    return full.index_select(0, keep)
