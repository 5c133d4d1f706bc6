"""Batch sharding of the codec path across the GPUs of one box.

Clips (and streams) are independent and the weights are ~50 MB, so the path is partitioned
along the batch dimension only: one process per GPU (torchrun), each rank runs the whole
encode -> RVQ -> decode on its contiguous shard, and there is NO collective inside the
forward.  `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is used
only for the optional gather of results and for reducing timing counters.
"""
from __future__ import annotations

import typing as tp

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> tp.Tuple[int, int]:
    """Contiguous [lo, hi) of `total` items owned by `rank`; the remainder goes to the first ranks."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world: int) -> tp.List[int]:
    return [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]


def _world(group=None) -> tp.Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def forward_sharded(compute: tp.Callable[[torch.Tensor], tp.Tuple[torch.Tensor, torch.Tensor]], x: torch.Tensor,
                    gather: bool = True, group=None):
    """Run `compute(x_shard) -> (indices[n,b,F], wav[b,1,T])` on this rank's shard of the full
    batch `x` [B,1,T] (every rank holds, or can synthesise, the full input) and optionally
    all-gather the results back into batch order.  Returns (indices, wav) for the full batch when
    gathering, else for the local shard."""
    rank, world = _world(group)
    B = x.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    idx, wav = compute(x[lo:hi])
    if not gather or world == 1:
        return idx, wav
    sizes = shard_sizes(B, world)
    n, b_loc, F = idx.shape
    T = wav.shape[2]
    smax = max(sizes)
    # all_gather wants equal shapes: pad the batch dim of ragged shards to the largest one
    idx_pad = torch.zeros(n, smax, F, dtype=idx.dtype, device=idx.device)
    wav_pad = torch.zeros(smax, 1, T, dtype=wav.dtype, device=wav.device)
    idx_pad[:, :b_loc] = idx
    wav_pad[:b_loc] = wav
    idx_parts = [torch.empty_like(idx_pad) for _ in sizes]
    wav_parts = [torch.empty_like(wav_pad) for _ in sizes]
    dist.all_gather(idx_parts, idx_pad, group=group)
    dist.all_gather(wav_parts, wav_pad, group=group)
    return (torch.cat([t[:, :s] for t, s in zip(idx_parts, sizes)], dim=1),
            torch.cat([t[:s] for t, s in zip(wav_parts, sizes)], dim=0))


def reduce_max(values: tp.Sequence[float], device=None, group=None) -> tp.List[float]:
    """MAX over ranks of a few scalars (device-timed milliseconds): a multi-GPU step is as slow
    as its slowest rank."""
    rank, world = _world(group)
    if world == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(v) for v in t]


def reduce_sum(values: tp.Sequence[float], device=None, group=None) -> tp.List[float]:
    rank, world = _world(group)
    if world == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(v) for v in t]
