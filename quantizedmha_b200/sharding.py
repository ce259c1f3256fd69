"""(batch × head) sharding of the attention forward across the B200s of one box.

Every (batch, head) pair is an independent attention problem (the reference already loops heads
independently, include/launchers.h:41-62), so the path shards with NO data-path collective:
rank r of W owns a contiguous range of the flattened unit index u = b*H + head.  Because the
tensor layout is [B, N, H*d], a contiguous unit range is a list of (batch, head-range) slabs.
"""
from __future__ import annotations

from typing import List, Tuple


def unit_range(units: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `units` items: first (units % world) ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_slabs(B: int, H: int, world: int, rank: int) -> List[Tuple[int, int, int]]:
    """[(b, h0, h1)] covering this rank's units in order; heads h0..h1-1 of batch b."""
    lo, hi = unit_range(B * H, world, rank)
    out = []
    u = lo
    while u < hi:
        b, h0 = divmod(u, H)
        h1 = min(H, h0 + (hi - u))
        out.append((b, h0, h1))
        u += h1 - h0
    return out


def slab_view(x, b: int, h0: int, h1: int, H: int):
    """View of heads [h0,h1) of batch b of a [B, N, H*d] array/tensor -> [N, (h1-h0)*d]."""
    d = x.shape[-1] // H
    return x[b, :, h0 * d:h1 * d]


def pack_units(out, B: int, H: int, world: int, rank: int):
    """This rank's output slabs of a [B, N, H*d] tensor as one contiguous [units_r, N, d] tensor."""
    import torch
    d = out.shape[-1] // H
    parts = [slab_view(out, b, h0, h1, H).reshape(out.shape[1], h1 - h0, d).permute(1, 0, 2)
             for (b, h0, h1) in shard_slabs(B, H, world, rank)]
    if not parts:
        return out.new_zeros((0, out.shape[1], d))
    return torch.cat(parts, dim=0).contiguous()


def gather_outputs(out, B: int, H: int, group=None):
    """Optional caller-side step, never part of the attention hot path: every rank
    holds valid data only in its own (batch, head) slabs of `out` [B, N, H*d]; after the call every
    rank holds the whole tensor.  One all-gather of unit-major [units, N, d] blocks — NCCL over
    NVLink 5 / NVSwitch for CUDA tensors, gloo for CPU tensors (tests).  Uneven splits are padded to
    the largest shard."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    N, d = out.shape[1], out.shape[-1] // H
    counts = [unit_range(B * H, world, r)[1] - unit_range(B * H, world, r)[0] for r in range(world)]
    cap = max(counts)
    mine = out.new_zeros((cap, N, d))
    mine[:counts[rank]] = pack_units(out, B, H, world, rank)
    bufs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine, group=group)
    for r in range(world):
        if r == rank:
            continue
        lo, _ = unit_range(B * H, world, r)
        for j in range(counts[r]):
            b, h = divmod(lo + j, H)
            out[b, :, h * d:(h + 1) * d] = bufs[r][j]
    return out
