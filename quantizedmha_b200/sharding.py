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
