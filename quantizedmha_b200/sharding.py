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


def forward_sharded(Q, K, V, H: int, kernel="int8", gran: int = -1, chunks: int = 4, group=None, forward_fn=None):
    """Sharded forward WITH a replicated result: Q, K, V [B, N, H*d] are present on every rank,
    rank r computes only its own (batch x head) units and every rank returns the whole output.  The rank's units are
    cut into `chunks` groups; the all-gather of a finished group (NCCL over NVLink 5 / NVSwitch for CUDA tensors) is
    enqueued on a side stream behind an event while the next group computes, so only the last group's transfer is
    exposed.  Never part of the attention hot path: callers that keep their shard use forward() on it directly.

    forward_fn(q, k, v, heads) -> [N, heads*d] computes one slab (default: the library's forward on CUDA tensors);
    the CPU tests inject a host implementation and run the same code over gloo."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B, N, d = Q.shape[0], Q.shape[1], Q.shape[2] // H
    units = B * H
    bounds = [unit_range(units, world, r) for r in range(world)]
    cap = max(hi - lo for lo, hi in bounds)
    chunks = max(1, min(chunks, cap))
    cu = -(-cap // chunks)                                # units per chunk (last chunk may be partly padding)
    lo, hi = bounds[rank]
    if forward_fn is None:
        from . import binding as qb
        forward_fn = lambda q, k, v, heads: qb.forward(q, k, v, heads, kernel=kernel, gran=gran)
    out = torch.empty_like(Q)
    cuda = Q.is_cuda
    comm = torch.cuda.Stream(device=Q.device) if cuda else None
    pending = []
    for c in range(chunks):
        mine = Q.new_zeros((cu, N, d))
        j0, j1 = c * cu, min((c + 1) * cu, hi - lo)        # this rank's unit offsets in the chunk
        u = lo + j0
        while u < lo + j1:                                 # slabs (b, h0..h1) inside the chunk
            b, h0 = divmod(u, H)
            h1 = min(H, h0 + (lo + j1 - u))
            sl = [slab_view(t, b, h0, h1, H).contiguous() for t in (Q, K, V)]
            o = forward_fn(sl[0], sl[1], sl[2], h1 - h0)   # [N, (h1-h0)*d]
            slab_view(out, b, h0, h1, H).copy_(o)
            mine[u - lo - j0:u - lo - j0 + (h1 - h0)] = o.reshape(N, h1 - h0, d).permute(1, 0, 2)
            u += h1 - h0
        recv = Q.new_empty((world * cu, N, d))
        if cuda:
            ready = torch.cuda.Event()
            ready.record()
            comm.wait_event(ready)
            with torch.cuda.stream(comm):
                work = dist.all_gather_into_tensor(recv, mine, group=group, async_op=True)
            mine.record_stream(comm)
            recv.record_stream(comm)
        else:                                              # CPU tensors (gloo, tests): list form of the same collective
            recv = [torch.empty_like(mine) for _ in range(world)]
            work = dist.all_gather(recv, mine, group=group, async_op=True)
        pending.append((c, recv, work))
    for c, recv, work in pending:
        work.wait()                                        # CUDA: makes the current stream wait for the collective
        for r in range(world):
            if r == rank:
                continue
            rlo, rhi = bounds[r]
            block = recv[r] if isinstance(recv, list) else recv[r * cu:(r + 1) * cu]
            j, jend = c * cu, min((c + 1) * cu, rhi - rlo)
            while j < jend:                                # one strided copy per (batch, head range) slab
                b, h0 = divmod(rlo + j, H)
                h1 = min(H, h0 + (jend - j))
                seg = block[j - c * cu:j - c * cu + (h1 - h0)]          # [heads, N, d]
                out[b, :, h0 * d:h1 * d] = seg.permute(1, 0, 2).reshape(N, (h1 - h0) * d)
                j += h1 - h0
    return out


# ---------------------------------------------------------------------------------------------------------
# Fused gather: the attention kernel's epilogue writes every finished output tile into the
# replicas of ALL ranks — TMA tensor stores into NVLink peer memory, issued tile by tile while the rest of the
# grid still computes — so the replicated result costs no all-gather, no extra pass over the output and no
# extra kernel.  NCCL only carries two tiny stream-ordered all-reduces that fence the replicas' reuse.

def launch_plan(B: int, H: int, world: int, rank: int) -> List[Tuple[int, int, int, int]]:
    """This rank's units as kernel launches [(b0, b1, h0, h1)]: heads h0..h1-1 of batch entries b0..b1-1.
    Consecutive whole batch entries are merged into one launch; partial head ranges stay one launch each."""
    plan: List[Tuple[int, int, int, int]] = []
    for b, h0, h1 in shard_slabs(B, H, world, rank):
        if plan and h0 == 0 and h1 == H and plan[-1][2] == 0 and plan[-1][3] == H and plan[-1][1] == b:
            plan[-1] = (plan[-1][0], b + 1, 0, H)
        else:
            plan.append((b, b + 1, h0, h1))
    return plan


def slab_offset(b0: int, h0: int, N: int, H: int, d: int) -> int:
    """Element offset of out[b0, 0, h0*d] inside a contiguous [B, N, H*d] tensor."""
    return (b0 * N * H + h0) * d


class ReplicatedOutput:
    """One [B, N, H*d] output tensor per rank, every rank's copy mapped into every other rank's address space
    (CUDA IPC handles exchanged through the process group; the mappings stay open until close()).
    `local` is this rank's tensor, `peer_base[r]` the device address of rank r's tensor in THIS process."""

    def __init__(self, B: int, N: int, H: int, d: int, dtype=None, device=None, group=None):
        import torch
        import torch.distributed as dist
        from . import binding as qb
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world - 1 > qb.MAX_PEERS:
            raise qb.QmhaError(f"at most {qb.MAX_PEERS + 1} ranks")
        self.B, self.N, self.H, self.d = B, N, H, d
        self.local = torch.empty((B, N, H * d), dtype=dtype or torch.float32, device=device)
        # Every step that can fail on one rank only is followed by an exchange of the outcome, so that all ranks raise
        # together instead of leaving the others inside a collective.
        try:
            mine = qb.ipc_export(self.local)
        except qb.QmhaError as e:
            mine = ("error", str(e))
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        bad = [f"rank {r}: {h[1]}" for r, h in enumerate(handles) if h[0] == "error"]
        if bad:
            raise qb.QmhaError("ReplicatedOutput: export failed (" + "; ".join(bad) + ")")
        self.peer_base, err = {}, None
        try:
            for r, (h, off) in enumerate(handles):
                if r != self.rank:
                    self.peer_base[r] = qb.ipc_open(h, off)
        except qb.QmhaError as e:
            err = str(e)
        outcomes = [None] * self.world
        dist.all_gather_object(outcomes, err, group=group)
        bad = [f"rank {r}: {o}" for r, o in enumerate(outcomes) if o]
        if bad:
            self.close()
            raise qb.QmhaError("ReplicatedOutput: mapping a peer's replica failed (" + "; ".join(bad) + ")")
        self._fence = torch.zeros(1, dtype=torch.int32, device=self.local.device)

    def fence(self):
        """Stream-ordered rendezvous of all ranks (one 4-byte NCCL all-reduce on the current stream): everything
        the ranks enqueued before it — kernels that write into the replicas, kernels that read them — is complete
        on every rank before anything enqueued after it starts."""
        import torch.distributed as dist
        dist.all_reduce(self._fence, op=dist.ReduceOp.MAX, group=self.group)   # (zeros stay zeros: nothing to overflow)

    def close(self):
        from . import binding as qb
        for addr in self.peer_base.values():      # only this object's mappings (they are reference-counted per allocation)
            qb.ipc_close(addr)
        self.peer_base = {}


def forward_fused_gather(Q, K, V, H: int, rep: "ReplicatedOutput", kernel="int8", gran: int = -1, forward_fn=None,
                         fence: bool = True):
    """Sharded forward with a replicated result and NO gather step: Q, K, V [B, N, H*d] are present on every
    rank, rank r computes only its own (batch x head) units and its attention kernel stores each finished tile
    into rep.local AND into every peer's replica (qmha_args.peer_O).  Returns rep.local, complete on the current
    stream once the trailing fence has passed.

    forward_fn(q, k, v, heads, out_view, peer_addrs) is injected by the CPU tests; the default is the library's
    forward on CUDA tensors."""
    B, N, d = Q.shape[0], Q.shape[1], Q.shape[2] // H
    if forward_fn is None:
        from . import binding as qb
        forward_fn = lambda q, k, v, heads, out_view, peers: qb.forward(q, k, v, heads, kernel=kernel, gran=gran,
                                                                         out=out_view, peer_outs=peers)
    out = rep.local
    esz = out.element_size()
    if fence:
        rep.fence()                      # nobody still reads the previous contents of any replica
    for b0, b1, h0, h1 in launch_plan(B, H, rep.world, rep.rank):
        q, k, v = (t[b0:b1, :, h0 * d:h1 * d] for t in (Q, K, V))   # strided slab views: read in place by the quantise pass
        view = out[b0:b1, :, h0 * d:h1 * d]
        off = slab_offset(b0, h0, N, H, d) * esz
        peers = [rep.peer_base[r] + off for r in sorted(rep.peer_base)]
        forward_fn(q, k, v, h1 - h0, view, peers)
    if fence:
        rep.fence()                      # every rank's stores into this replica have landed
    return out
