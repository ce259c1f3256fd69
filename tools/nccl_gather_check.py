"""Multi-GPU check under torchrun (NCCL): every rank computes ITS (batch x head) slabs of one problem through the
C-ABI, sharding.gather_outputs() replicates the result over NVLink (one all-gather, caller-side, never on the hot
path), and every rank compares the replicated tensor with a full single-GPU forward — bit-identical, because the
scales are per head or finer.  usage: python -m torch.distributed.run --nproc-per-node N tools/nccl_gather_check.py"""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantizedmha_b200 as qm
from quantizedmha_b200.sharding import (shard_slabs, slab_view, gather_outputs, forward_sharded, forward_fused_gather,
                                        ReplicatedOutput)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, H, N, d = 3, 5, 1000, 64          # 15 units: uneven split for every world size > 1
gen = torch.Generator(device=dev).manual_seed(1234)      # same inputs on every rank
q, k, v = (torch.rand((B, N, H * d), device=dev, generator=gen) for _ in range(3))
ok = True
for kernel, gran in (("int8", qm.GRAN_BLOCK), ("int8", qm.GRAN_HEAD), ("f16", qm.GRAN_HEAD)):
    out = torch.zeros_like(q)
    for (b, h0, h1) in shard_slabs(B, H, world, rank):
        sl = [slab_view(t, b, h0, h1, H).contiguous() for t in (q, k, v)]
        slab_view(out, b, h0, h1, H).copy_(qm.forward(*sl, h1 - h0, kernel=kernel, gran=gran))
    torch.cuda.synchronize()
    qm.binding.check_async_error()
    gather_outputs(out, B, H)
    full = qm.forward(q, k, v, H, kernel=kernel, gran=gran)
    torch.cuda.synchronize()
    same = bool(torch.equal(out, full))
    ok = ok and same
    if rank == 0:
        print(f"{kernel} gran={gran}: gathered == single-GPU forward: {same}", flush=True)
# sharded forward with the chunked gather that overlaps the transfer of finished units with the next units' compute
for chunks in (1, 3):
    out = forward_sharded(q, k, v, H, kernel="int8", gran=qm.GRAN_BLOCK, chunks=chunks)
    full = qm.forward(q, k, v, H, kernel="int8", gran=qm.GRAN_BLOCK)
    torch.cuda.synchronize()
    same = bool(torch.equal(out, full))
    ok = ok and same
    if rank == 0:
        print(f"forward_sharded(chunks={chunks}) == single-GPU forward: {same}", flush=True)
# fused gather: no collective at all — the attention epilogue stores every finished tile into the replicas of ALL
# ranks (TMA tensor stores into CUDA-IPC mapped peer memory over NVLink); NCCL only carries two 4-byte fences
for kernel, gran, odt in (("int8", qm.GRAN_BLOCK, torch.float32), ("f16", qm.GRAN_HEAD, torch.float32),
                          ("int8", qm.GRAN_BLOCK, torch.float16)):
    rep_out = ReplicatedOutput(B, N, H, d, dtype=odt, device=dev)
    rep_out.local.fill_(float("nan"))
    torch.cuda.synchronize()
    for it in range(2):                                    # the second pass reuses the replicas behind the fence
        out = forward_fused_gather(q, k, v, H, rep_out, kernel=kernel, gran=gran)
    full = qm.forward(q, k, v, H, kernel=kernel, gran=gran, out_dtype=odt)
    torch.cuda.synchronize()
    qm.binding.check_async_error()
    same = bool(torch.equal(out, full))
    ok = ok and same
    if rank == 0:
        print(f"forward_fused_gather({kernel}, gran={gran}, {odt}) == single-GPU forward: {same}", flush=True)
    dist.barrier()
    rep_out.close()
# timing at a BASELINE-sized shape: 8 batch entries of C4 split over the ranks, result replicated on every rank
Bt, Ht, Nt, dt = 8, 32, 8192, 128
gen = torch.Generator(device=dev).manual_seed(99)
qt, kt, vt = (torch.rand((Bt, Nt, Ht * dt), device=dev, generator=gen) for _ in range(3))
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
def compute_only():
    for (b, h0, h1) in shard_slabs(Bt, Ht, world, rank):
        sl = [slab_view(t, b, h0, h1, Ht).contiguous() for t in (qt, kt, vt)]
        qm.forward(*sl, h1 - h0, kernel="int8", gran=qm.GRAN_BLOCK)
t_comp = timed(compute_only)
t_1 = timed(lambda: forward_sharded(qt, kt, vt, Ht, kernel="int8", gran=qm.GRAN_BLOCK, chunks=1))
t_4 = timed(lambda: forward_sharded(qt, kt, vt, Ht, kernel="int8", gran=qm.GRAN_BLOCK, chunks=4))
t_8 = timed(lambda: forward_sharded(qt, kt, vt, Ht, kernel="int8", gran=qm.GRAN_BLOCK, chunks=8))
rep_t = ReplicatedOutput(Bt, Nt, Ht, dt, dtype=torch.float32, device=dev)
t_f = timed(lambda: forward_fused_gather(qt, kt, vt, Ht, rep_t, kernel="int8", gran=qm.GRAN_BLOCK))
full_t = qm.forward(qt, kt, vt, Ht, kernel="int8", gran=qm.GRAN_BLOCK)
torch.cuda.synchronize()
same = bool(torch.equal(rep_t.local, full_t))
ok = ok and same
rep_16 = ReplicatedOutput(Bt, Nt, Ht, dt, dtype=torch.float16, device=dev)
t_f16 = timed(lambda: forward_fused_gather(qt, kt, vt, Ht, rep_16, kernel="int8", gran=qm.GRAN_BLOCK))
if rank == 0:
    print(f"C4 split over {world} ranks, FUSED gather (epilogue stores into every replica over NVLink, no all-gather): "
          f"{t_f:.2f} ms with the fp32 result replicated (== single-GPU forward: {same}), {t_f16:.2f} ms with an fp16 result",
          flush=True)
if rank == 0:
    gb = Bt * Nt * Ht * dt * 4 / 1e9
    print(f"C4 split over {world} ranks, replicated {gb:.2f} GB result: own slabs only {t_comp:.2f} ms; "
          f"compute + gather, 1 chunk (no overlap) {t_1:.2f} ms; 4 chunks {t_4:.2f} ms; 8 chunks {t_8:.2f} ms", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("NCCL gather check", "OK" if int(flag.item()) == 1 else "FAILED", f"({world} ranks)", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
