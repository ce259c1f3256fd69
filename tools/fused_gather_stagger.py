"""Fused gather (qmha_args.peer_O) with and without the first-wave stagger, under torchrun:
C4 (B = 8 in total) split over the ranks, fp32 result replicated.  QMHA_PEER_STAGGER_NS is read per launch."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantizedmha_b200 as qm
from quantizedmha_b200.sharding import ReplicatedOutput, unit_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, H, N, d = 8, 32, 8192, 128
dm = H * d
lo, hi = unit_range(B, world, rank)
gen = torch.Generator(device=dev).manual_seed(7700 + rank)
q, k, v = (torch.rand((hi - lo, N, dm), device=dev, generator=gen) for _ in range(3))
rep = ReplicatedOutput(B, N, H, d, dtype=torch.float32, device=dev)
peers = [rep.peer_base[r] + lo * N * dm * 4 for r in sorted(rep.peer_base)]
own = torch.empty_like(q)

def timed(fn, reps=6):
    fn(); fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

def fused():
    rep.fence()
    qm.forward(q, k, v, H, kernel="int8", gran=qm.GRAN_BLOCK, out=rep.local[lo:hi], peer_outs=peers)
    rep.fence()

t_own = timed(lambda: qm.forward(q, k, v, H, kernel="int8", gran=qm.GRAN_BLOCK, out=own))
res = {}
for ns in sys.argv[1:] or ["0", "auto", "30000", "60000", "90000", "120000"]:
    if ns == "auto": os.environ.pop("QMHA_PEER_STAGGER_NS", None)
    else: os.environ["QMHA_PEER_STAGGER_NS"] = ns
    res[ns] = timed(fused)
os.environ.pop("QMHA_PEER_STAGGER_NS", None)
gathered = torch.empty((B, N, dm), device=dev)
dist.all_gather_into_tensor(gathered, own)
torch.cuda.synchronize(); qm.binding.check_async_error()
same = bool(torch.equal(rep.local, gathered))
flag = torch.tensor([1 if same else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"{world} GPUs: own slabs {t_own:.3f} ms; fused gather by first-wave stagger (ns): " +
          ", ".join(f"{k_}: {v_:.3f} ms" for k_, v_ in res.items()) + f"; replicas == NCCL result: {bool(int(flag.item()))}", flush=True)
dist.barrier(); rep.close(); dist.destroy_process_group()
