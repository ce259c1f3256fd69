"""Runs the quantise + attention kernels a few times on the C4 shape (for ncu captures).
usage: prof_one.py [int8|f16|bf16] [B,H,N,d] [reps] [block|head]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantizedmha_b200 as qm
kern = sys.argv[1] if len(sys.argv) > 1 else "int8"
B, H, N, d = (int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "8,32,8192,128".split(",")))
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
gran = qm.GRAN_BLOCK if (len(sys.argv) > 4 and sys.argv[4] == "block") else qm.GRAN_HEAD
dev = torch.device("cuda:0")
torch.manual_seed(1)
dm = H * d
tq, tk, tv = (torch.rand((B, N, dm), device=dev) for _ in range(3))
out = torch.empty_like(tq)
if kern == "int8":
    Qp, Kp, Vt, sc = qm.quantize_qkv(tq, tk, tv, H, gran)
else:
    Qp, Kp, Vt = qm.convert_qkv_f16(tq, tk, tv, H, kernel=kern); sc = None
for _ in range(reps):
    qm.attention_prepared(Qp, Kp, Vt, sc, B, N, dm, H, kern, out=out, gran=gran)
torch.cuda.synchronize()
qm.binding.check_async_error()
print("ok", float(out.abs().mean().item()))
