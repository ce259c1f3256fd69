"""Times the INT8 quantise pass of several library builds (QMHA per-head / per-block / per-tensor scales) on one box.
usage: python tools/quant_ab.py name=path [...] [--gran head|block|tensor] [--shape B,H,N,d] [--reps R]
Every build is loaded side by side; codes and scales of each are compared with the first one (must be identical)."""
import argparse, ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantizedmha_b200 import binding as qb
ap = argparse.ArgumentParser()
ap.add_argument("libs", nargs="+")
ap.add_argument("--gran", default="head")
ap.add_argument("--shape", default="8,32,8192,128")
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
B, H, N, d = (int(x) for x in a.shape.split(","))
dm = H * d
gran = {"tensor": 0, "head": 1, "block": 2}[a.gran]
torch.manual_seed(0)
tq, tk, tv = (torch.rand((B, N, dm), device="cuda") for _ in range(3))
n_pad = (N + 255) // 256 * 256
u = B * H
ref = None
for spec in a.libs:
    name, path = spec.split("=", 1)
    L = qb.declare(C.CDLL(os.path.abspath(path)))
    Qp = torch.empty((u, n_pad, d), dtype=torch.int8, device="cuda"); Kp = torch.empty_like(Qp)
    Vt = torch.empty((u, d, n_pad), dtype=torch.float16, device="cuda")
    sc = torch.empty((3, u, n_pad // 32) if gran == 2 else (3, u), device="cuda")
    def run():
        rc = L.qmha_quantize_qkv(tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), B, N, dm, H, gran, Qp.data_ptr(), Kp.data_ptr(), Vt.data_ptr(), sc.data_ptr(), None)
        assert rc == 0, L.qmha_last_error()
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    E = B * N * dm
    same = "first"
    if ref is None: ref = (Qp.clone(), Kp.clone(), Vt.clone(), sc.clone())
    else: same = "identical" if all(torch.equal(x, y) for x, y in zip(ref, (Qp, Kp, Vt, sc))) else "DIFFERENT"
    print(f"{name:12s} {ms:7.3f} ms  {15 * E / ms / 1e6:7.0f} GB/s algorithmic (15E)  outputs {same}", flush=True)
