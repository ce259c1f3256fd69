"""Timing loop in the spirit of the reference's extensions/torch/run_torch_example.py:43-59.
Usage: python run_torch_example.py [--N 8192 --d_model 1024 --heads 32 --kernel fa_tc_int8_b]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch_ext  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=8192)
ap.add_argument("--d_model", type=int, default=1024)
ap.add_argument("--heads", type=int, default=32)
ap.add_argument("--kernel", default="fa_tc_int8_b")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
torch.manual_seed(0)
Q, K, V = (torch.randn(a.N, a.d_model, device="cuda") for _ in range(3))
for _ in range(3):
    out = torch_ext.flash_solve(Q, K, V, a.d_model, a.heads, a.kernel)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    out = torch_ext.flash_solve(Q, K, V, a.d_model, a.heads, a.kernel)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print(f"{a.kernel}: {ms:.3f} ms / call, {4 * a.N * a.N * a.d_model / ms / 1e9:.1f} TFLOP/s, out {tuple(out.shape)} {out.dtype}")
