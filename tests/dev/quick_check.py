"""Tiny correctness probe for a kernel variant (QMHA_ATTN_VARIANT) with small shapes first."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import quantizedmha_b200 as qm
from oracle import load_oracle
orc = load_oracle()
for (N, dm, h) in [(128, 128, 1), (256, 128, 1), (300, 128, 1), (64, 128, 1), (1024, 256, 2), (512, 64, 1), (50, 64, 8), (2048, 512, 4)]:
    q, k, v = orc.golden_inputs(N, dm, h)
    ref = orc.mha(q, k, v, h, "f64")
    tq, tk, tv = (torch.from_numpy(a).cuda() for a in (q, k, v))
    for kern, tol in (("int8", 2e-2), ("f16", 2e-3)):
        out = qm.forward(tq, tk, tv, h, kernel=kern)
        torch.cuda.synchronize()
        qm.binding.check_async_error()
        err = float(np.abs(out.cpu().numpy() - ref).max())
        print(N, dm, h, kern, f"{err:.3e}", "OK" if err <= tol else "FAIL", flush=True)
