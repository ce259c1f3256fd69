import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import quantizedmha_b200 as qm
from oracle import load_oracle
orc = load_oracle()
np.set_printoptions(linewidth=220, precision=4, suppress=True)
N, dm, h = 1024, 128, 1
q, k, v = (a[None].copy() for a in orc.golden_inputs(N, dm, h))
v[:] = 0
for n in range(N):
    v[0, n, (n // 64) % 128] = 1.0
tq, tk, tv = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (q, k, v))
ref = orc.mha(q, k, v, h, "f64")[0]
shown = 0
for r in range(40):
    out = qm.forward(tq, tk, tv, h, kernel="int8", gran=qm.GRAN_HEAD)
    torch.cuda.synchronize(); qm.binding.check_async_error()
    o = out.cpu().numpy()[0]
    d = np.abs(o - ref)
    bad = np.where(d.max(axis=1) > 5e-3)[0]
    if len(bad) and shown < 3:
        shown += 1
        print("run", r, "bad rows", len(bad), "warps", sorted({int(b) // 32 for b in bad}))
        for rr in bad[:3]:
            print(" row", rr, "\n  got", o[rr, :16], "\n  ref", ref[rr, :16], "\n  sum got", o[rr].sum())
        colbad = (d[bad] > 5e-3).sum(axis=0)[:16]
        print(" per-half-step count of bad rows:", colbad)
print("done")
