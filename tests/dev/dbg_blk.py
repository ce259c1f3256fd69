import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import quantizedmha_b200 as qm
from oracle import load_oracle
orc = load_oracle()
for (B, N, dm, h) in [(1, 512, 128, 1), (1, 1024, 128, 1), (1, 1024, 256, 2)]:
    q, k, v = (np.stack([a] * B) for a in orc.golden_inputs(N, dm, h))
    ref = orc.mha(q, k, v, h, "f64")
    tq, tk, tv = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (q, k, v))
    out = qm.forward(tq, tk, tv, h, kernel="int8", gran=qm.GRAN_BLOCK)
    torch.cuda.synchronize(); qm.binding.check_async_error()
    o = out.cpu().numpy()
    outh = qm.forward(tq, tk, tv, h, kernel="int8", gran=qm.GRAN_HEAD).cpu().numpy()
    err = np.abs(o - ref)[0]          # [N, dm]
    rowerr = err.max(axis=1)
    print(B, N, dm, h, "blk max", err.max(), "head max", np.abs(outh-ref).max())
    bad = np.where(rowerr > 5e-3)[0]
    print(" bad rows:", len(bad), bad[:40], bad[-10:] if len(bad) else "")
    if len(bad):
        r = bad[0]
        print(" row", r, "got", o[0, r, :6], "ref", ref[0, r, :6], "ratio", (o[0, r, :6] / ref[0, r, :6]))
        colerr = err.max(axis=0); print(" bad cols:", np.where(colerr > 5e-3)[0][:20])
