import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import quantizedmha_b200 as qm
from oracle import load_oracle
orc = load_oracle()
mode = sys.argv[1]  # "ref" (old lib: save) or "test"
kern = sys.argv[2] if len(sys.argv) > 2 else "int8"
OUT = os.path.join(ROOT, "gpurun_out", "tmp"); os.makedirs(OUT, exist_ok=True)
for N in (512, 640, 768, 1024, 2048):
    dm, h = 128, 1
    q, k, v = (a[None] for a in orc.golden_inputs(N, dm, h))
    tq, tk, tv = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (q, k, v))
    f = os.path.join(OUT, f"ref_{kern}_{N}.npy")
    if mode == "ref":
        out = qm.forward(tq, tk, tv, h, kernel=kern, gran=qm.GRAN_HEAD)
        torch.cuda.synchronize(); np.save(f, out.cpu().numpy()); continue
    ref = torch.from_numpy(np.load(f)).cuda()
    nbad = 0; rows = set(); maxd = 0.0
    for r in range(30):
        out = qm.forward(tq, tk, tv, h, kernel=kern, gran=qm.GRAN_HEAD)
        torch.cuda.synchronize(); qm.binding.check_async_error()
        d = (out - ref).abs().amax(dim=-1)[0]
        bad = torch.nonzero(d > 1e-5).flatten().tolist()
        if bad: nbad += 1; rows.update(bad); maxd = max(maxd, float(d.max()))
    rows = sorted(rows)
    print(N, "runs with mismatch:", nbad, "/30 maxdiff", maxd, "rows", len(rows), rows[:12], "warps", sorted({r // 32 for r in rows})[:40], flush=True)
