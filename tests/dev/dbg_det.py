import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import quantizedmha_b200 as qm
from oracle import load_oracle
orc = load_oracle()
for (B, N, dm, h) in [(1, 1024, 256, 2), (1, 2048, 512, 4), (4, 4096, 1024, 8)]:
    q, k, v = (np.stack([a] * B) for a in orc.golden_inputs(N, dm, h))
    tq, tk, tv = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (q, k, v))
    for kern, gran in (("int8", qm.GRAN_BLOCK), ("int8", qm.GRAN_HEAD), ("f16", qm.GRAN_HEAD)):
        outs = []
        for r in range(6):
            out = qm.forward(tq, tk, tv, h, kernel=kern, gran=gran)
            torch.cuda.synchronize(); qm.binding.check_async_error()
            outs.append(out.clone())
        diffs = [float((outs[0] - o).abs().max()) for o in outs]
        nbad = [int(((outs[0] - o).abs().amax(dim=-1) > 1e-4).sum()) for o in outs]
        print(B, N, dm, h, kern, gran, "maxdiff vs run0", diffs, "rows differing", nbad, flush=True)
