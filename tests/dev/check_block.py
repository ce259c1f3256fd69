"""Block-scale mode probe: quantiser bit-exactness vs the oracle, attention vs emulated/FP, timing."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import quantizedmha_b200 as qm
from oracle import load_oracle
orc = load_oracle()

def unpack_rows(Qp, B, N, h, d):
    a = Qp.cpu().numpy().reshape(B, h, Qp.shape[1], Qp.shape[2])[:, :, :N, :d]
    return np.ascontiguousarray(a.transpose(0, 2, 1, 3)).reshape(B, N, h * d)
def unpack_vt(Vt, B, N, h, d):
    a = Vt.float().cpu().numpy().reshape(B, h, Vt.shape[1], Vt.shape[2])[:, :, :d, :N]
    return np.ascontiguousarray(a.transpose(0, 3, 1, 2)).reshape(B, N, h * d)

for (B, N, dm, h, kind) in [(1, 128, 128, 1, "golden"), (1, 256, 128, 1, "golden"), (2, 300, 256, 2, "golden"), (1, 1024, 256, 2, "golden"),
                            (1, 512, 64, 1, "golden"), (1, 50, 64, 8, "golden"), (1, 2048, 512, 4, "profile"), (1, 1000, 128, 4, "peaked")]:
    d = dm // h
    if kind == "profile":
        q, k, v = (a.reshape(B, N, dm) for a in orc.profile_inputs(B * N, dm))
    else:
        q, k, v = (np.stack([a] * B) for a in orc.golden_inputs(N, dm, h))
        if kind == "peaked":
            q, k, v = q * 4, k * 4, v * 4
        if B > 1:
            q[1] *= 1.5
    ref = orc.mha(q, k, v, h, "f64")
    tq, tk, tv = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (q, k, v))
    Qp, Kp, Vt, sc = qm.quantize_qkv(tq, tk, tv, h, qm.GRAN_BLOCK)
    torch.cuda.synchronize()
    nb = -(-N // 32)
    exact = {}
    emu_in = []
    for i, (x, packed, un) in enumerate(((q, Qp, unpack_rows), (k, Kp, unpack_rows), (v, Vt, unpack_vt))):
        codes, s = orc.quantize(x, h, "block", 32)
        got_s = sc[i].cpu().numpy()[:, :nb].reshape(-1)
        exact["QKV"[i]] = bool(np.array_equal(un(packed, B, N, h, d), codes.astype(np.float32) if i == 2 else codes)) and bool(np.array_equal(got_s, s))
        emu_in.append((codes, s))
    emu = orc.mha_int8_emulated_block(emu_in[0][0], emu_in[1][0], emu_in[2][0], emu_in[0][1], emu_in[1][1], emu_in[2][1], h, 32, "f16")
    out = qm.forward(tq, tk, tv, h, kernel="int8", gran=qm.GRAN_BLOCK)
    torch.cuda.synchronize(); qm.binding.check_async_error()
    o = out.cpu().numpy()
    outh = qm.forward(tq, tk, tv, h, kernel="int8", gran=qm.GRAN_HEAD).cpu().numpy()
    e = lambda a, b: (float(np.abs(a - b).max()), float(np.linalg.norm(a - b) / np.linalg.norm(b)))
    print(json.dumps({"shape": [B, N, dm, h], "kind": kind, "quant_exact": exact, "blk_vs_fp": e(o, ref), "blk_vs_emu": e(o, emu),
                      "head_vs_fp": e(outh, ref)}), flush=True)
