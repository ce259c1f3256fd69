"""Per-warp clock64 timeline of one CTA of the traced INT8 d=128 kernel (qmha_debug_attention_trace).
usage: python tools/trace_dump.py [B,H,N,d] [first_step] [n_steps]      (QMHA_DEBUG_NO_MMA etc. from the env)
Stamps per softmax warp and half-step i: t0 step start, t1 before / t2 after the wait for S(i+1), t3 P(i) published
(arrive on p_full).  MMA warp of tile 0 (row 8): m0 P.V(i) about to issue (v_full, p_full satisfied), m1 P.V(i)
committed, m3 Q.K^T(i+3) issued + committed.  Development aid."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantizedmha_b200 as qm
B, H, N, d = (int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "8,32,8192,128".split(",")))
first = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cnt = int(sys.argv[3]) if len(sys.argv) > 3 else 6
dm = H * d
torch.manual_seed(1)
tq, tk, tv = (torch.rand((B, N, dm), device="cuda") for _ in range(3))
out = torch.empty_like(tq)
Qp, Kp, Vt, sc = qm.quantize_qkv(tq, tk, tv, H, qm.GRAN_BLOCK)
L = qm.lib()
nt = (N + 63) // 64
buf = np.zeros(9 * nt * 4 + 16, np.int64)
for rep in range(2):
    rc = L.qmha_debug_attention_trace(C.c_void_p(Qp.data_ptr()), C.c_void_p(Kp.data_ptr()), C.c_void_p(Vt.data_ptr()),
                                      C.c_void_p(sc.data_ptr()), C.c_void_p(out.data_ptr()), B, N, dm, H, 0, buf.ctypes.data_as(C.c_void_p))
    assert rc == 0, L.qmha_last_error()
tr = buf[:9 * nt * 4].reshape(9, nt, 4)
ph = buf[9 * nt * 4:]
print("CTA phases: setup %d, first scores %d, main loop %d, o_final wait %d, stores %d; total %d" % (
    ph[1] - ph[0], ph[2] - ph[1], ph[3] - ph[2], ph[4] - ph[3], ph[5] - ph[4], ph[6] - ph[0]))
mid = slice(nt // 4, 3 * nt // 4)
print("median step (clk): " + " ".join(f"w{w}:{np.median(np.diff(tr[w, mid, 0])):.0f}" for w in range(8)) + f"  mma0:{np.median(np.diff(tr[8, mid, 0])):.0f}")
t00 = tr[0, first, 0]
for i in range(first, first + cnt):
    print(f"--- half-step {i}   (clocks relative to warp 0's start of step {first})")
    for w in range(4):
        a = tr[w, i] - t00
        print(f"  tile0 warp{w}: start {a[0]:6d}  S({i+1}) wait {a[1]:6d} -> {a[2]:6d} ({a[2]-a[1]:4d})  P({i}) published {a[3]:6d}")
    m = tr[8, i] - t00
    print(f"  mma tile0  : P.V({i}) issue {m[0]:6d}  committed {m[1]:6d}  K tile there {m[2]:6d}  Q.K({i+3}) issued {m[3]:6d}   [p_full({i}) last arrival {max(tr[w, i, 3] for w in range(4)) - t00:6d}]")
    for w in range(4, 8):
        a = tr[w, i] - t00
        print(f"  tile1 warp{w}: start {a[0]:6d}  S({i+1}) wait {a[1]:6d} -> {a[2]:6d} ({a[2]-a[1]:4d})  P({i}) published {a[3]:6d}")
