"""Same-box A/B of several builds of libqmha.so (attention kernel only, prepared operands shared).

usage: python tools/ab_libs.py name=path [name=path ...] [--shape B,H,N,d] [--rounds R] [--reps K]
                               [--kernel int8|f16] [--gran block|head] [--data uniform|normal]

All libraries are loaded side by side in ONE process and timed in alternation (R rounds), so box-to-box
and warm-up differences cancel.  Every variant's output is compared with the first library's FP16-kernel
output (rel-L2) as a sanity check — the parity tests proper are tests/test_gpu_parity.py.
Writes gpurun_out/ab_libs.json.  Development aid, not part of the tests."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantizedmha_b200 import binding as qb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("libs", nargs="+")
ap.add_argument("--shape", default="8,32,8192,128")
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--kernel", default="int8")
ap.add_argument("--gran", default="block")
ap.add_argument("--data", default="uniform")
ap.add_argument("--out", default="ab_libs.json")
ap.add_argument("--nomma", action="store_true", help="QMHA_DEBUG_NO_MMA: softmax side alone")
args = ap.parse_args()

os.environ["QMHA_CYCLES"] = "1"
if args.nomma:
    os.environ["QMHA_DEBUG_NO_MMA"] = "1"
B, H, N, d = (int(x) for x in args.shape.split(","))
dm = H * d
dev = torch.device("cuda:0")
libs = []
for spec in args.libs:
    name, path = spec.split("=", 1)
    libs.append((name, qb.declare(C.CDLL(os.path.abspath(path)))))

samples = []
_p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                       "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"], stdout=subprocess.PIPE, text=True)


def _rd():
    for ln in _p.stdout:
        f = [x.strip() for x in ln.split(",")]
        try:
            samples.append((time.time(), float(f[0]), float(f[1]), f[2]))
        except Exception:  # noqa: BLE001
            pass


threading.Thread(target=_rd, daemon=True).start()


def clocks_between(t0, t1):
    s = [x for x in samples if t0 <= x[0] <= t1]
    if not s:
        return {}
    mhz = sorted(x[1] for x in s)
    return {"sm_mhz": mhz[len(mhz) // 2], "power_w": max(x[2] for x in s),
            "power_cap": any(x[3].lower().startswith("active") for x in s)}


def chk(L, rc):
    if rc != 0:
        raise RuntimeError(L.qmha_last_error().decode())


torch.manual_seed(1)
if args.data == "uniform":
    tq, tk, tv = (torch.rand((B, N, dm), device=dev) for _ in range(3))
else:
    tq, tk, tv = (torch.randn((B, N, dm), device=dev) for _ in range(3))
n_pad = (N + 255) // 256 * 256
d_pad = 32 if d <= 32 else (64 if d <= 64 else 128)
units = B * H
gran = {"block": qb.GRAN_BLOCK, "head": qb.GRAN_HEAD}[args.gran]
L0 = libs[0][1]
out = torch.empty_like(tq)
# FP16 anchor output from the first library
Qh = torch.empty((units, n_pad, d_pad), dtype=torch.float16, device=dev)
Kh = torch.empty_like(Qh)
Vt = torch.empty((units, d_pad, n_pad), dtype=torch.float16, device=dev)
chk(L0, L0.qmha_convert_qkv_f16(tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), B, N, dm, H, Qh.data_ptr(), Kh.data_ptr(), Vt.data_ptr(), None))
anchor = torch.empty_like(tq)
chk(L0, L0.qmha_attention_prepared(Qh.data_ptr(), Kh.data_ptr(), Vt.data_ptr(), None, anchor.data_ptr(), B, N, dm, H, qb.KERNEL_F16, qb.GRAN_HEAD, None))
torch.cuda.synchronize()
if args.kernel == "int8":
    del Qh, Kh
    Qp = torch.empty((units, n_pad, d_pad), dtype=torch.int8, device=dev)
    Kp = torch.empty_like(Qp)
    sc = torch.empty((3, units, n_pad // 32) if gran == qb.GRAN_BLOCK else (3, units), dtype=torch.float32, device=dev)
    chk(L0, L0.qmha_quantize_qkv(tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), B, N, dm, H, gran, Qp.data_ptr(), Kp.data_ptr(), Vt.data_ptr(), sc.data_ptr(), None))
    kid, scp = qb.KERNEL_INT8, sc.data_ptr()
else:
    Qp, Kp, kid, scp, gran = Qh, Kh, qb.KERNEL_F16, None, qb.GRAN_HEAD
torch.cuda.synchronize()
del tq, tk, tv


def launch(L):
    chk(L, L.qmha_attention_prepared(Qp.data_ptr(), Kp.data_ptr(), Vt.data_ptr(), scp, out.data_ptr(), B, N, dm, H, kid, gran, None))


res = {name: [] for name, _ in libs}
for rnd in range(args.rounds):
    for name, L in libs:
        try:
            for _ in range(2):
                launch(L)
            torch.cuda.synchronize()
            chk(L, L.qmha_check_async_error())
            cyc = (C.c_ulonglong * 2)()
            has_cyc = hasattr(L, "qmha_debug_cycles") and L.qmha_debug_cycles(cyc, 1) == 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0 = time.time()
            e0.record()
            for _ in range(args.reps):
                launch(L)
            e1.record()
            torch.cuda.synchronize()
            w1 = time.time()
            ms = e0.elapsed_time(e1) / args.reps
            time.sleep(0.06)
            rec = {"ms": ms, "tflops": 4.0 * B * H * N * N * d / ms / 1e9,
                   "rel_l2_vs_f16": float(((out - anchor).norm() / anchor.norm()).item()),
                   "max_abs_vs_f16": float((out - anchor).abs().max().item())}
            if has_cyc and L.qmha_debug_cycles(cyc, 1) == 0 and cyc[1]:
                rec["cta_kclk"] = cyc[0] / cyc[1] / 1e3     # mean residency of a CTA, exact SM clocks
            if has_cyc and hasattr(L, "qmha_debug_sm_spans"):   # one launch: per-SM span against the sum of CTA residencies
                L.qmha_debug_sm_spans.argtypes = [C.POINTER(C.c_ulonglong), C.c_int, C.c_int]
                sp = (C.c_ulonglong * 384)()
                L.qmha_debug_cycles(cyc, 1)
                launch(L)
                torch.cuda.synchronize()
                if L.qmha_debug_cycles(cyc, 0) == 0 and L.qmha_debug_sm_spans(sp, 192, 1) == 0:
                    spans = [sp[2 * i] for i in range(192) if sp[2 * i]]
                    rec["one_launch"] = {"sms": len(spans), "ctas": cyc[1], "resid_sum_mclk": cyc[0] / 1e6,
                                         "span_sum_mclk": sum(spans) / 1e6, "span_max_kclk": max(spans) / 1e3,
                                         "span_mean_kclk": sum(spans) / len(spans) / 1e3,
                                         "gap_frac_of_span": 1.0 - cyc[0] / sum(spans)}
            rec.update(clocks_between(w0 + 0.03, w1))
            if rec.get("sm_mhz"):
                rec["mclk"] = ms * 1e-3 * rec["sm_mhz"]
        except Exception as e:  # noqa: BLE001
            rec = {"error": str(e)}
        res[name].append(rec)
        print(name, json.dumps(rec), flush=True)
_p.terminate()
print("---- summary (best ms / median ms / Mclk at median clock)")
for name, _ in libs:
    ok = [r for r in res[name] if "ms" in r]
    if not ok:
        print(f"{name:24s} failed: {res[name][0].get('error')}")
        continue
    ms = sorted(r["ms"] for r in ok)
    mclk = sorted(r.get("mclk", 0.0) for r in ok)
    kc = sorted(r.get("cta_kclk", 0.0) for r in ok)
    print(f"{name:24s} best {ms[0]:.3f} ms  median {ms[len(ms) // 2]:.3f} ms  {mclk[len(mclk) // 2]:.2f} Mclk  CTA {kc[len(kc) // 2]:.1f} kclk  "
          f"rel-L2 vs f16 {ok[0]['rel_l2_vs_f16']:.2e}  max-abs {ok[0]['max_abs_vs_f16']:.2e}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", args.out), "w"), indent=1)
