"""A/B timing of attention-kernel variants on the GPU (QMHA_ATTN_VARIANT) + timeline traces.
Writes gpurun_out/tune.json and gpurun_out/trace_v*.npy.  Development aid, not part of the tests."""
import ctypes as C
import json
import os
import sys

import subprocess
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantizedmha_b200 as qm  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
dev = torch.device("cuda:0")
variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,8,4".split(","))]
B, H, N, d = (int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "8,32,8192,128".split(",")))
dm = H * d
torch.manual_seed(1)
tq, tk, tv = (torch.rand((B, N, dm), device=dev) for _ in range(3))
out = torch.empty_like(tq)
res = []
L = qm.lib()

samples = []  # (time, sm_mhz, power_w)
_p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                       "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"], stdout=subprocess.PIPE, text=True)
def _rd():
    for ln in _p.stdout:
        f = [x.strip() for x in ln.split(",")]
        try:
            samples.append((time.time(), float(f[0]), float(f[1]), f[2], f[3], f[4]))
        except Exception:
            pass
threading.Thread(target=_rd, daemon=True).start()

def clocks_between(t0, t1):
    s = [x for x in samples if t0 <= x[0] <= t1]
    if not s:
        return {}
    mhz = sorted(x[1] for x in s)
    return {"sm_mhz_med": mhz[len(mhz) // 2], "sm_mhz_min": mhz[0], "power_w_max": max(x[2] for x in s),
            "power_cap": any(x[3].lower().startswith("active") for x in s), "n": len(s)}

for kern in ("int8", "f16"):
    if kern == "int8":
        GR = qm.GRAN_BLOCK if os.environ.get("TUNE_BLOCK") else qm.GRAN_HEAD
        Qp, Kp, Vt, sc = qm.quantize_qkv(tq, tk, tv, H, GR)
    else:
        Qp, Kp, Vt = qm.convert_qkv_f16(tq, tk, tv, H)
        sc = None
    base = None
    for v in variants:
        os.environ["QMHA_ATTN_VARIANT"] = str(v)
        try:
            for _ in range(2):
                qm.attention_prepared(Qp, Kp, Vt, sc, B, N, dm, H, kern, out=out, gran=(GR if kern == "int8" else qm.GRAN_HEAD))
            torch.cuda.synchronize()
            qm.binding.check_async_error()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 40
            w0 = time.time()
            e0.record()
            for _ in range(reps):
                qm.attention_prepared(Qp, Kp, Vt, sc, B, N, dm, H, kern, out=out, gran=(GR if kern == "int8" else qm.GRAN_HEAD))
            e1.record()
            torch.cuda.synchronize()
            w1 = time.time()
            ms = e0.elapsed_time(e1) / reps
            time.sleep(0.05)
            if base is None:
                base = out.clone()
            rec = {"kernel": kern, "variant": v, "poly_every": v, "ms": ms,
                   "tflops": 4.0 * B * H * N * N * d / ms / 1e9,
                   "max_abs_vs_first": float((out - base).abs().max().item()),
                   "rel_l2_vs_first": float(((out - base).norm() / base.norm()).item())}
            ck = clocks_between(w0 + 0.05, w1)
            rec.update(ck)
            if ck.get("sm_mhz_med"):
                rec["mclk_per_launch"] = ms * 1e-3 * ck["sm_mhz_med"]
        except Exception as e:  # noqa: BLE001
            rec = {"kernel": kern, "variant": v, "error": str(e)}
        print(json.dumps(rec), flush=True)
        res.append(rec)
    if kern == "int8" and d == 128:
        nt = (N + 63) // 64
        for v in (0,):
            buf = np.zeros(9 * nt * 4 + 16, np.int64)
            rc = L.qmha_debug_attention_trace(C.c_void_p(Qp.data_ptr()), C.c_void_p(Kp.data_ptr()), C.c_void_p(Vt.data_ptr()),
                                              C.c_void_p(sc.data_ptr()), C.c_void_p(out.data_ptr()), B, N, dm, H, v,
                                              buf.ctypes.data_as(C.c_void_p))
            if rc != 0:
                print("trace failed:", L.qmha_last_error().decode())
                continue
            tr = buf[:9 * nt * 4].reshape(9, nt, 4)
            ph = buf[9 * nt * 4:]
            np.save(os.path.join(OUT, f"trace_v{v}.npy"), tr)
            print("CTA phases (clk): setup %d, to first scores %d, main loop %d, wait O final %d, stores %d, exit barrier %d; total %d" % (
                ph[1] - ph[0], ph[2] - ph[1], ph[3] - ph[2], ph[4] - ph[3], ph[5] - ph[4], ph[6] - ph[5], ph[6] - ph[0]), flush=True)
            print("  epilogue detail (clk from O final): all tiles handed to the TMA %d, end %d" % tuple(int(ph[k] - ph[4]) for k in (11, 5)), flush=True)
            mid = slice(nt // 4, 3 * nt // 4)
            print(f"trace v{v}: median MMA iteration {np.median(np.diff(tr[8, mid, 0])):.0f} clk", flush=True)
            for w in range(8):
                st = np.diff(tr[w, mid, 0])
                wait = (tr[w, mid, 2] - tr[w, mid, 1])
                print(f"  softmax warp {w}: step median {np.median(st):.0f} mean {st.mean():.0f} clk; s_full wait mean {wait.mean():.0f} max {wait.max()}; "
                      f"start->prefetch {np.median(tr[w, mid, 1] - tr[w, mid, 0]):.0f}", flush=True)
    del Qp, Kp, Vt
os.environ.pop("QMHA_ATTN_VARIANT", None)
json.dump(res, open(os.path.join(OUT, "tune.json"), "w"), indent=1)
_p.terminate()
