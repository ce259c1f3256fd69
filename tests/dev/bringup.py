"""GPU bring-up / diagnostics: runs a matrix of small cases through the C-ABI and REPORTS error
statistics against the CPU oracle instead of asserting, so that one gpurun call yields as much
information as possible.  Writes gpurun_out/bringup.json (+ .npy dumps for failing cases).
Not part of the product or the test-suite; tests/ holds the gated versions of these checks."""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import quantizedmha_b200 as qm  # noqa: E402
from oracle import load_oracle  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
orc = load_oracle()
report = {"cases": []}
dev = torch.device("cuda:0")


def stats(got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    fin = bool(np.isfinite(got).all())
    diff = np.abs(got - ref)
    return {"finite": fin, "max_abs": float(np.nanmax(diff)), "rel_l2": float(np.linalg.norm(np.nan_to_num(got - ref)) / max(np.linalg.norm(ref), 1e-30)),
            "ref_absmax": float(np.abs(ref).max())}


def unpack_q(Qp, B, N, h, d):
    # [B*h, n_pad, d_pad] -> [B, N, h*d]
    a = Qp.cpu().numpy().reshape(B, h, Qp.shape[1], Qp.shape[2])[:, :, :N, :d]
    return np.ascontiguousarray(a.transpose(0, 2, 1, 3)).reshape(B, N, h * d)


def unpack_vt(Vt, B, N, h, d):
    # [B*h, d_pad, n_pad] -> [B, N, h*d]
    a = Vt.float().cpu().numpy().reshape(B, h, Vt.shape[1], Vt.shape[2])[:, :, :d, :N]
    return np.ascontiguousarray(a.transpose(0, 3, 1, 2)).reshape(B, N, h * d)


def run_case(name, q, k, v, h, kernels=("int8", "f16"), dump=False):
    q, k, v = (np.ascontiguousarray(a, np.float32) for a in (q, k, v))
    B = 1 if q.ndim == 2 else q.shape[0]
    N, dm = q.shape[-2], q.shape[-1]
    d = dm // h
    ref = orc.mha(q, k, v, h, "f64").reshape(B, N, dm)
    tq, tk, tv = (torch.from_numpy(a.reshape(B, N, dm)).to(dev) for a in (q, k, v))
    for kern in kernels:
        rec = {"name": name, "kernel": kern, "B": B, "N": N, "d_model": dm, "h": h}
        try:
            if kern == "int8":
                Qp, Kp, Vt, sc = qm.quantize_qkv(tq, tk, tv, h)
                torch.cuda.synchronize()
                qq, sq = orc.quantize(q.reshape(B, N, dm), h, "head")
                kq, sk = orc.quantize(k.reshape(B, N, dm), h, "head")
                vq, sv = orc.quantize(v.reshape(B, N, dm), h, "head")
                scn = sc.cpu().numpy()
                rec["quant_bit_exact"] = {
                    "Q": bool(np.array_equal(unpack_q(Qp, B, N, h, d), qq)),
                    "K": bool(np.array_equal(unpack_q(Kp, B, N, h, d), kq)),
                    "V": bool(np.array_equal(unpack_vt(Vt, B, N, h, d), vq.astype(np.float32))),
                    "scales": bool(np.array_equal(scn, np.stack([sq, sk, sv]))),
                    "pad_zero": bool((Qp[:, N:, :] == 0).all().item() and (Qp[:, :, d:] == 0).all().item()
                                     and (Vt[:, d:, :] == 0).all().item() and (Vt[:, :, N:] == 0).all().item()),
                }
                emu = orc.mha_int8_emulated(qq, kq, vq, sq, sk, sv, h, "f16").reshape(B, N, dm)
                rec["emu_vs_fp"] = stats(emu, ref)
            out = qm.forward(tq, tk, tv, h, kernel=kern)
            torch.cuda.synchronize()
            qm.binding.check_async_error()
            o = out.cpu().numpy()
            rec["vs_fp"] = stats(o, ref)
            if kern == "int8":
                rec["vs_emu"] = stats(o, emu)
            bad = (not rec["vs_fp"]["finite"]) or rec["vs_fp"]["max_abs"] > (2e-2 if kern == "int8" else 2e-3)
            rec["ok"] = not bad
            if dump or bad:
                np.save(os.path.join(OUT, f"dump_{name}_{kern}_got.npy"), o[0][:256, :256])
                np.save(os.path.join(OUT, f"dump_{name}_{kern}_ref.npy"), ref[0][:256, :256].astype(np.float32))
        except Exception as e:  # noqa: BLE001
            rec["error"] = f"{type(e).__name__}: {e}"
            rec["ok"] = False
            traceback.print_exc()
        print(json.dumps(rec), flush=True)
        report["cases"].append(rec)


def main():
    print("device:", torch.cuda.get_device_name(0), "| lib:", qm.lib().qmha_version().decode(), flush=True)
    rng = np.random.default_rng(0)
    # --- diagnostics first: structure-revealing inputs, N=128/256, d=128, one head
    n, d = 128, 128
    eye = np.eye(n, d, dtype=np.float32)
    q0 = np.zeros((n, d), np.float32)
    g = (rng.standard_normal((n, d)) * 0.5).astype(np.float32)
    run_case("ones_128", np.ones((n, d), np.float32), np.ones((n, d), np.float32), np.ones((n, d), np.float32), 1)
    run_case("uniformP_randV_128", q0, g, (rng.standard_normal((n, d))).astype(np.float32), 1, dump=True)
    run_case("randQK_eyeV_128", g, (rng.standard_normal((n, d)) * 0.5).astype(np.float32), eye, 1, dump=True)
    run_case("rand_128", g, (rng.standard_normal((n, d)) * 0.5).astype(np.float32), (rng.standard_normal((n, d))).astype(np.float32), 1)
    for (N, dm, h) in [(256, 128, 1), (512, 128, 1), (384, 256, 2), (1024, 128, 2), (512, 64, 1), (512, 128, 4), (300, 128, 1), (50, 64, 8), (8, 32, 4)]:
        q, k, v = orc.golden_inputs(N, dm, h)
        run_case(f"golden_N{N}_dm{dm}_h{h}", q, k, v, h)
    q, k, v = orc.profile_inputs(2 * 640, 256)
    run_case("profile_B2_N640_dm256_h2", q.reshape(2, 640, 256), k.reshape(2, 640, 256), v.reshape(2, 640, 256), 2)
    q, k, v = orc.profile_inputs(2048, 512)
    run_case("profile_N2048_dm512_h4", q, k, v, 4)
    q, k, v = orc.profile_inputs(2048, 128)
    run_case("profile_N2048_dm128_h4_d32", q, k, v, 4)

    # --- timing: headline shape (C4) and a smaller one; inputs U[0,1) generated on device
    for (B, H, N, d, tag) in [(1, 8, 4096, 64, "C3"), (2, 32, 8192, 128, "C4_quarter"), (8, 32, 8192, 128, "C4")]:
        try:
            dm = H * d
            torch.manual_seed(1)
            tq = torch.rand((B, N, dm), device=dev)
            tk = torch.rand((B, N, dm), device=dev)
            tv = torch.rand((B, N, dm), device=dev)
            out = torch.empty_like(tq)
            for kern in ("int8", "f16"):
                rec = {"name": f"time_{tag}", "kernel": kern, "B": B, "H": H, "N": N, "d": d}
                if kern == "int8":
                    Qp, Kp, Vt, sc = qm.quantize_qkv(tq, tk, tv, H)
                else:
                    Qp, Kp, Vt = qm.convert_qkv_f16(tq, tk, tv, H)
                    sc = None
                torch.cuda.synchronize()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                for _ in range(2):
                    qm.attention_prepared(Qp, Kp, Vt, sc, B, N, dm, H, kern, out=out)
                torch.cuda.synchronize()
                qm.binding.check_async_error()
                reps = 5
                ev[0].record()
                for _ in range(reps):
                    qm.attention_prepared(Qp, Kp, Vt, sc, B, N, dm, H, kern, out=out)
                ev[1].record()
                torch.cuda.synchronize()
                ms = ev[0].elapsed_time(ev[1]) / reps
                flops = 4.0 * B * H * N * N * d
                rec["attn_ms"] = ms
                rec["attn_tflops"] = flops / ms / 1e9
                ev[2].record()
                for _ in range(reps):
                    if kern == "int8":
                        qm.quantize_qkv(tq, tk, tv, H)
                    else:
                        qm.convert_qkv_f16(tq, tk, tv, H)
                ev[3].record()
                torch.cuda.synchronize()
                pms = ev[2].elapsed_time(ev[3]) / reps
                E = B * N * dm
                rec["prep_ms"] = pms
                rec["prep_alg_GBs"] = (3 * E * 4 + 3 * E * (1 if kern == "int8" else 2) + (E if kern == "int8" else 0)) / pms / 1e6
                print(json.dumps(rec), flush=True)
                report["cases"].append(rec)
                del Qp, Kp, Vt
            del tq, tk, tv, out
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            traceback.print_exc()
            report["cases"].append({"name": f"time_{tag}", "error": str(e)})
    json.dump(report, open(os.path.join(OUT, "bringup.json"), "w"), indent=1)
    nbad = sum(1 for c in report["cases"] if c.get("ok") is False)
    print("BRINGUP DONE; failing cases:", nbad)


if __name__ == "__main__":
    main()
