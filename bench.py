#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json: attention fwd TFLOP/s & ms at
B8/H32/N8192/d128 INT8).

    python bench.py --gpus N --steps K --warmup W          # this framework (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W  # the reference's CPU path, host cores

One "step" = one pass of the hot path (INT8 quantise kernels + fused attention kernel, i.e. what
qmha_forward()/solve() enqueue) over one batch of synthetic inputs resident in HBM.  Under
torchrun (N>1) every rank owns its own (batch × head) units — no collective on the data path —
and the time is the max over ranks of the CUDA-event time of the K steps.

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for the definition of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B, H, N, d, kernel, scaling)
    "c4": (8, 32, 8192, 128, "int8", "weak"),    # headline roofline point (per GPU)
    "c3": (1, 8, 4096, 64, "int8", "weak"),      # reference fa_tc_int8_b comparison shape
    "c2": (1, 32, 8192, 32, "f16", "weak"),      # reference default config.h shape, FP16 anchor
    "c4f16": (8, 32, 8192, 128, "f16", "weak"),  # FP16 anchor at the headline shape
    "c5": (32, 32, 16384, 128, "int8", "strong"),  # long-context sweep: units split over ranks
}
METRIC = "attention fwd TFLOP/s at B8/H32/N8192/d128 INT8 (4*B*H*N^2*d FLOPs per step)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"src": "measured", "hbm": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"],
                "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"])}
    return {"src": "fallback", "hbm": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU golden path (tests/generate_golden.cpp cpu_mha compiled where
# it lies into oracle/_ref) or, when that is absent, the oracle's restatement of it.
# ------------------------------------------------------------------------------------------------
def cpu_step(n, d, heads, threads):
    """Runs `heads` independent single-head attention problems [n, d] on `threads` host threads.
    Returns (seconds, kind)."""
    import numpy as np
    from oracle import load_oracle, load_ref
    orc = load_oracle()
    ref = load_ref()
    q, k, v = orc.profile_inputs(n, d)  # inputs/data.cu generator, one head worth
    kind = "reference" if ref is not None else "port"

    def one(_):
        if ref is not None:
            ref.cpu_mha(q, k, v, 1)       # generate_golden.cpp:53-92, unmodified, single thread
        else:
            orc.mha(q, k, v, 1, "f32", threads=1)

    t0 = time.perf_counter()
    if threads == 1:
        for i in range(heads):
            one(i)
    else:
        from concurrent.futures import ThreadPoolExecutor  # ctypes releases the GIL in the call
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(one, range(heads)))
    return time.perf_counter() - t0, kind


def cpu_baseline(args):
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, args.cpu_threads or cores))
    heads = args.cpu_sample_heads or 16 * threads   # ~10-20 s of CPU work on the GPU box's cores
    n, d = args.cpu_sample_n, WORKLOADS[args.workload][3]
    sec, kind = cpu_step(n, d, heads, threads)
    flops = 4.0 * heads * n * n * d
    return {"value": flops / sec / 1e12, "unit": "TFLOP/s", "cores": threads, "kind": kind,
            "sample": f"{heads} heads of the workload truncated to N={n}, d={d}; {threads} threads, one head at a time each; "
                      f"{sec:.2f} s; reference cpu_mha (tests/generate_golden.cpp:53-92)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    B, H, N, d, kernel, scaling = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, args.cpu_threads or cores))
    heads = args.cpu_sample_heads or 4 * threads   # a few seconds per step: K+W steps stay within minutes
    n = args.cpu_sample_n
    kind = "port"
    for _ in range(args.warmup):
        cpu_step(n, d, heads, threads)
    total = 0.0
    for _ in range(args.steps):
        sec, kind = cpu_step(n, d, heads, threads)
        total += sec
    ms = total / max(args.steps, 1) * 1e3
    flops = 4.0 * heads * n * n * d
    val = flops / (ms / 1e3) / 1e12
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: B={B} H={H} N={N} d={d}; CPU step = bounded sample of "
                               f"{heads} heads truncated to N={n}"},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": threads, "kind": kind,
                         "sample": f"{heads} heads, N={n}, d={d}, {threads} threads, one head at a time each"},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_native(args):
    import torch
    import quantizedmha_b200 as qm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # Bind this rank to the CPUs next to its GPU (NUMA node of the PCIe root): the pinned host buffers of
    # the e2e arm are then first-touched in local memory instead of all ranks sharing one node.
    numa = "default"
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        numa = "cpu affinity set to the GPU's local CPUs (nvml)"
    except Exception:  # noqa: BLE001  (no nvml / not permitted: keep the default placement)
        pass
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    B, H, N, d, kernel, scaling = WORKLOADS[args.workload]
    from quantizedmha_b200.sharding import unit_range
    if scaling == "strong":
        # fixed total work: split the batch axis (units = B*H stay whole per batch entry)
        lo, hi = unit_range(B, world, rank)
        Bl = hi - lo
    else:
        Bl = B
    dm = H * d
    kid = qm.kernel_id(kernel)
    L = qm.lib()

    gen = torch.Generator(device=dev)
    gen.manual_seed(42 + rank)
    # U[0,1) like inputs/data.cu:15-22 (counter-based device RNG for the multi-GB shape)
    tq = torch.rand((Bl, N, dm), device=dev, generator=gen)
    tk = torch.rand((Bl, N, dm), device=dev, generator=gen)
    tv = torch.rand((Bl, N, dm), device=dev, generator=gen)
    out = torch.empty_like(tq)
    n_pad, d_pad = qm.workspace_dims(N, dm, H)
    units = Bl * H
    elt = torch.int8 if kernel == "int8" else torch.float16
    Qp = torch.empty((units, n_pad, d_pad), dtype=elt, device=dev)
    Kp = torch.empty_like(Qp)
    Vt = torch.empty((units, d_pad, n_pad), dtype=torch.float16, device=dev)
    gran = {"head": qm.GRAN_HEAD, "block": qm.GRAN_BLOCK, "tensor": qm.GRAN_TENSOR}[args.scales]
    sc = torch.empty((3, units, n_pad // 32) if gran == qm.GRAN_BLOCK else (3, units), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream()
    sp = int(stream.cuda_stream)

    def chk(rc):
        if rc != 0:
            raise RuntimeError(L.qmha_last_error().decode())

    def prep():
        if kernel == "int8":
            chk(L.qmha_quantize_qkv(tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), Bl, N, dm, H, gran,
                                    Qp.data_ptr(), Kp.data_ptr(), Vt.data_ptr(), sc.data_ptr(), sp))
        else:
            chk(L.qmha_convert_qkv_f16(tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), Bl, N, dm, H,
                                       Qp.data_ptr(), Kp.data_ptr(), Vt.data_ptr(), sp))

    def attn():
        chk(L.qmha_attention_prepared(Qp.data_ptr(), Kp.data_ptr(), Vt.data_ptr(),
                                      sc.data_ptr() if kernel == "int8" else None, out.data_ptr(),
                                      Bl, N, dm, H, kid, gran, sp))

    def barrier():
        if dist is not None:
            dist.barrier()

    for _ in range(max(args.warmup, 0)):
        prep(); attn()
    torch.cuda.synchronize()
    chk(L.qmha_check_async_error())

    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = qm.launch_count()
    barrier(); torch.cuda.synchronize()
    for i in range(K):
        ev[i][0].record(stream)
        prep()
        ev[i][1].record(stream)
        attn()
        ev[i][2].record(stream)
    torch.cuda.synchronize(); barrier()
    launches = qm.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    chk(L.qmha_check_async_error())
    total_ms = ev[0][0].elapsed_time(ev[K - 1][2])
    prep_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    attn_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    ms_step = total_ms / K

    # ---- e2e: host buffers through the C-ABI, H2D + compute + D2H inside the timed region.
    # Pinned host memory is capped at ~8.6 GB: larger workloads (c5) time the first `Be` batch
    # entries — the path pipelines per batch entry, so the rate is the same — and say so.
    e2e_steps = max(1, min(K, args.e2e_steps))
    Be = max(1, min(Bl, int(8.6e9 // (4 * N * dm * 4))))
    hq = torch.empty((Be, N, dm), dtype=torch.float32, pin_memory=True)
    hk = torch.empty_like(hq, pin_memory=True)
    hv = torch.empty_like(hq, pin_memory=True)
    ho = torch.empty_like(hq, pin_memory=True)
    hq.copy_(tq[:Be]); hk.copy_(tk[:Be]); hv.copy_(tv[:Be])
    torch.cuda.synchronize()
    chk(L.qmha_forward_host(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), ho.data_ptr(), Be, N, dm, H, kid, gran))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        chk(L.qmha_forward_host(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), ho.data_ptr(), Be, N, dm, H, kid, gran))
    e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3 * (Bl / Be)   # scaled to the full per-rank batch
    e2e_maxdiff = float((ho.to(dev) - out[:Be]).abs().max().item())

    t = torch.tensor([ms_step, attn_ms, prep_ms, e2e_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, attn_ms, prep_ms, e2e_ms = [float(x) for x in t.tolist()]

    flops_rank = 4.0 * Bl * H * N * N * d
    flops_all = 4.0 * (B if scaling == "strong" else B * world) * H * N * N * d
    E = Bl * N * dm
    prep_bytes = 3 * E * 4 + (3 * E if kernel == "int8" else 0) + (E * 2 if kernel == "int8" else 3 * E * 2)
    # int8: Q,K int8 (1 B) + V codes as fp16 (2 B) = 4E out; f16: 6E out.  3*E*4 in.
    prep_bytes = 3 * E * 4 + (4 * E if kernel == "int8" else 6 * E)
    pk = peaks()
    traffic = None
    try:   # per-launch DRAM traffic measured once with `ncu --set full` (profiles/r01/traffic.json)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01", "traffic.json")))
        traffic = tj["attn_fwd_kernel"].get(f"{args.workload}:{kernel}")
    except (OSError, KeyError, ValueError):
        pass
    value = flops_all / (ms_step / 1e3) / 1e12
    attn_tflops = flops_rank / (attn_ms / 1e3) / 1e12
    # Library INT8 GEMM rate of this box (cuBLASLt through torch._int_mm, best of 5, outside every timed
    # region): with the measured bf16 rate it gives the time-weighted ceiling of a kernel whose Q.K^T half
    # runs on the INT8 pipe and whose P.V half runs on the 16-bit pipe.
    int8_gemm = None
    if rank == 0 and kernel == "int8":
        try:
            ga = torch.randint(-127, 127, (8192, 8192), dtype=torch.int8, device=dev)
            gb = torch.randint(-127, 127, (8192, 8192), dtype=torch.int8, device=dev).t()
            torch._int_mm(ga, gb)
            best = 1e9
            for _ in range(5):
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(); torch._int_mm(ga, gb); g1.record(); torch.cuda.synchronize()
                best = min(best, g0.elapsed_time(g1))
            int8_gemm = 2.0 * 8192 ** 3 / (best / 1e3) / 1e12
            del ga, gb
        except Exception:  # noqa: BLE001  (not available on this torch build: leave the field empty)
            int8_gemm = None
    mixed_peak = 2.0 / (1.0 / int8_gemm + 1.0 / pk["bf16_sustained"]) if int8_gemm else None
    line = {
        "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "s8*s8->s32 (Q.K^T), f16*f16->f32 (P.V), f32 softmax" if kernel == "int8" else "f16*f16->f32, f32 softmax",
        "data": "synthetic U[0,1) (inputs/data.cu distribution), random on device",
        "config": {"workload": f"{args.workload}: B={B}{' per GPU' if scaling == 'weak' and world > 1 else ''} H={H} N={N} d={d} "
                               f"kernel={kernel} scales={args.scales}", "l2": f"inputs+outputs {4 * Bl * N * dm * 4 / 1e9:.2f} GB per GPU vs 126 MB L2 (no flush needed when larger)",
                   "parallelism": f"(batch x head) units sharded over {world} GPU(s), no collective"},
        "attn_ms": attn_ms, "attn_tflops_per_gpu": attn_tflops, "prep_ms": prep_ms,
        "prep_gbs_algorithmic": prep_bytes / (prep_ms / 1e3) / 1e9, "prep_frac_of_hbm": prep_bytes / (prep_ms / 1e3) / 1e9 / pk["hbm"],
        "roofline": {"bound": "tensor", "kernel": "attn_fwd_kernel", "achieved": attn_tflops, "peak": pk["bf16_sustained"],
                     "unit": "TFLOP/s", "frac": attn_tflops / pk["bf16_sustained"], "traffic": traffic,
                     "traffic_unit": "bytes of DRAM read+write per launch (ncu); algorithmic operand+output bytes = %d" % (E * (1 + 1 + 2) + E * 4) if kernel == "int8" else "bytes",
                     "peak_src": f"{pk['src']} dense bf16 cuBLAS GEMM, sustained (kernel timed inside the step loop)",
                     "frac_of_nominal_int8_4500": attn_tflops / 4500.0,
                     "frac_of_nominal_mixed_3000": attn_tflops / 3000.0,
                     "int8_gemm_tflops_measured": int8_gemm,
                     "frac_of_measured_mixed": (attn_tflops / mixed_peak) if mixed_peak else None,
                     "note": "peak is the bf16 GEMM rate of MEASURED_PEAKS.json; frac can exceed 1 for the INT8 kernel because "
                             "the Q.K^T half of the FLOPs runs on the INT8 pipe; frac_of_measured_mixed uses the harmonic "
                             "mean of the live INT8 GEMM rate and the bf16 rate (half of the FLOPs each)"},
        "e2e": {"value": flops_all / (e2e_ms / 1e3) / 1e12, "unit": "TFLOP/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 3 * E * 4, "d2h_bytes_per_step": E * 4, "steps": e2e_steps,
                "timed_batch_entries": Be, "of_batch_entries": Bl,
                "api": "qmha_forward_host (pinned host buffers, copies pipelined per batch entry)",
                "host_placement": numa,
                "max_abs_vs_device_path": e2e_maxdiff},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-n", type=int, default=2048)
    ap.add_argument("--cpu-sample-heads", type=int, default=0, help="0 = one head per host thread")
    ap.add_argument("--cpu-threads", type=int, default=0, help="0 = all host cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scales", default="block", choices=["head", "block", "tensor"],
                    help="granularity of the dynamic INT8 scales (block = the reference's 32-row tiles)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
