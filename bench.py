#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json: attention fwd TFLOP/s & ms at
B8/H32/N8192/d128 INT8).

    python bench.py --gpus N --steps K --warmup W          # this framework (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W  # the reference's CPU path, host cores

One "step" = one pass of the hot path (INT8 quantise kernel + fused attention kernel, i.e. what
qmha_forward()/solve() enqueue) over one batch of synthetic inputs resident in HBM.  Under
torchrun (N>1) every rank owns its own (batch x head) units — no collective on the data path —
and the time is the max over ranks of the CUDA-event time of the K steps.  With N>1 the same
process also times BASELINE config 5 (B=32, N=16384) split over the ranks — the strong-scaling
sweep — and attaches it as `scaling_c5`.

Prints ONE JSON line (rank 0).  See DESIGN.md §6 for the definition of every field.
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
    "c4bf16": (8, 32, 8192, 128, "bf16", "weak"),  # BF16 anchor at the headline shape
    "c4pv8": (8, 32, 8192, 128, "int8_pv8", "weak"),  # INT8 kernel with INT8 P.V (the reference's P semantics), opt-in mode
    "c5": (32, 32, 16384, 128, "int8", "strong"),  # long-context sweep: units split over ranks
}
N_SM, MUFU_PER_SM_CLK = 148, 16   # B200: 148 SMs, 16 ex2 per SM per clock (4 per sub-partition)


def metric_name(workload):
    B, H, N, d, kernel, _ = WORKLOADS[workload]
    return f"attention fwd TFLOP/s at B{B}/H{H}/N{N}/d{d} {kernel.upper()} (4*B*H*N^2*d FLOPs per step)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"src": "MEASURED_PEAKS.json", "hbm": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"],
                "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"])}
    return {"src": "fallback of B200_PROFILING.md", "hbm": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU golden path (tests/generate_golden.cpp cpu_mha compiled where
# it lies into oracle/_ref) or, when that is absent, the oracle's restatement of it.
# ------------------------------------------------------------------------------------------------
def cpu_step(n, d, heads, threads):
    """Runs `heads` independent single-head attention problems [n, d] on `threads` host threads.
    Returns (seconds, kind)."""
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
        "impl": "reference", "metric": metric_name(args.workload), "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
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
        # "under load" = samples in the upper half of the observed power range (the sampler also sees idle time)
        if pw:
            thr = 0.5 * (min(pw) + max(pw))
            sm_load = sorted(s for s, p in zip(sm, pw) if p >= thr) or sorted(sm)
        else:
            sm_load = sorted(sm)
        return {"sm_mhz": sm_load[len(sm_load) // 2] if sm_load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


class Job:
    """Inputs, prepared operands and output of one workload on one GPU, driven through the C-ABI."""

    def __init__(self, torch, qm, dev, Bl, H, N, d, kernel, gran, seed, family="uniform"):
        self.torch, self.qm, self.L = torch, qm, qm.lib()
        self.Bl, self.H, self.N, self.d, self.kernel, self.gran = Bl, H, N, d, kernel, gran
        self.dm = H * d
        self.kid = qm.kernel_id(kernel)
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        shape = (Bl, N, self.dm)
        self.tq, self.tk, self.tv = (torch.empty(shape, device=dev) for _ in range(3))
        self.fill(family, gen)
        self.out = torch.empty(shape, device=dev)
        n_pad, d_pad = qm.workspace_dims(N, self.dm, H)
        units = Bl * H
        self.int8 = kernel in ("int8", "int8_pv8")
        elt = torch.int8 if self.int8 else torch.float16   # bf16 operands are 2 bytes as well
        self.Qp = torch.empty((units, n_pad, d_pad), dtype=elt, device=dev)
        self.Kp = torch.empty_like(self.Qp)
        self.Vt = torch.empty((units, d_pad, n_pad), dtype=torch.int8 if kernel == "int8_pv8" else torch.float16, device=dev)
        self.sc = torch.empty((3, units, n_pad // 32) if gran == qm.GRAN_BLOCK else (3, units), dtype=torch.float32, device=dev)
        self.stream = torch.cuda.current_stream()
        self.sp = int(self.stream.cuda_stream)

    def fill(self, family, gen):
        for t in (self.tq, self.tk, self.tv):
            if family == "uniform":     # U[0,1) like inputs/data.cu:15-22 (counter-based device RNG for the multi-GB shape)
                t.uniform_(0.0, 1.0, generator=gen)
            else:                       # 0.5*N(0,1) like tests/generate_golden.cpp:123-138: signed scores, moving row max
                t.normal_(0.0, 0.5, generator=gen)

    def chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.qmha_last_error().decode())

    def prep(self):
        if self.int8:
            self.chk(self.L.qmha_quantize_qkv_k(self.tq.data_ptr(), self.tk.data_ptr(), self.tv.data_ptr(), 0, self.Bl, self.N,
                                                self.dm, self.H, self.kid, self.gran, -1, 0.0, self.Qp.data_ptr(),
                                                self.Kp.data_ptr(), self.Vt.data_ptr(), self.sc.data_ptr(), self.sp))
        else:
            self.chk(self.L.qmha_convert_qkv_16(self.tq.data_ptr(), self.tk.data_ptr(), self.tv.data_ptr(), 0, self.Bl, self.N,
                                                self.dm, self.H, self.kid, -1, 0.0, self.Qp.data_ptr(), self.Kp.data_ptr(),
                                                self.Vt.data_ptr(), self.sp))

    def attn(self):
        self.chk(self.L.qmha_attention_prepared(self.Qp.data_ptr(), self.Kp.data_ptr(), self.Vt.data_ptr(),
                                                self.sc.data_ptr() if self.int8 else None, self.out.data_ptr(),
                                                self.Bl, self.N, self.dm, self.H, self.kid, self.gran, self.sp))

    def time(self, steps, warmup, barrier, attn_only=False):
        """-> (ms per step, attention ms, quantise ms, kernel launches), CUDA events on the launching stream"""
        torch = self.torch
        for _ in range(max(warmup, 0)):
            self.prep(); self.attn()
        torch.cuda.synchronize()
        self.chk(self.L.qmha_check_async_error())
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        l0 = self.qm.launch_count()
        barrier(); torch.cuda.synchronize()
        for i in range(steps):
            ev[i][0].record(self.stream)
            if not attn_only:
                self.prep()
            ev[i][1].record(self.stream)
            self.attn()
            ev[i][2].record(self.stream)
        torch.cuda.synchronize(); barrier()
        launches = self.qm.launch_count() - l0
        self.chk(self.L.qmha_check_async_error())
        total = ev[0][0].elapsed_time(ev[steps - 1][2])
        prep_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
        attn_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
        return total / steps, attn_ms, prep_ms, launches

    def parity(self, rng_seed=0, n_units=6, n_rows=5):
        """Checker (not timed, not on the product path): a sample of (batch, head, row) triples of the output the
        timed steps left behind, against the CPU oracle's per-row routine in float64 over all N keys."""
        import numpy as np
        from oracle import load_oracle
        orc = load_oracle()
        rng = np.random.default_rng(rng_seed)
        Bl, H, N, d = self.Bl, self.H, self.N, self.d
        units = {(0, 0), (Bl - 1, H - 1)}
        while len(units) < min(n_units, Bl * H):
            units.add((int(rng.integers(0, Bl)), int(rng.integers(0, H))))
        num = den = 0.0
        mx = 0.0
        rows_total = 0
        for b, head in sorted(units):
            rows = np.unique(np.concatenate([[0, N - 1], rng.integers(0, N, max(n_rows - 2, 0))])).astype(np.int64)
            sl = slice(head * d, (head + 1) * d)
            q, k, v = (t[b, :, sl].contiguous().cpu().numpy() for t in (self.tq, self.tk, self.tv))
            ref = orc.mha_head_rows(q[rows], k, v, "f64").astype(np.float64)
            got = self.out[b, :, sl].contiguous().cpu().numpy()[rows].astype(np.float64)
            if not np.isfinite(got).all():
                return {"max_abs": float("nan"), "rel_l2": float("nan"), "rows": int(rows_total), "ok": False}
            mx = max(mx, float(np.abs(got - ref).max()))
            num += float(((got - ref) ** 2).sum()); den += float((ref ** 2).sum())
            rows_total += len(rows)
        rel = (num / max(den, 1e-300)) ** 0.5
        tol_abs, tol_rel = {"int8": (2e-2, 1e-2), "int8_pv8": (2e-2, 1e-2), "f16": (2e-3, None), "bf16": (1e-2, None)}[self.kernel]
        ok = mx <= tol_abs and (tol_rel is None or rel <= tol_rel)
        return {"max_abs": mx, "rel_l2": rel, "rows": int(rows_total), "units": len(units), "ok": bool(ok),
                "checker": "oracle.mha_head_rows float64 (generate_golden.cpp:69-90 per-row routine), all N keys per row",
                "tolerance": {"max_abs": tol_abs, "rel_l2": tol_rel}}


def flops_of(B, H, N, d):
    return 4.0 * B * H * N * N * d


def copy_roof_ms(torch, dev, bufs_in, buf_out, reps=2):
    """Raw platform roof of the e2e arm: the same pinned buffers moved by one cudaMemcpyAsync each, all host->device
    copies on one stream and the device->host copy on another (full duplex), nothing else running."""
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    d_in = [torch.empty_like(b, device=dev) for b in bufs_in]
    d_out = torch.empty_like(buf_out, device=dev)
    best = None
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_in):
            for h, dd in zip(bufs_in, d_in):
                dd.copy_(h, non_blocking=True)
        with torch.cuda.stream(s_out):
            buf_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        best = ms if best is None else min(best, ms)
    return best


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

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    B, H, N, d, kernel, scaling = WORKLOADS[args.workload]
    from quantizedmha_b200.sharding import unit_range
    if scaling == "strong":
        # fixed total work: split the batch axis (units = B*H stay whole per batch entry)
        lo, hi = unit_range(B, world, rank)
        Bl = hi - lo
    else:
        Bl = B
    dm = H * d
    gran = {"head": qm.GRAN_HEAD, "block": qm.GRAN_BLOCK, "tensor": qm.GRAN_TENSOR}[args.scales]
    L = qm.lib()
    job = Job(torch, qm, dev, Bl, H, N, d, kernel, gran, 42 + rank)
    K = args.steps

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, attn_ms, prep_ms, launches = job.time(K, args.warmup, barrier)
    clocks = sampler.stop() if rank == 0 else None
    parity = job.parity() if rank == 0 else None

    # ---- e2e: host buffers through the C-ABI, H2D + compute + D2H inside the timed region.
    # Pinned host memory is capped at ~8.6 GB: larger workloads (c5) time the first `Be` batch
    # entries — the path pipelines per (batch entry, head group), so the rate is the same — and say so.
    e2e_steps = max(1, min(K, args.e2e_steps))
    gran_e2e = qm.GRAN_HEAD if gran == qm.GRAN_TENSOR else gran   # the host-buffer path chunks by (batch, head group)
    Be = max(1, min(Bl, int(8.6e9 // (4 * N * dm * 4))))
    hq = torch.empty((Be, N, dm), dtype=torch.float32, pin_memory=True)
    hk = torch.empty_like(hq, pin_memory=True)
    hv = torch.empty_like(hq, pin_memory=True)
    ho = torch.empty_like(hq, pin_memory=True)
    hq.copy_(job.tq[:Be]); hk.copy_(job.tk[:Be]); hv.copy_(job.tv[:Be])
    torch.cuda.synchronize()
    job.chk(L.qmha_forward_host(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), ho.data_ptr(), Be, N, dm, H, job.kid, gran_e2e))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        job.chk(L.qmha_forward_host(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), ho.data_ptr(), Be, N, dm, H, job.kid, gran_e2e))
    e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3 * (Bl / Be)   # scaled to the full per-rank batch
    e2e_maxdiff = float((ho.to(dev) - job.out[:Be]).abs().max().item())
    if gran_e2e == gran and not e2e_maxdiff <= 1e-6:
        raise SystemExit(f"bench.py: the host-buffer path and the device path disagree (max abs {e2e_maxdiff})")
    barrier()
    roof_ms = copy_roof_ms(torch, dev, [hq, hk, hv], ho) * (Bl / Be)
    # the same call for a 16-bit caller (qmha_forward_host_ex, fp16 host buffers in and out): half the PCIe bytes
    e2e16 = None
    if not args.no_e2e16:
        h16 = [t.to(torch.float16).pin_memory() for t in (hq, hk, hv)]
        ho16 = torch.empty((Be, N, dm), dtype=torch.float16, pin_memory=True)
        call16 = lambda: job.chk(L.qmha_forward_host_ex(h16[0].data_ptr(), h16[1].data_ptr(), h16[2].data_ptr(), ho16.data_ptr(),
                                                        Be, N, dm, H, job.kid, gran_e2e, 1, 1))
        call16()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            call16()
        e2e16_ms = (time.perf_counter() - t0) / e2e_steps * 1e3 * (Bl / Be)
        diff16 = float((ho16.to(dev).float() - job.out[:Be]).abs().max().item())   # vs the fp32-input device result
        (e2e16_ms,) = max_over_ranks([e2e16_ms])
        e2e16 = {"ms_per_step": e2e16_ms, "h2d_bytes_per_step": 3 * Bl * N * dm * 2, "d2h_bytes_per_step": Bl * N * dm * 2,
                 "api": "qmha_forward_host_ex, fp16 pinned host buffers in and out", "max_abs_vs_fp32_path": diff16}
        del h16, ho16
    del hq, hk, hv, ho

    ms_step, attn_ms, prep_ms, e2e_ms, roof_ms = max_over_ranks([ms_step, attn_ms, prep_ms, e2e_ms, roof_ms])

    # ---- the same kernel on signed inputs (0.5*N(0,1)): the lazy O-rescale path really runs there
    signed = None
    is_int8 = kernel in ("int8", "int8_pv8")
    if is_int8 and not args.no_signed:
        gen = torch.Generator(device=dev); gen.manual_seed(4242 + rank)
        job.fill("normal", gen)
        s_ms, s_attn, s_prep, _ = job.time(min(K, 5), 1, barrier)
        s_par = job.parity(1) if rank == 0 else None
        (s_attn,) = max_over_ranks([s_attn])
        signed = {"attn_ms": s_attn, "attn_tflops_per_gpu": 4.0 * Bl * H * N * N * d / (s_attn / 1e3) / 1e12,
                  "data": "0.5*N(0,1) (tests/generate_golden.cpp:123-138 distribution)", "parity": s_par}

    # ---- the opt-in INT8 P.V mode (P as 8-bit codes on the INT8 pipe: the reference's P semantics) on the same inputs
    pv8 = None
    if kernel == "int8" and not args.no_pv8:
        j8 = Job(torch, qm, dev, Bl, H, N, d, "int8_pv8", gran, 42 + rank)
        p_ms, p_attn, p_prep, _ = j8.time(min(K, 5), 2, barrier)
        p_par = j8.parity() if rank == 0 else None
        p_ms, p_attn, p_prep = max_over_ranks([p_ms, p_attn, p_prep])
        pv8 = {"kernel": "int8_pv8 (QMHA_KERNEL_INT8_PV8): s8*s8->s32 (Q.K^T), u8*s8->s32 (P.V)", "ms_per_step": p_ms, "attn_ms": p_attn,
               "prep_ms": p_prep, "tflops": flops_of(Bl, H, N, d) * (world if scaling == "weak" else 1) / (p_ms / 1e3) / 1e12,
               "attn_tflops_per_gpu": flops_of(Bl, H, N, d) / (p_attn / 1e3) / 1e12, "parity": p_par}
        del j8

    # ---- BASELINE config 5 (B=32, H=32, N=16384, d=128), units split over the ranks: strong scaling
    scaling_c5 = None
    if world > 1 and args.workload == "c4" and not args.no_c5:
        del job
        torch.cuda.empty_cache()
        B5, H5, N5, d5, k5, _ = WORKLOADS["c5"]
        lo, hi = unit_range(B5, world, rank)
        j5 = Job(torch, qm, dev, hi - lo, H5, N5, d5, k5, gran, 4200 + rank)
        c5_ms, c5_attn, c5_prep, _ = j5.time(min(K, 5), 2, barrier)
        c5_par = j5.parity(2, n_units=3, n_rows=3) if rank == 0 else None
        c5_ms, c5_attn = max_over_ranks([c5_ms, c5_attn])
        c5_tf = 4.0 * B5 * H5 * N5 * N5 * d5 / (c5_ms / 1e3) / 1e12
        one = None
        try:   # the committed 1-GPU line of the same workload (python bench.py --workload c5)
            one = json.load(open(os.path.join(ROOT, "profiles", "r02", "bench_c5_1gpu.json")))["ms_per_step"]
        except (OSError, KeyError, ValueError):
            pass
        scaling_c5 = {"workload": f"c5: B={B5} H={H5} N={N5} d={d5} INT8 scales={args.scales}, batch split over {world} GPUs",
                      "scaling": "strong", "ms_per_step": c5_ms, "attn_ms": c5_attn, "tflops": c5_tf, "steps": min(K, 5),
                      "ms_per_step_1gpu_line": one, "speedup_vs_1gpu_line": (one / c5_ms) if one else None,
                      "parity": c5_par}
        del j5

    # ---- fused gather (N > 1): C4 (B = 8 in total) split over the ranks, the 1.07 GB fp32 result replicated on every rank.
    # Not the headline path (the north star keeps collectives off it): it shows the compute -> all-gather pair as ONE kernel —
    # the attention epilogue TMA-stores every finished tile into all replicas over NVLink (qmha_args.peer_O, CUDA IPC) —
    # beside the same computation followed by an NCCL all-gather, and checks that both give the same bits.
    fused = None
    if world > 1 and args.workload == "c4" and not args.no_fused_gather and B % world == 0:
        try:
            torch.cuda.empty_cache()
            from quantizedmha_b200.sharding import ReplicatedOutput
            lo, hi = unit_range(B, world, rank)
            gen = torch.Generator(device=dev).manual_seed(7700 + rank)
            fq, fk, fv = (torch.rand((hi - lo, N, dm), device=dev, generator=gen) for _ in range(3))
            rep = ReplicatedOutput(B, N, H, d, dtype=torch.float32, device=dev)
            off = lo * N * dm * 4
            peers = [rep.peer_base[r] + off for r in sorted(rep.peer_base)]
            own = torch.empty_like(fq)
            gathered = torch.empty((B, N, dm), device=dev)

            def timed(fn, reps=5):
                fn(); fn()
                torch.cuda.synchronize(); barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return max_over_ranks([e0.elapsed_time(e1) / reps])[0]

            def f_compute():
                qm.forward(fq, fk, fv, H, kernel=kernel, gran=gran, out=own)

            def f_nccl():
                qm.forward(fq, fk, fv, H, kernel=kernel, gran=gran, out=own)
                dist.all_gather_into_tensor(gathered, own)

            def f_fused():
                rep.fence()
                qm.forward(fq, fk, fv, H, kernel=kernel, gran=gran, out=rep.local[lo:hi], peer_outs=peers)
                rep.fence()

            t_c, t_n, t_f = timed(f_compute), timed(f_nccl), timed(f_fused)
            torch.cuda.synchronize()
            qm.binding.check_async_error()
            same = bool(torch.equal(rep.local, gathered))
            flag = torch.tensor([1 if same else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            fused = {"workload": f"c4 split over {world} GPUs (B={B} in total), fp32 result replicated on every rank",
                     "replicated_bytes": B * N * dm * 4, "ms_own_slabs_only": t_c, "ms_compute_then_nccl_all_gather": t_n,
                     "ms_fused_epilogue_stores": t_f, "replicas_equal_nccl_result_on_every_rank": bool(int(flag.item())),
                     "api": "qmha_forward_ex with peer_O (TMA stores into CUDA-IPC mapped peer tensors) between two 4-byte NCCL fences"}
            barrier()
            rep.close()
            del rep, own, gathered, fq, fk, fv
        except Exception as e:  # noqa: BLE001  (an optional section must never take the bench line down)
            fused = {"error": f"{type(e).__name__}: {e}"[:300]}

    flops_rank = 4.0 * Bl * H * N * N * d
    flops_all = 4.0 * (B if scaling == "strong" else B * world) * H * N * N * d
    E = Bl * N * dm
    # quantise pass: SURVEY §8(d) algorithmic bytes = one fp32 read + one int8 write of Q, K, V = 15*E; the kernel
    # writes the V codes as fp16 (the P.V MMA is 16-bit), i.e. 16*E actually moved.  FP16/BF16: 12*E + 6*E.
    prep_alg = 15 * E if is_int8 else 18 * E
    prep_act = (16 * E if kernel == "int8" else 15 * E) if is_int8 else 18 * E
    pk = peaks()
    traffic = None
    for rnd in ("r02", "r01"):   # per-launch DRAM traffic from one `ncu --set full` capture of this command
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", rnd, "traffic.json")))
            traffic = tj["attn_fwd_kernel"].get(f"{args.workload}:{kernel}")
            if traffic is not None:
                break
        except (OSError, KeyError, ValueError):
            pass
    value = flops_all / (ms_step / 1e3) / 1e12
    attn_tflops = flops_rank / (attn_ms / 1e3) / 1e12
    # Library INT8 GEMM rate of this box (cuBLASLt through torch._int_mm, best of 5, outside every timed
    # region): with the measured bf16 rate it gives the time-weighted ceiling of a kernel whose Q.K^T half
    # runs on the INT8 pipe and whose P.V half runs on the 16-bit pipe.
    int8_gemm = None
    if rank == 0 and is_int8:
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
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    mufu_bound = N_SM * MUFU_PER_SM_CLK * sm_mhz * 1e6 * 4.0 * d / 1e12   # one ex2 per score element = 4d FLOPs
    if kernel == "int8_pv8":
        roof_peak = int8_gemm or 4500.0
        peak_src = ("INT8 GEMM rate measured live on this box (torch._int_mm 8192^3): both GEMMs of this mode run on the INT8 pipe"
                    if int8_gemm else "nominal 4500 (live INT8 GEMM measurement unavailable)")
    elif kernel == "int8":
        mixed_peak = 2.0 / (1.0 / int8_gemm + 1.0 / pk["bf16_sustained"]) if int8_gemm else 2.0 / (1.0 / 4500.0 + 1.0 / pk["bf16_sustained"])
        roof_peak = mixed_peak
        peak_src = (f"harmonic mean (half of the FLOPs each) of the INT8 GEMM rate measured live on this box "
                    f"({'torch._int_mm 8192^3' if int8_gemm else 'unavailable: nominal 4500'}) and the sustained dense bf16 "
                    f"rate of {pk['src']}: the Q.K^T half runs on the INT8 pipe, the P.V half on the 16-bit pipe")
    else:
        roof_peak = pk["bf16_sustained"]
        peak_src = f"{pk['src']} dense bf16 cuBLAS GEMM, sustained (kernel timed inside the step loop)"
    operand_bytes = (E * (1 + 1 + (2 if kernel == "int8" else 1)) + E * 4) if is_int8 else E * 6 + E * 4
    line = {
        "metric": metric_name(args.workload), "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": {"int8": "s8*s8->s32 (Q.K^T), f16*f16->f32 (P.V), f32 softmax",
                  "int8_pv8": "s8*s8->s32 (Q.K^T), u8*s8->s32 (P.V), f32 softmax", "f16": "f16*f16->f32, f32 softmax",
                  "bf16": "bf16*bf16->f32, f32 softmax"}[kernel],
        "data": "synthetic U[0,1) (inputs/data.cu distribution), random on device",
        "config": {"workload": f"{args.workload}: B={B}{' per GPU' if scaling == 'weak' and world > 1 else ''} H={H} N={N} d={d} "
                               f"kernel={kernel} scales={args.scales}", "l2": f"inputs+outputs {4 * Bl * N * dm * 4 / 1e9:.2f} GB per GPU vs 126 MB L2 (no flush needed when larger)",
                   "parallelism": f"(batch x head) units sharded over {world} GPU(s), no collective"},
        "attn_ms": attn_ms, "attn_tflops_per_gpu": attn_tflops, "prep_ms": prep_ms,
        "prep": {"kernel": "block_quantize_kernel" if (is_int8 and gran == qm.GRAN_BLOCK) else "quantise / convert",
                 "bound": "hbm", "peak_gbs": pk["hbm"], "algorithmic_bytes_15E": prep_alg, "moved_bytes_16E": prep_act,
                 "gbs_algorithmic": prep_alg / (prep_ms / 1e3) / 1e9, "frac_algorithmic": prep_alg / (prep_ms / 1e3) / 1e9 / pk["hbm"],
                 "gbs_moved": prep_act / (prep_ms / 1e3) / 1e9, "frac_moved": prep_act / (prep_ms / 1e3) / 1e9 / pk["hbm"]},
        "roofline": {"bound": "tensor", "kernel": "attn_fwd_kernel", "achieved": attn_tflops, "peak": roof_peak,
                     "unit": "TFLOP/s", "frac": attn_tflops / roof_peak, "traffic": traffic,
                     "traffic_note": f"DRAM read+write bytes per launch from the ncu --set full capture under profiles/; algorithmic operand+output bytes = {operand_bytes}",
                     "peak_src": peak_src,
                     "frac_of_nominal_int8_4500": attn_tflops / 4500.0 if is_int8 else None,
                     "frac_of_bf16_sustained": attn_tflops / pk["bf16_sustained"],
                     "int8_gemm_tflops_measured": int8_gemm,
                     "mufu_bound_tflops": mufu_bound, "frac_of_mufu_bound": attn_tflops / mufu_bound,
                     "mufu_note": f"{N_SM} SMs x {MUFU_PER_SM_CLK} ex2/clk x {sm_mhz:.0f} MHz (median SM clock under load) x 4d FLOPs per score element"},
        "e2e": {"value": flops_all / (e2e_ms / 1e3) / 1e12, "unit": "TFLOP/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 3 * E * 4, "d2h_bytes_per_step": E * 4, "steps": e2e_steps,
                "timed_batch_entries": Be, "of_batch_entries": Bl,
                "api": "qmha_forward_host (pinned host buffers; copies pipelined per batch entry x head group)",
                "host_placement": numa, "copy_roof_ms": roof_ms, "frac_of_copy_roof": roof_ms / e2e_ms,
                "copy_roof_note": "same pinned buffers, one cudaMemcpyAsync each, H2D and D2H on two streams, no compute",
                "max_abs_vs_device_path": e2e_maxdiff,
                "fp16_buffers": (dict(e2e16, value=flops_all / (e2e16["ms_per_step"] / 1e3) / 1e12, unit="TFLOP/s") if e2e16 else None)},
        "parity": parity,
        "signed_inputs": signed,
        "int8_pv_mode": pv8,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if scaling_c5 is not None:
        line["scaling_c5"] = scaling_c5
    if fused is not None:
        line["fused_gather"] = fused
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        print("bench.py: sampled rows of the timed output are outside the tolerance", file=sys.stderr)
        return 1
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
    ap.add_argument("--no-signed", action="store_true", help="skip the extra timing on signed inputs")
    ap.add_argument("--no-pv8", action="store_true", help="skip the extra timing of the opt-in INT8 P.V mode")
    ap.add_argument("--no-e2e16", action="store_true", help="skip the fp16-host-buffer variant of the e2e measurement")
    ap.add_argument("--no-c5", action="store_true", help="N>1: skip the C5 strong-scaling measurement")
    ap.add_argument("--no-fused-gather", action="store_true", help="N>1: skip the fused-gather (peer-memory epilogue) measurement")
    ap.add_argument("--scales", default="block", choices=["head", "block", "tensor"],
                    help="granularity of the dynamic INT8 scales (block = the reference's 32-row tiles)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
