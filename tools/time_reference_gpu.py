"""Times the REFERENCE's own GPU kernels (built for sm_100 by baseline/Makefile from the sources in
/root/reference; binaries in the git-ignored baseline/_ref/) against this repository's driver on
the same B200 and the same shape (the reference's compile-time config: N=8192, d_model=1024,
h=32 => d=32, B=1).  The reference driver takes no timing, so solve() time is the wall-clock
difference between --runs=11 and --runs=1 divided by 10 (input generation and H2D cancel out).
Writes gpurun_out/ref_gpu.json."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out"); os.makedirs(OUT, exist_ok=True)
SCRATCH = tempfile.mkdtemp(prefix="qmha_refgpu_")   # the binaries write a 100 MB .cache/ here

def wall(cmd):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, cwd=SCRATCH, capture_output=True, text=True)
    return time.perf_counter() - t0, r

res = {"shape": {"B": 1, "N": 8192, "d_model": 1024, "h": 32, "d": 32}, "flops": 4.0 * 32 * 8192 * 8192 * 32, "rows": []}
for k in ("fa_tc_int8_b", "fa_tc_v1b", "fa"):
    exe = os.path.join(ROOT, "baseline", "_ref", f"profile_{k}")
    if not os.path.exists(exe):
        continue
    wall([exe, "--no-check", "--warmup=1", "--runs=1"])          # populate .cache/input_random_*.bin
    best = None
    for _ in range(3):
        t1, r1 = wall([exe, "--no-check", "--warmup=1", "--runs=1"])
        t11, r11 = wall([exe, "--no-check", "--warmup=1", "--runs=11"])
        d = (t11 - t1) / 10 * 1e3
        best = d if best is None else min(best, d)
    ms = best
    res["rows"].append({"impl": f"reference {k} (WMMA/SIMT, sm_100 build)", "solve_ms": ms,
                        "tflops": res["flops"] / ms / 1e9, "rc": r11.returncode})
for k in ("fa_tc_int8_b", "fa_tc_v2a"):
    exe = os.path.join(ROOT, "bin", f"profile_{k}")
    if not os.path.exists(exe):
        continue
    t, r = wall([exe, "--no-check", "--warmup=3", "--runs=20", "--json", "--N=8192", "--d_model=1024", "--h=32"])
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if line:
        d = json.loads(line[-1])
        res["rows"].append({"impl": f"this repo {k} (tcgen05)", "solve_ms": d["ms_median"], "tflops": d["tflops_median"], "rc": r.returncode})
json.dump(res, open(os.path.join(OUT, "ref_gpu.json"), "w"), indent=1)
for row in res["rows"]:
    print(f"{row['impl']:48s} solve() {row['solve_ms']:9.3f} ms  {row['tflops']:8.2f} TFLOP/s  rc={row['rc']}")
