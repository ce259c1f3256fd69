"""Summarises an ncu report: key metrics + SASS opcode mix + hottest stall lines.
usage: python tools/ncu_hot.py report.ncu-rep [top_n]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = [r"^gpu__time_duration.sum$", r"sm__cycles_elapsed.avg$", r"sm__pipe_tensor_cycles_active_realtime.avg.pct", r"sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        r"sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", r"sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", r"sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        r"sm__issue_active.sum.pct", r"smsp__inst_executed.sum$", r"dram__bytes_(read|write).sum$", r"launch__registers_per_thread$", r"gpc__cycles_elapsed.avg.per_second",
        r"smsp__average_warps_issue_stalled_.*_per_issue_active", r"l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct", r"sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, vals):
    if any(re.search(w, h) for w in want):
        print(f"{h:90s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot_s = sum(int(r[ix["# Samples"]]) for r in data); tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
op = collections.Counter(); ops = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]]); k = m.group(2) if m else "?"
    op[k] += int(r[ix["Instructions Executed"]]); ops[k] += int(r[ix["# Samples"]])
print(f"--- total samples {tot_s}, warp instructions {tot_i}")
for k, v in op.most_common(22):
    print(f"{k:34s} {v / tot_i * 100:6.2f}% inst  {ops[k] / tot_s * 100:6.2f}% samples")
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter()
for r in data:
    for h in stall: agg[h] += int(r[ix[h]])
print("--- stall reasons (all samples):", ", ".join(f"{k[6:]}={v / tot_s * 100:.1f}%" for k, v in agg.most_common(10)))
print("--- hottest lines")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
    st = sorted(((h, int(r[ix[h]])) for h in stall), key=lambda kv: -kv[1])[:2]
    print(f"{int(r[ix['# Samples']]) / tot_s * 100:5.2f}% exec {int(r[ix['Instructions Executed']]):>10d} {r[ix['Source']].strip()[:64]:64s} {st}")
