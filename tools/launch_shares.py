"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.
usage: python tools/launch_shares.py launches.csv "description of the profiled command" """
import collections, csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
    name = r[ix["Kernel Name"]]
    m = re.search(r"(attn_fwd_kernel|block_quantize_kernel|block_aux_kernel|fused_quantize_kernel|stream_quantize_kernel|prepare_kernel|absmax_kernel|finalize_scales_kernel)", name)
    key = m.group(1) if m else "torch/other"
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v_ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    tot[key] += v_ms; cnt[key] += 1
ours = sum(v for k, v in tot.items() if k != "torch/other")
print(f"ncu launch list of `{sys.argv[2] if len(sys.argv) > 2 else '?'}`; per-launch times are cold-cache/serialised: compare shares.")
for k, v in sorted(tot.items(), key=lambda kv: kv[1]):
    share = f"{v / ours * 100:5.1f}%" if k != "torch/other" else "   - "
    print(f"{k:24s} launches {cnt[k]:3d}  total {v:9.3f} ms  each {v / cnt[k]:8.4f} ms  share of our kernels {share}")
