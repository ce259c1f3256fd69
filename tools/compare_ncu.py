#!/usr/bin/env python3
"""Side-by-side markdown report of two (or more) Nsight Compute captures.

Same purpose as the reference's tools/compare_ncu.py (it lines up the text tables of two `ncu`
outputs, tools/compare_ncu.py:1-193); this one works on *metrics*, so the inputs may be any mix of

  * a `.ncu-rep` report              (read through `ncu -i <rep> --page raw --csv`; needs ncu on PATH),
  * a raw-page CSV saved from that command,
  * a `metric value unit` summary as committed under profiles/ (tools/ncu_hot.py output).

    python tools/compare_ncu.py A B [C ...] [--names a,b,c] [--kernel REGEX] [--all] [-o report.md]

Rows are the union of the metrics, the first capture is the baseline of the delta columns.  Without
--all only the metrics that explain a kernel on this GPU are kept (time, cycles, clocks, tensor / MUFU /
ALU / FMA pipe use, issue slots, DRAM bytes, registers, stall reasons).
"""
import argparse
import csv
import io
import re
import subprocess
import sys

KEEP = [r"gpu__time_duration", r"sm__cycles_elapsed\.avg$", r"cycles_elapsed\.avg\.per_second", r"pipe_tensor", r"mem_tensor",
        r"inst_executed_pipe_(xu|alu|fma|fmaheavy)\.", r"issue_active", r"dram__bytes_(read|write)\.sum$", r"dram__throughput",
        r"registers_per_thread", r"issue_stalled_.*_per_issue_active", r"smsp__inst_executed\.sum$", r"wavefronts_mem_shared"]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
              "hz": 1.0, "Khz": 1e3, "Mhz": 1e6, "Ghz": 1e9}


def _num(text):
    try:
        return float(text.replace(",", ""))
    except ValueError:
        return None


def _normalise(value, unit):
    """bytes / seconds / hertz are brought to one scale so that captures taken at different sizes compare."""
    if value is not None and unit in UNIT_SCALE:
        base = "byte" if "byte" in unit else ("hz" if "hz" in unit.lower() else "s")
        return value * UNIT_SCALE[unit], base
    return value, unit


def parse_summary(text):
    """`metric value unit` lines (profiles/*.txt).  Lines that do not look like a metric are ignored."""
    out = {}
    for line in text.splitlines():
        parts = line.split()
        if len(parts) >= 2 and re.match(r"^[A-Za-z_][\w.]*$", parts[0]) and _num(parts[1]) is not None:
            out[parts[0]] = _normalise(_num(parts[1]), parts[2] if len(parts) > 2 else "")
    return out


def parse_raw_csv(text, kernel=None):
    """Raw page of `ncu --csv`: header row, unit row, then one row per profiled launch.  The launches that
    match `kernel` (all, when None) are averaged."""
    rows = [r for r in csv.reader(io.StringIO(text)) if r]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[start], rows[start + 1], rows[start + 2:]
    kcol = hdr.index("Kernel Name")
    if kernel:
        data = [r for r in data if re.search(kernel, r[kcol])]
    if not data:
        raise SystemExit(f"no launch matches --kernel {kernel!r}")
    out = {}
    for c, (h, u) in enumerate(zip(hdr, units)):
        vals = [_num(r[c]) for r in data if c < len(r)]
        vals = [v for v in vals if v is not None]
        if vals and re.match(r"^[A-Za-z_][\w.]*$", h) and "__" in h:
            out[h] = _normalise(sum(vals) / len(vals), u)
    return out


def load(path, kernel=None):
    if path.endswith(".ncu-rep"):
        res = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
        if res.returncode != 0:
            raise SystemExit(f"ncu could not read {path}: {res.stderr.strip()[:200]}")
        return parse_raw_csv(res.stdout, kernel)
    text = open(path).read()
    if "Kernel Name" in text.split("\n", 3)[0] or '"Kernel Name"' in text[:4096]:
        return parse_raw_csv(text, kernel)
    return parse_summary(text)


def _fmt(v, unit=""):
    if v is None:
        return "–"
    if (unit not in ("byte", "s", "hz", "cycle", "inst")) or (unit == "inst" and abs(v) < 1e3):   # percentages, ratios, counts: as they are
        return f"{v:.4g}"
    a = abs(v)
    if a >= 1e9:
        return f"{v / 1e9:.3f} G"
    if a >= 1e6:
        return f"{v / 1e6:.3f} M"
    if a >= 1e3:
        return f"{v / 1e3:.3f} k"
    if a and a < 1e-3:
        return f"{v * 1e6:.3f} µ"
    if a and a < 1:
        return f"{v * 1e3:.3f} m"
    return f"{v:.3f}"


def report(caps, names, keep_all=False):
    keys = []
    for cap in caps:
        for k in cap:
            if k not in keys and (keep_all or any(re.search(p, k) for p in KEEP)):
                keys.append(k)
    keys.sort()
    head = ["metric", "unit"] + names + [f"Δ {n} vs {names[0]}" for n in names[1:]]
    lines = ["| " + " | ".join(head) + " |", "|" + "---|" * len(head)]
    for k in keys:
        vals = [cap.get(k, (None, ""))[0] for cap in caps]
        unit = next((cap[k][1] for cap in caps if k in cap), "")
        deltas = []
        for v in vals[1:]:
            if v is None or vals[0] in (None, 0):
                deltas.append("–")
            else:
                deltas.append(f"{(v / vals[0] - 1) * 100:+.1f} %")
        lines.append("| " + " | ".join([f"`{k}`", unit] + [_fmt(v, unit) for v in vals] + deltas) + " |")
    return "\n".join(lines) + "\n"


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("captures", nargs="+")
    ap.add_argument("--names", default="")
    ap.add_argument("--kernel", default=None, help="regex on the kernel name (raw CSV / .ncu-rep inputs)")
    ap.add_argument("--all", action="store_true", help="keep every metric")
    ap.add_argument("-o", "--output", default=None)
    a = ap.parse_args(argv)
    if len(a.captures) < 2:
        ap.error("need at least two captures")
    names = a.names.split(",") if a.names else [re.sub(r"\.(ncu-rep|csv|txt)$", "", c.rsplit("/", 1)[-1]) for c in a.captures]
    if len(names) != len(a.captures):
        ap.error("--names must list one name per capture")
    md = report([load(c, a.kernel) for c in a.captures], names, a.all)
    if a.output:
        open(a.output, "w").write(md)
    else:
        sys.stdout.write(md)
    return 0


if __name__ == "__main__":
    sys.exit(main())
