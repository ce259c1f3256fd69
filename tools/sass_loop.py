"""Prints the instruction order of the hot softmax loop of a built library as a compact string, to see how
ptxas interleaved the MUFU work with the integer / FMA work.  usage: sass_loop.py lib.so [kernel-substring]
M = MUFU.EX2, v = VIADD/IADD3 (magic add), F = FFMA2, a = FADD2, p = F2FP, x = VIMNMX*, L = LDTM, S = STTM,
w = SYNCS / barrier ops, b = branch, . = anything else.  One line per basic block (split at branches/labels)."""
import re, subprocess, sys
lib = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else "ILb1ELi128ELi0ELb1ELb0"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None; funcs = {}
for ln in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", ln)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur: funcs[cur].append(m.group(2).strip())
for name, ins in funcs.items():
    if want not in name: continue
    print(name[:100], len(ins), "instructions")
    line = ""; blocks = []
    for i in ins:
        op = re.sub(r"^@!?U?P\d+\s+", "", i).split()[0]
        if op.startswith("MUFU"): c = "M"
        elif (op in ("VIADD", "IADD3")) and "0x4b400000" in i: c = "v"
        elif op.startswith("FFMA2"): c = "F"
        elif op.startswith("FADD2"): c = "a"
        elif op.startswith("F2FP"): c = "p"
        elif op.startswith("VIMNMX"): c = "x"
        elif op.startswith("LDTM"): c = "L"
        elif op.startswith("STTM"): c = "S"
        elif op.startswith("SYNCS") or op.startswith("BAR"): c = "w"
        elif op in ("BRA", "BRA.U", "EXIT", "BSYNC", "BSSY", "WARPSYNC.ALL", "BSYNC.RECONVERGENT", "BRA.DIV"): c = "b"
        else: c = "."
        line += c
        if c == "b": blocks.append(line); line = ""
    blocks.append(line)
    for b in blocks:
        if b.count("M") >= 16:
            runs = [len(r) for r in re.findall(r"M+", b)]
            print(f"[{len(b)} instr, {b.count('M')} MUFU, longest MUFU run {max(runs)}, runs>=4: {sum(1 for r in runs if r >= 4)}]")
            print(b)
