#!/usr/bin/env python
"""Static instruction mix of the largest loops of a kernel: python tools/loopstat.py file.cu-or-cubin [kernel substring]"""
import re, subprocess, sys
from collections import Counter
src = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
if not src.endswith(".cubin"):
    out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-cubin", "-Xptxas", "-v", "-o", "/tmp/loopstat.cubin", src],
                         capture_output=True, text=True)
    print("\n".join(l for l in out.stderr.split("\n") if "registers" in l or "spill" in l))
    src = "/tmp/loopstat.cubin"
sass = subprocess.run(["cuobjdump", "-sass", src], capture_output=True, text=True).stdout
cur = None; funcs = {}
for l in sass.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, L in funcs.items():
    if sub not in name: continue
    addr = [a for a, _ in L]
    print(name[:80], "total", len(L))
    loops = []
    for i, (a, t) in enumerate(L):
        m = re.search(r"BRA.*?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr: loops.append((i - addr.index(tgt) + 1, addr.index(tgt), i))
    for n, j, i in sorted(loops, reverse=True)[:2]:
        c = Counter(); mv = 0
        for _, t in L[j:i + 1]:
            ws = t.split(); op = ws[1] if ws[0].startswith("@") else ws[0]
            c[op.split(".")[0]] += 1
            mv += op.startswith("IMAD.MOV") or op == "MOV"
        print(f"  loop {n} instrs, moves {mv}:", ", ".join(f"{k} {v}" for k, v in c.most_common(24)))
