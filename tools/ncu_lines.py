#!/usr/bin/env python
"""Joins an `ncu --page source --csv` SASS dump with `nvdisasm -g -c` line info and prints the
executed-instruction and stall-sample share per CUDA source line.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:pass1 > src.csv
    nvcc ... -cubin -o k.cubin x.cu && nvdisasm -g -c k.cubin > k.sass
    python tools/ncu_lines.py src.csv k.sass '<mangled kernel name substring>' [top]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# --- nvdisasm: instruction index -> (file, line) for the chosen function
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
loc = None
ins_loc = []
for l in lines[start + 1:]:
    if l.startswith("//---------------------") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        loc = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        ins_loc.append(loc)

rows = list(csv.reader(open(src_csv)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[h]
ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[h + 1:]:
    if r and r[0] == "Kernel Name":      # next launch of the report: keep the first one only
        break
    if len(r) > ci:
        data.append(r)
if len(data) != len(ins_loc):
    print(f"warning: {len(data)} profiled instructions vs {len(ins_loc)} disassembled", file=sys.stderr)
agg = defaultdict(lambda: [0, 0])
for r, l in zip(data, ins_loc):
    agg[l][0] += int(r[ci])
    agg[l][1] += int(r[si])
ti = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print(f"total warp-instructions {ti}, samples {ts}")
for l, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n / ti * 100:6.2f}% instr {s / max(ts, 1) * 100:6.2f}% samples  {l}")

# optional coarse grouping: python tools/ncu_lines.py ... top  file:lo-hi=name ...
groups = [a for a in sys.argv[5:] if "=" in a]
if groups:
    print("--- groups")
    rest_i, rest_s = ti, ts
    for gdef in groups:
        rng, name = gdef.split("=")
        f, lohi = rng.split(":")
        lo, hi = map(int, lohi.split("-"))
        gi = sum(v[0] for l, v in agg.items() if l and l[0] == f and lo <= l[1] <= hi)
        gs = sum(v[1] for l, v in agg.items() if l and l[0] == f and lo <= l[1] <= hi)
        rest_i -= gi
        rest_s -= gs
        print(f"{gi / ti * 100:6.2f}% instr {gs / max(ts, 1) * 100:6.2f}% samples  {name}")
    print(f"{rest_i / ti * 100:6.2f}% instr {rest_s / max(ts, 1) * 100:6.2f}% samples  (other)")
