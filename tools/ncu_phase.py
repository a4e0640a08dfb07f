#!/usr/bin/env python
"""Per-phase / per-opcode breakdown of an `ncu --page source --csv` SASS dump.

    python tools/ncu_phase.py src.csv k.sass '<mangled kernel substring>' main_file.cuh  lo-hi=name ...

Every SASS instruction is attributed to the last line of `main_file` seen in the nvdisasm line
info at or before it (inlined helpers carry header lines, so the enclosing kernel line is used),
then lines are bucketed into the given ranges.  Prints executed warp-instructions, stall samples
and the opcode mix per bucket."""
import csv, re, sys
from collections import defaultdict, Counter

src_csv, sass, kname, mainf = sys.argv[1:5]
ranges = []
for a in sys.argv[5:]:
    r, name = a.split("=")
    lo, hi = map(int, r.split("-"))
    ranges.append((lo, hi, name))

lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
cur_main = None
info = []
for l in lines[start + 1:]:
    if l.startswith("//---------------------") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if m.group(1).endswith(mainf):
            cur_main = int(m.group(2))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        txt = m.group(2).strip()
        t = txt.split()
        op = t[1] if t[0].startswith("@") else t[0]
        info.append((cur_main, op, txt))

rows = list(csv.reader(open(src_csv)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[h]
ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[h + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) > ci:
        data.append(r)
assert len(data) == len(info), (len(data), len(info))

def bucket(line):
    for lo, hi, name in ranges:
        if line is not None and lo <= line <= hi:
            return name
    return "other"

bi, bs = Counter(), Counter()
bop = defaultdict(Counter)
top = Counter()
for r, (ln, op, txt) in zip(data, info):
    n, s = int(r[ci]), int(r[si])
    b = bucket(ln)
    bi[b] += n; bs[b] += s
    base = op.split(".")[0]
    bop[b][base] += n
    top[base] += n
ti, ts = sum(bi.values()), sum(bs.values())
print(f"total warp-instr {ti}  samples {ts}")
for lo, hi, name in ranges + [(0, 0, "other")]:
    if bi[name] == 0: continue
    ops = ", ".join(f"{o} {c/ti*100:.1f}" for o, c in bop[name].most_common(12))
    print(f"{bi[name]/ti*100:6.2f}% instr {bs[name]/max(ts,1)*100:6.2f}% samples  {name:16s} | {ops}")
print("opcode mix:", ", ".join(f"{o} {c/ti*100:.1f}%" for o, c in top.most_common(30)))
