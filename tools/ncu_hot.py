#!/usr/bin/env python
"""Top SASS instructions by stall samples: python tools/ncu_hot.py src.csv k.sass '<kernel substr>' main.cuh [top]"""
import csv, re, sys
src_csv, sass, kname, mainf = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
cur_main = None; cur = None; info = []
for l in lines[start + 1:]:
    if l.startswith("//---------------------") or l.startswith("\t.section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        if m.group(1).endswith(mainf): cur_main = int(m.group(2))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: info.append((cur_main, cur, m.group(2).strip()))
rows = list(csv.reader(open(src_csv)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[h]
ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in rows[h + 1:] if len(r) > ci and r[0] != "Kernel Name"][:len(info)]
ts = sum(int(r[si]) for r in data)
order = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:top]
for i in order:
    r = data[i]
    st = sorted(((int(r[j] or 0), c[6:]) for j, c in stall_cols), reverse=True)[:3]
    print(f"{int(r[si])/ts*100:5.2f}%  L{info[i][0]} {info[i][1][0]}:{info[i][1][1]:<4}  {info[i][2][:60]:60s} " + " ".join(f"{n}={v}" for v, n in st if v))
