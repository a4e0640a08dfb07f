#!/usr/bin/env python
"""Static SASS instruction counts per CUDA source line / line group (no GPU needed).

    nvcc ... -cubin -o k.cubin x.cu && nvdisasm -g -c k.cubin > k.sass
    python tools/sass_static.py k.sass '<mangled kernel substring>' [file:lo-hi=name ...]
"""
import re
import sys
from collections import Counter, defaultdict

sass, kname = sys.argv[1:3]
groups = [a for a in sys.argv[3:] if "=" in a]
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
loc, per_line, ops = None, Counter(), defaultdict(Counter)
n = 0
for l in lines[start + 1:]:
    if l.startswith("//---------------------") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        loc = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m2 = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*)", l)
    if m2:
        n += 1
        per_line[loc] += 1
        t = m2.group(1).split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[loc][op] += 1
print("static instructions:", n)
if groups:
    rest = n
    for g in groups:
        rng, name = g.split("=")
        f, lohi = rng.split(":")
        lo, hi = map(int, lohi.split("-"))
        c = sum(v for k, v in per_line.items() if k and k[0] == f and lo <= k[1] <= hi)
        oc = Counter()
        for k, v in ops.items():
            if k and k[0] == f and lo <= k[1] <= hi:
                oc.update(v)
        rest -= c
        print(f"{c:6d}  {name:24s} " + " ".join(f"{o}:{k}" for o, k in oc.most_common(8)))
    print(f"{rest:6d}  (other)")
else:
    for k, v in per_line.most_common(40):
        print(v, k)
