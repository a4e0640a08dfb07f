#!/usr/bin/env python
"""Shared-memory wavefronts per opcode from `ncu --page source --csv`: python tools/ncu_smem.py src.csv rows"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[h]
iw, ie, ic, isrc = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("Instructions Executed"), hdr.index("Source")
agg = {}
for r in rows[h + 1:]:
    if len(r) <= iw or not r[iw].isdigit():
        continue
    w, e = int(r[iw]), int(r[ie] or 0)
    if w == 0:
        continue
    t = r[isrc].split()
    op = t[1] if t[0].startswith('@') else t[0]
    a = agg.setdefault(op, [0, 0, 0]); a[0] += w; a[1] += e; a[2] += int(r[ic])
tot = sum(v[0] for v in agg.values())
print("total shared wavefronts", tot, "per unit %.1f" % (tot / units))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:10s} wf {v[0]:>10d} excess {v[1]:>9d} instr {v[2]:>9d}  wf/instr {v[0]/max(v[2],1):.2f}  per unit {v[0]/units:.1f}")
