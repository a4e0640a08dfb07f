#!/usr/bin/env python
"""Fills the result placeholders of DESIGN.md section 8 from profiles/r01_bench_c2.json."""
import json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_c2.json")))
p = os.path.join(ROOT, "DESIGN.md")
s = open(p).read()
k = d["roofline_step"]["kernels_us_cupti"]
table = ", ".join(f"`{n}` {v:.1f} us" for n, v in sorted(k.items(), key=lambda kv: -kv[1]))
rep = {"RESULT_MS": f"{d['ms_per_step']:.3f}", "RESULT_MPX": f"{d['value']:,.0f}".replace(",", " "),
       "RESULT_FRAC": f"{d['roofline_step']['frac'] * 100:.1f}", "KERNEL_TABLE": table,
       "EAGER_MS": f"{d['torch_cuda_eager']['ms_per_step']:.2f}" if d.get("torch_cuda_eager") else "n/a",
       "CPU_MPX": f"{d['cpu_baseline']['value']:.1f}" if d.get("cpu_baseline") else "n/a"}
for a, b in rep.items():
    s = s.replace(a, b)
open(p, "w").write(s)
print(rep)
