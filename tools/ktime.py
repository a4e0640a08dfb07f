#!/usr/bin/env python
"""Per-kernel device times of one fused step (CUPTI through torch.profiler; kernels launched via
ctypes are recorded too).  python tools/ktime.py [workload] [flags-int]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench, vlg_b200
from vlg_b200 import _cabi, ops as vops

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
N, H, W, K, sigma, far, dtype = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
lib = _cabi.load()
tdt = torch.float32 if dtype == "f32" else torch.bfloat16
sets = [bench.make_inputs(N, H, W, K, sigma, far, dtype, dev, seed=1024 + s) for s in range(2)]
cfg = vops.WarpLossConfig(w_tv=0.5)
prob = vops._problem(N, H, W, K, tdt, cfg)
prob.flags |= flags
ws = vops._workspace(prob, True, dev)
loss = torch.zeros(_cabi.LOSS_SLOTS, dtype=torch.float32, device=dev)
d_c = torch.empty(N, H, W, 2, dtype=torch.float32, device=dev)
d_a = vops.empty_nhwc((N, 3, H, W), tdt, dev)
d_b = vops.empty_nhwc((N, K, H, W), tdt, dev)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ptr = vops._ptr
def step(i):
    s = sets[i & 1]
    vops.check(lib.vlg_warp_loss_fwd_bwd(C.byref(prob), ptr(s["src_rgb"]), ptr(s["src_layout"]), ptr(s["flow"]),
               ptr(s["tgt_rgb"]), ptr(s["tgt_label"]), ptr(loss), ptr(d_c), ptr(d_a), ptr(d_b), None, ptr(ws), ws.numel(), sp))
for i in range(6): step(i)
torch.cuda.synchronize()
reps = 20
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(reps): step(i)
    torch.cuda.synchronize()
tot = 0.0
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if e.device_time_total <= 0: continue
    us = e.device_time_total / reps
    tot += us
    print(f"{us:9.1f} us  x{e.count / reps:4.1f}  {e.key[:110]}")
hdr = ws[:384].cpu().numpy().view("uint64")
prof = hdr[32:48]
if prof.any():
    names = ["plan:logic+tma", "taps+disp", "ring wait", "softmax", "grad dots", "obuf:issue", "tv+dcoords", "plan:ldg issue", "plan:xy+redux", "label+addr", "gather", "obuf:sts", "obuf:wait_read", "-", "-", "-"]
    print(f"fallback lanes {int(prof[13])}, rows with a fallback lane {int(prof[14])} of {int(prof[15])} warp-rows")
    prof = prof.copy(); prof[13:] = 0
    tot_c = float(prof.sum())
    print("lay_strip sections (cycles summed over warps, last step): " + ", ".join(f"{n} {c / tot_c * 100:.1f}%" for n, c in zip(names, prof) if c))
print(f"{tot:9.1f} us  total per step   loss {loss.tolist()[:6]}")
