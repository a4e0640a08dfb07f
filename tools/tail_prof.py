#!/usr/bin/env python
"""When do the warps / CTAs of the persistent pass-1 kernels start and finish?  Needs a tuning build with
-DVLG_PROFILE_TAIL (tools/variants.sh tail "-DVLG_PROFILE_TAIL"), selected with VLG_B200_LIB=build/variants/lib_tail.so.
    VLG_B200_LIB=build/variants/lib_tail.so python tools/tail_prof.py [workload]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from vlg_b200 import _cabi, ops as vops

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
N, H, W, K, sigma, far, dtype = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
lib = _cabi.load()
tdt = torch.float32 if dtype == "f32" else torch.bfloat16
sets = [bench.make_inputs(N, H, W, K, sigma, far, dtype, dev, seed=1024 + s) for s in range(2)]
prob = vops._problem(N, H, W, K, tdt, vops.WarpLossConfig(w_tv=0.5))
ws = vops._workspace(prob, True, dev)
loss = torch.zeros(_cabi.LOSS_SLOTS, dtype=torch.float32, device=dev)
d_c = torch.empty(N, H, W, 2, dtype=torch.float32, device=dev)
d_a, d_b = vops.empty_nhwc((N, 3, H, W), tdt, dev), vops.empty_nhwc((N, K, H, W), tdt, dev)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ptr = vops._ptr
def align(v, a=256):
    return (v + a - 1) // a * a
n_blocks = N * ((W + 31) // 32) * ((H + 7) // 8)
off = align(384)
off = align(off + n_blocks * 4); off = align(off + n_blocks * 4); off = align(off + n_blocks * 32)
off_rgb = off
off_lay = align(off_rgb + 8192 * 16)
res = {"rgb_strip (per warp)": [], "lay_tile (per CTA)": []}
for i in range(12):
    s = sets[i & 1]
    vops.check(lib.vlg_warp_loss_fwd_bwd(C.byref(prob), ptr(s["src_rgb"]), ptr(s["src_layout"]), ptr(s["flow"]), ptr(s["tgt_rgb"]),
                                         ptr(s["tgt_label"]), ptr(loss), ptr(d_c), ptr(d_a), ptr(d_b), None, ptr(ws), ws.numel(), sp))
    torch.cuda.synchronize()
    if i < 4:
        continue
    hdr = ws[:384].cpu().numpy()
    prof = hdr.view("uint64")[8:16]
    n_rgb, n_lay = int(hdr.view("uint32")[9]), int(hdr.view("uint32")[13])
    M = 0xFFFFFFFFFFFFFFFF
    for name, o, base, cnt in (("rgb_strip (per warp)", 0, off_rgb, n_rgb), ("lay_tile (per CTA)", 4, off_lay, n_lay)):
        t0 = ((M - int(prof[o])) >> 4) & 0xFFFFFF
        st = ws[base:base + cnt * 16].cpu().numpy().view("uint32").reshape(-1, 4)[:, 3].astype(np.int64)
        d = (((st & 0xFFFFFF) - t0) & 0xFFFFFF) * 16 / 1e3
        sm = st >> 24
        ok = d > 5.0              # units without work leave at once
        res[name].append(np.percentile(d[ok], [0, 10, 50, 90, 100]))
        if i == 11:
            per_sm = np.array([d[ok & (sm == k)].mean() for k in range(int(sm.max()) + 1) if (ok & (sm == k)).any()])
            spread_in_sm = np.array([d[ok & (sm == k)].max() - d[ok & (sm == k)].min() for k in range(int(sm.max()) + 1) if (ok & (sm == k)).any()])
            idx = np.arange(len(d))[ok]
            print(f"{name}: mean finish per SM: min {per_sm.min():.1f} median {np.median(per_sm):.1f} max {per_sm.max():.1f} us; "
                  f"spread inside an SM: median {np.median(spread_in_sm):.1f} us; correlation of finish time with unit index {np.corrcoef(idx, d[ok])[0, 1]:.2f}")
            q = np.array_split(d[ok], 8)
            print("   mean finish by eighths of the unit index:", " ".join(f"{x.mean():.1f}" for x in q))
for name, r in res.items():
    m = np.median(np.array(r), 0)
    print(f"{name}: units finish (us after the first start) min {m[0]:.1f}, p10 {m[1]:.1f}, median {m[2]:.1f}, p90 {m[3]:.1f}, max {m[4]:.1f}")
