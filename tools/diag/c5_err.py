"""Diagnostic: where does the d_src_layout error of the C5-shaped case live?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import vlg_b200
from oracle import torch_oracle as TO
import test_gpu_parity as T
DEV = "cuda"
N, H, W, K = 1, 375, 1242, 20
d = T._make_case(N, H, W, K, 48.0, seed=1024, layout="soft", far_frac=0.05)
a = T._cl(d["src_rgb"]).requires_grad_(True); b = T._cl(d["src_layout"]).requires_grad_(True)
f = d["flow"].to(DEV).requires_grad_(True)
cfg = vlg_b200.WarpLossConfig(w_tv=0.1, padding_mode="border", want_argmax=True)
total, vec, arg = vlg_b200.warp_loss(a, b, f, T._cl(d["tgt_rgb"]), d["tgt_label"].to(DEV), cfg)
total.backward()
r32 = T._gpu_oracle(d, torch.float32, w_tv=0.1); r64 = T._gpu_oracle(d, torch.float64, w_tv=0.1)
for name, g, k in (("d_src_layout", b.grad, "d_src_layout"), ("d_src_rgb", a.grad, "d_src_rgb"), ("d_flow", f.grad, "d_flow")):
    g = g.double(); x32 = r32[k].double(); x64 = r64[k]
    e64 = (g - x64); e32 = g - x32; o = x32 - x64
    print(name, "max|ref|", x64.abs().max().item(), "rms ref", x64.square().mean().sqrt().item())
    print("   rms(got-r64) %.3e rms(got-r32) %.3e rms(r32-r64) %.3e" % tuple(t.square().mean().sqrt().item() for t in (e64, e32, o)))
    print("   max(got-r64) %.3e max(got-r32) %.3e max(r32-r64) %.3e" % tuple(t.abs().max().item() for t in (e64, e32, o)))
    if g.dim() == 4 and g.shape[1] in (3, 20):
        ep = e64.square().sum(1)[0]; rp = x64.square().sum(1)[0]     # per pixel
        tot = ep.sum().item()
        print("   share of squared error: top row %.3f bottom row %.3f left col %.3f right col %.3f interior %.3f" % (
            ep[0].sum().item() / tot, ep[-1].sum().item() / tot, ep[:, 0].sum().item() / tot, ep[:, -1].sum().item() / tot,
            ep[1:-1, 1:-1].sum().item() / tot))
        print("   share of squared ref  : border %.3f" % (1 - rp[1:-1, 1:-1].sum().item() / rp.sum().item()))
        idx = torch.topk(ep.flatten(), 8).indices
        for i in idx.tolist():
            y, x = divmod(i, W)
            print("     worst px (%d,%d): err2 %.3e ref2 %.3e got %s r64 %s r32 %s" % (y, x, ep[y, x].item(), rp[y, x].item(),
                  g[0, :3, y, x].tolist(), x64[0, :3, y, x].tolist(), x32[0, :3, y, x].tolist()))
        ei = e64[0, :, 1:-1, 1:-1]; ri = x64[0, :, 1:-1, 1:-1]
        print("   interior only: rms err / rms ref = %.3e ; max err / max ref = %.3e" % (
            (ei.square().mean().sqrt() / ri.square().mean().sqrt()).item(), (ei.abs().max() / ri.abs().max()).item()))
