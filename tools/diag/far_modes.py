"""Diagnostic: packed vs wide far accumulators -- error against the fp64 oracle on fuzz-like cases, and the far-path
kernel times + segment-count histogram on the C5 flow (375x1242, sigma 48 px box-filtered 9x9, 5 % outliers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, vlg_b200
from oracle import torch_oracle as TO
dev = "cuda"
cl = lambda t: t.to(dev).contiguous(memory_format=torch.channels_last)
for (N, H, W, sigma, padding, seed) in ((2, 37, 70, 9.0, "zeros", 0), (1, 20, 33, 9.0, "border", 3), (2, 64, 96, 30.0, "zeros", 5)):
    g = torch.Generator().manual_seed(seed)
    K = 20
    a0 = torch.randn(N, 3, H, W, generator=g); b0 = torch.randn(N, K, H, W, generator=g); t0 = torch.randn(N, 3, H, W, generator=g)
    lab = torch.randint(0, K, (N, H, W), generator=g); flow = torch.randn(N, H, W, 2, generator=g) * sigma
    res = {}
    for mode in (True, False):
        a, b, f = cl(a0).requires_grad_(True), cl(b0).requires_grad_(True), flow.to(dev).requires_grad_(True)
        tot, vec, _ = vlg_b200.warp_loss(a, b, f, cl(t0), lab.to(dev), vlg_b200.WarpLossConfig(w_tv=0.7, padding_mode=padding, far_packed=mode))
        tot.backward()
        res[mode] = (a.grad.clone(), b.grad.clone())
    ref = TO.warp_loss_fwd_bwd(a0, b0, flow, t0, lab, w_tv=0.7, padding_mode=padding, dtype=torch.float64)
    for i, name in enumerate(("d_src_rgb", "d_src_layout")):
        r = ref[name].to(dev).float()
        mx = r.abs().max().item()
        ep = (res[True][i] - r).abs().max().item() / mx
        ew = (res[False][i] - r).abs().max().item() / mx
        d = (res[True][i] - res[False][i]).abs()
        print(f"{(N,H,W,sigma,padding)} {name}: packed err {ep:.2e} wide err {ew:.2e} packed-wide {d.max().item()/mx:.2e}")

# C5 flow: kernel times of both modes
from torch.profiler import profile, ProfilerActivity
N, H, W, K = 16, 375, 1242, 20
g = torch.Generator().manual_seed(7)
flow = torch.randn(N, 2, H, W, generator=g) * 48.0
flow = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(flow, (4, 4, 4, 4), mode="replicate"), 9, 1)
m = torch.rand(N, 1, H, W, generator=g) < 0.05
flow = torch.where(m, (torch.rand(N, 2, H, W, generator=g) - 0.5) * 2 * W, flow).permute(0, 2, 3, 1).contiguous().to(dev)
a0 = cl(torch.randn(N, 3, H, W, generator=g)); b0 = cl(torch.randn(N, K, H, W, generator=g)); t0 = cl(torch.randn(N, 3, H, W, generator=g))
lab = torch.randint(0, K, (N, H, W), generator=g).to(dev)
out = {}
for mode in (True, False):
    cfg = vlg_b200.WarpLossConfig(w_tv=0.7, far_packed=mode)
    def step():
        a, b, f = a0.detach().requires_grad_(True), b0.detach().requires_grad_(True), flow.detach().requires_grad_(True)
        tot, vec, _ = vlg_b200.warp_loss(a, b, f, t0, lab, cfg)
        tot.backward()
        return a.grad, b.grad
    for _ in range(3): step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): out[mode] = step()
        torch.cuda.synchronize()
    print("far_packed =", mode)
    for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:7]:
        print(f"  {ev.device_time_total / 5:9.1f} us  {ev.key[:70]}")
for i, name in enumerate(("d_src_rgb", "d_src_layout")):
    mx = out[False][i].abs().max().item()
    print(f"C5 {name}: packed-wide {(out[True][i] - out[False][i]).abs().max().item() / mx:.2e}")
