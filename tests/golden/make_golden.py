"""Generates tests/golden/*.npz by running the REAL reference modules in the build container.

    python tests/golden/make_golden.py        # needs /root/reference (absent on the GPU box)

What is run, unmodified, from /root/reference/src/loss.py: `GradientLoss` (:16-25) and `SsimLoss`
(:64-91).  CE and L1 are constructed exactly as src/trainer.py:124,130 and composed with the
weights of src/trainer.py:248-250 (VGG term excluded: needs a download + CUDA).  The warp is
torch's `F.grid_sample(bilinear, align_corners=True)` on a grid built the src/models/modules.py:69
way (the reference has no warp of its own -- SURVEY.md section 0), and the TV term is the
first-difference stencil of src/loss.py:22,24 on the pixel-unit flow.

Nothing under oracle/ is imported here on purpose: the fixtures are what the oracle restatement is
pinned against (tests/test_oracle.py).  Seed convention 1024 follows src/main.py:121.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, "/root/reference/src")
import loss as ref_loss  # noqa: E402  (the real reference module)

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_grid(N, H, W):
    # src/models/modules.py:69-70 generalised from 256 to (H, W)
    xx = (torch.arange(W).repeat(1, H, 1).float() / (W - 1)) * 2 - 1          # [1,H,W]
    yy = ((torch.arange(H).repeat(1, W, 1).float() / (H - 1)) * 2 - 1).transpose(1, 2)
    return torch.stack([xx, yy], dim=-1).repeat(N, 1, 1, 1)


def make_case(name, N, H, W, K, padding, flow_kind, layout_kind, w_tv, seed=1024, ignore_frac=0.0):
    g = torch.Generator().manual_seed(seed)
    mean = torch.tensor([0.485, 0.456, 0.406])[None, :, None, None]   # src/trainer.py:122-123
    std = torch.tensor([0.229, 0.224, 0.225])[None, :, None, None]
    src_rgb = (torch.rand(N, 3, H, W, generator=g) - mean) / std
    tgt_rgb = (torch.rand(N, 3, H, W, generator=g) - mean) / std
    blk = 8
    lab_src = torch.randint(0, K, (N, (H + blk - 1) // blk, (W + blk - 1) // blk), generator=g)
    lab_src = lab_src.repeat_interleave(blk, 1).repeat_interleave(blk, 2)[:, :H, :W]
    lab_tgt = torch.randint(0, K, (N, (H + blk - 1) // blk, (W + blk - 1) // blk), generator=g)
    lab_tgt = lab_tgt.repeat_interleave(blk, 1).repeat_interleave(blk, 2)[:, :H, :W].contiguous()
    if ignore_frac > 0:
        m = torch.rand(N, H, W, generator=g) < ignore_frac
        lab_tgt[m] = -100
    if layout_kind == "onehot":
        src_layout = torch.eye(K)[lab_src.long()].permute(0, 3, 1, 2).contiguous()  # net_utils.py:23
    else:
        src_layout = torch.randn(N, K, H, W, generator=g)
    if flow_kind == "smooth":
        flow = torch.randn(N, 2, H, W, generator=g) * 4.0
        flow = F.avg_pool2d(F.pad(flow, (4, 4, 4, 4), mode="replicate"), 9, 1)
    elif flow_kind == "half":      # half-pixel adversarial: exact ties / integer coordinates
        flow = torch.round(torch.randn(N, 2, H, W, generator=g) * 3.0) / 2
    elif flow_kind == "large":     # large displacement, runs off the image
        flow = torch.randn(N, 2, H, W, generator=g) * 12.0
        m = torch.rand(N, 1, H, W, generator=g) < 0.05
        far = (torch.rand(N, 2, H, W, generator=g) - 0.5) * 2 * max(H, W)
        flow = torch.where(m, far, flow)
    elif flow_kind == "zero":
        flow = torch.zeros(N, 2, H, W)
    flow = flow.permute(0, 2, 3, 1).contiguous()          # [N,H,W,2] pixels

    a = src_rgb.clone().requires_grad_(True)
    b = src_layout.clone().requires_grad_(True)
    f = flow.clone().requires_grad_(True)
    scale = torch.tensor([2.0 / (W - 1), 2.0 / (H - 1)], dtype=torch.float64).float()
    grid = ref_grid(N, H, W) + f * scale
    w_rgb = F.grid_sample(a, grid, mode="bilinear", padding_mode=padding, align_corners=True)
    w_lay = F.grid_sample(b, grid, mode="bilinear", padding_mode=padding, align_corners=True)

    crit_l1 = torch.nn.L1Loss()                                   # src/trainer.py:130
    crit_ce = torch.nn.CrossEntropyLoss(reduction="mean")         # src/trainer.py:124
    crit_gd = ref_loss.GradientLoss()                             # src/loss.py:16
    crit_ssim = ref_loss.SsimLoss()                               # src/loss.py:64
    t_l1 = crit_l1(w_rgb, tgt_rgb)
    t_gd = crit_gd(w_rgb, tgt_rgb)
    t_ssim = crit_ssim(w_rgb, tgt_rgb)
    t_ce = crit_ce(input=w_lay, target=lab_tgt)
    t_tv = (f[:, 1:] - f[:, :-1]).abs().mean() + (f[:, :, 1:] - f[:, :, :-1]).abs().mean()
    total = t_l1 * 40 + (t_gd + t_ssim) * 20 + t_ce * 10 + w_tv * t_tv   # src/trainer.py:248-251
    total.backward()

    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        padding=padding, w_tv=np.float64(w_tv),
        src_rgb=src_rgb.numpy(), src_layout=src_layout.numpy(), flow=flow.numpy(),
        tgt_rgb=tgt_rgb.numpy(), tgt_label=lab_tgt.numpy(),
        grid=grid.detach().numpy(), warped_rgb=w_rgb.detach().numpy(),
        warped_layout=w_lay.detach().numpy(), argmax=torch.argmax(w_lay, dim=1).numpy(),
        terms=np.array([t_l1.item(), t_gd.item(), t_ssim.item(), t_ce.item(), t_tv.item()], np.float64),
        total=np.float64(total.item()),
        d_src_rgb=a.grad.numpy(), d_src_layout=b.grad.numpy(), d_flow=f.grad.numpy(),
        torch_version=torch.__version__,
    )
    print(name, "terms", [t.item() for t in (t_l1, t_gd, t_ssim, t_ce, t_tv)], "total", total.item())


if __name__ == "__main__":
    torch.set_num_threads(1)
    make_case("c1s_border_smooth_onehot", 2, 24, 40, 20, "border", "smooth", "onehot", 0.5)
    make_case("c1s_border_half_onehot", 1, 19, 33, 20, "border", "half", "onehot", 0.0)
    make_case("c1s_zeros_large_soft", 2, 21, 35, 20, "zeros", "large", "soft", 1.0, ignore_frac=0.1)
    make_case("c1s_border_large_soft", 1, 16, 48, 20, "border", "large", "soft", 0.25)
    make_case("c1s_border_zero_onehot", 1, 12, 20, 20, "border", "zero", "onehot", 0.0)
    make_case("k5_border_smooth_soft", 1, 13, 17, 5, "border", "smooth", "soft", 0.1)
