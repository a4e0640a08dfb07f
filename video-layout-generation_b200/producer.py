"""Flow-predicting producer for the training-step harness.

The reference's generator is a GridNet (src/models/gridnet.py:7-58: three resolution rows of
widths 32/64/96, six columns, residual 3x3-conv blocks with PReLU, stride-2 downsampling on the
left half and bilinear x2 upsampling on the right half, two output heads).  Dense convolutions are
library work (cuDNN) and out of scope for the hand-written path; this module re-expresses that
architecture in plain torch only so that the fused warp+loss op can be measured inside a real
training step: the image head is re-purposed as a 2-channel FLOW head.
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F


def _conv_pair(cin: int, cout: int, stride: int = 1) -> nn.Sequential:
    return nn.Sequential(nn.PReLU(), nn.Conv2d(cin, cout, 3, stride, 1), nn.PReLU(), nn.Conv2d(cout, cout, 3, 1, 1))


class Lateral(nn.Module):
    """Residual block; a 1x1-free projection conv on the skip when the width changes."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.body = _conv_pair(cin, cout)
        self.skip = nn.Conv2d(cin, cout, 3, 1, 1) if cin != cout else None

    def forward(self, x):
        return self.body(x) + (x if self.skip is None else self.skip(x))


class Down(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.body = _conv_pair(cin, cout, stride=2)

    def forward(self, x):
        return self.body(x)


class Up(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.body = _conv_pair(cin, cout)

    def forward(self, x, size):
        return self.body(F.interpolate(x, size=size, mode="bilinear", align_corners=True))


class FlowGridNet(nn.Module):
    """3 x `cols` grid of feature maps; returns (flow [N,2,H,W] in pixels, seg logits | None)."""

    def __init__(self, in_channels: int = 8, widths=(32, 64, 96), cols: int = 6, seg_classes: int = 0,
                 max_flow: float = 8.0):
        super().__init__()
        self.cols, self.max_flow = cols, max_flow
        w0, w1, w2 = widths
        half = cols // 2
        self.stem = Lateral(in_channels, w0)
        self.lat = nn.ModuleList([nn.ModuleList([Lateral(w, w) for _ in range(cols - 1)]) for w in widths])
        self.down01 = nn.ModuleList([Down(w0, w1) for _ in range(half)])
        self.down12 = nn.ModuleList([Down(w1, w2) for _ in range(half)])
        self.up21 = nn.ModuleList([Up(w2, w1) for _ in range(cols - half)])
        self.up10 = nn.ModuleList([Up(w1, w0) for _ in range(cols - half)])
        self.flow_head = Lateral(w0, 2)
        self.seg_head = Lateral(w0, seg_classes) if seg_classes else None
        nn.init.zeros_(self.flow_head.body[-1].weight)   # start from (almost) zero flow
        nn.init.zeros_(self.flow_head.body[-1].bias)

    def forward(self, x):
        half = self.cols // 2
        r0 = self.stem(x)
        r1 = self.down01[0](r0)
        r2 = self.down12[0](r1)
        for c in range(1, self.cols):
            if c < half:      # descending half: information flows down the rows
                r0 = self.lat[0][c - 1](r0)
                r1 = self.down01[c](r0) + self.lat[1][c - 1](r1)
                r2 = self.down12[c](r1) + self.lat[2][c - 1](r2)
            else:             # ascending half: information flows back up
                r2 = self.lat[2][c - 1](r2)
                r1 = self.up21[c - half](r2, r1.shape[-2:]) + self.lat[1][c - 1](r1)
                r0 = self.up10[c - half](r1, r0.shape[-2:]) + self.lat[0][c - 1](r0)
        flow = self.max_flow * torch.tanh(self.flow_head(r0))
        seg = self.seg_head(r0) if self.seg_head is not None else None
        return flow, seg


def flow_nhw2(flow_nchw: torch.Tensor) -> torch.Tensor:
    """[N,2,H,W] (channels_last storage) -> [N,H,W,2] without a copy when possible."""
    return flow_nchw.permute(0, 2, 3, 1).contiguous()
