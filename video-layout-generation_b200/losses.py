"""Drop-in `nn.Module`s with the reference's call signatures.

    criterionL1 = L1Loss()                      # src/trainer.py:130   criterionL1(img, frame3)
    loss        = CombinedLoss()                # src/loss.py:54-62    loss(output=img, target=frame3)
    ce          = CrossEntropyLoss()            # src/trainer.py:124   ce(input=seg, target=seg3)

Each returns a 0-dim fp32 tensor and is differentiable w.r.t. its first argument.  `PixelLosses`
evaluates the whole src/trainer.py:248-251 composition in ONE kernel launch; `WarpLoss` is the
fused flow-guided op of SURVEY.md section 3.5.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _cabi
from .ops import WarpLossConfig, pixel_losses, warp_loss


class _Criterion(nn.Module):
    _mask = 0
    _slot = _cabi.LOSS_TOTAL

    def _cfg(self) -> WarpLossConfig:
        return WarpLossConfig(w_l1=0.0, w_gd=0.0, w_ssim=0.0, w_ce=0.0, w_tv=0.0, term_mask=self._mask)


class L1Loss(_Criterion):
    """torch.nn.L1Loss() as constructed at src/trainer.py:130."""
    _mask = _cabi.TERM_L1

    def forward(self, input, target):
        cfg = self._cfg()
        cfg.w_l1 = 1.0
        return pixel_losses(input, target, None, None, cfg)[0]


class GradientLoss(_Criterion):
    """src/loss.py:16-25."""
    _mask = _cabi.TERM_GD

    def forward(self, a, b):
        cfg = self._cfg()
        cfg.w_gd = 1.0
        return pixel_losses(a, b, None, None, cfg)[0]


class SsimLoss(_Criterion):
    """src/loss.py:64-91 (the unused `opt` argument is kept for signature parity)."""
    _mask = _cabi.TERM_SSIM

    def forward(self, x, y, opt=None):
        cfg = self._cfg()
        cfg.w_ssim = 1.0
        return pixel_losses(x, y, None, None, cfg)[0]


class CombinedLoss(_Criterion):
    """src/loss.py:54-62 without the VGG term (dense conv stack, out of scope: SURVEY section 2 #1).
    An optional torch `vgg` module can be passed in and is added exactly like the reference does."""
    _mask = _cabi.TERM_GD | _cabi.TERM_SSIM

    def __init__(self, vgg: Optional[nn.Module] = None):
        super().__init__()
        self.vgg = vgg

    def forward(self, output, target) -> torch.Tensor:
        cfg = self._cfg()
        cfg.w_gd = cfg.w_ssim = 1.0
        out = pixel_losses(output, target, None, None, cfg)[0]
        if self.vgg is not None:
            out = out + self.vgg(output, target)
        return out


class CrossEntropyLoss(_Criterion):
    """nn.CrossEntropyLoss(reduction='mean') as constructed at src/trainer.py:124."""
    _mask = _cabi.TERM_CE

    def __init__(self, weight: Optional[torch.Tensor] = None, ignore_index: int = -100, reduction: str = "mean"):
        """reduction='mean' is nn.CrossEntropyLoss (weighted mean when `weight` is given);
        reduction='sum_over_known' is the reference's class-weighted variant
        `F.cross_entropy(weight=w, reduction='sum') / n_known` (src/models/simple.py:56-59)."""
        super().__init__()
        if reduction not in ("mean", "sum_over_known"):
            raise ValueError("reduction must be 'mean' or 'sum_over_known'")
        self.ignore_index = ignore_index
        self.reduction = reduction
        self.register_buffer("weight", None if weight is None else weight.detach().float().contiguous())

    def forward(self, input, target):
        cfg = self._cfg()
        cfg.w_ce = 1.0
        cfg.ignore_index = self.ignore_index
        cfg.ce_norm = "count" if self.reduction == "sum_over_known" else "torch"
        if self.weight is not None:
            cfg.class_weight = self.weight.to(input.device)
        return pixel_losses(None, None, input, target, cfg)[0]


class PixelLosses(nn.Module):
    """40*L1 + 20*(GD+SSIM) + 10*CE of src/trainer.py:248-251 in one launch.
    forward(img, frame3, seg, seg3) -> total; the individual terms are in `.last_terms`."""

    def __init__(self, w_l1=40.0, w_style=20.0, w_ce=10.0, ignore_index=-100, want_argmax=False):
        super().__init__()
        self.cfg = WarpLossConfig(w_l1=w_l1, w_gd=w_style, w_ssim=w_style, w_ce=w_ce, ignore_index=ignore_index,
                                  want_argmax=want_argmax, term_mask=_cabi.TERM_L1 | _cabi.TERM_GD | _cabi.TERM_SSIM | _cabi.TERM_CE)
        self.last_terms = None
        self.last_argmax = None

    def forward(self, img, frame3, seg, seg3):
        total, terms, arg = pixel_losses(img, frame3, seg, seg3, self.cfg)
        self.last_terms, self.last_argmax = terms, arg
        return total


class WarpLoss(nn.Module):
    """Fused flow-guided warp + losses.  forward(src_rgb, src_layout, flow, tgt_rgb, tgt_label).
    `src_layout` is a [N,K,H,W] layout (one-hot or soft; differentiable) or the int64 [N,H,W] class-id map it
    is the one-hot of (the reference's data, src/models/net_utils.py:14-24; label sources are not differentiable).
    `debug=True` reads the device status word after every call (synchronises) and raises on out-of-range labels."""

    def __init__(self, weights=(40.0, 20.0, 10.0, 0.0), padding_mode="border", coords_are_grid=False,
                 ignore_index=-100, global_batch=0, assume_near=False, want_argmax=False, debug=False):
        super().__init__()
        w_l1, w_style, w_ce, w_tv = weights
        self.cfg = WarpLossConfig(w_l1=w_l1, w_gd=w_style, w_ssim=w_style, w_ce=w_ce, w_tv=w_tv,
                                  padding_mode=padding_mode, coords_are_grid=coords_are_grid,
                                  ignore_index=ignore_index, global_batch=global_batch,
                                  assume_near=assume_near, want_argmax=want_argmax, debug=debug)
        self.last_terms = None
        self.last_argmax = None

    def forward(self, src_rgb, src_layout, flow, tgt_rgb, tgt_label):
        total, terms, arg = warp_loss(src_rgb, src_layout, flow, tgt_rgb, tgt_label, self.cfg)
        self.last_terms, self.last_argmax = terms, arg
        return total
