"""Loss objects of the joint UNet + prior step, callable like the reference's:
``loss(output, target)`` with ``output = cat([sigmoid(seg), sigmoid(prior)], dim=1)``
(``awesome/model/wrapper_module.py:157-228``).  These run on the segmentation side of the step (plain torch ops on
``[B,1,H,W]`` tensors next to a 13 M-parameter UNet); the prior's own per-pixel losses are folded into the fused fit
kernel (``LossConfig``).  Device-side MIOU for the in-loop checks lives in ``pretrain.mask_iou``."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F


class SE:
    """``awesome/measures/se.py:21-23``: squared error, ``reduction`` in {"mean", "sum", "none"}."""

    def __init__(self, reduction: str = "mean", **kwargs):
        self.reduction = reduction

    def __call__(self, output: torch.Tensor, target: torch.Tensor, **kwargs) -> torch.Tensor:
        l = (target - output) ** 2
        return l.mean() if self.reduction == "mean" else l.sum() if self.reduction == "sum" else l


class WeightedBCE:
    """``WeightedLoss(nn.BCELoss(), mode="sssdms", noneclass=2)`` (``awesome/measures/weighted_loss.py:41-92``): pixels
    of the none class are dropped, pixels with ``target == 0`` (foreground) weigh ``round(bg / fg / 10) + 1``."""

    def __init__(self, mode: str = "sssdms", noneclass: Optional[float] = 2.0):
        if mode not in ("sssdms", "equal", "none"):
            raise ValueError(f"Mode {mode} is not supported")
        self.mode, self.noneclass = mode, noneclass

    def __call__(self, output: torch.Tensor, target: torch.Tensor, **kwargs) -> torch.Tensor:
        o, t = output.reshape(-1), target.reshape(-1).to(output.dtype)
        if self.noneclass is not None:
            keep = t != self.noneclass
            o, t = o[keep], t[keep]
        l = F.binary_cross_entropy(o, t, reduction="none")
        if self.mode != "none":
            fg, bg = (t == 0).sum().to(o.dtype), (t == 1).sum().to(o.dtype)
            ratio = bg / fg
            if self.mode == "sssdms":
                ratio = torch.round(ratio / 10) + 1
            l = l * torch.where(t == 0, ratio, torch.ones_like(t))
        return l.mean()


class UnariesConversionLoss:
    """``awesome/measures/unaries_conversion_loss.py:20-22``: hard targets from soft unaries."""

    def __init__(self, criterion, **kwargs):
        self.criterion = criterion

    def __call__(self, output, target, **kwargs):
        return self.criterion(output, (target >= 0.5).float(), **kwargs)


class FBMSJointLoss:
    """``awesome/measures/fbms_joint_loss.py:35-59``: ``alpha * criterion(seg, target) + beta * penalty(prior, seg)``;
    the penalty is soft-clipped to the segmentation loss through a detached ratio, gradients of both terms reach the
    segmentation net and the prior."""

    def __init__(self, criterion=None, penalty_criterion=None, alpha: float = 1.0, beta: float = 1.0,
                 clip_penalty: bool = True, **kwargs):
        self.criterion = criterion if criterion is not None else WeightedBCE("sssdms", noneclass=2.0)
        self.penalty_criterion = penalty_criterion if penalty_criterion is not None else SE("mean")
        self.alpha, self.beta, self.clip_penalty = alpha, beta, clip_penalty
        self.last = {}

    def __call__(self, output: torch.Tensor, target: torch.Tensor, **kwargs) -> torch.Tensor:
        half = output.shape[1] // 2
        seg, pri = output[:, :half], output[:, half:]
        seg_raw = self.criterion(seg, target)
        pen_raw = self.penalty_criterion(pri, seg)
        seg_loss, pen = self.alpha * seg_raw, self.beta * pen_raw
        if self.clip_penalty and bool(pen > seg_loss):
            pen = pen * (seg_loss / pen).detach()
        self.last = {"segmentation_loss": seg_raw.detach(), "penalty_loss": pen_raw.detach()}
        return seg_loss + pen
