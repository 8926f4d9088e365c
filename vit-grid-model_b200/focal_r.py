"""Focal-R regression loss (README.md:16 of the reference names it; the reference has no implementation).

loss = mean(|e| * (2*sigmoid(beta*|e|) - 1)**gamma), e = pred - target   (Yang et al. 2021, L1 form; ``mse=True``
uses e**2).  Forward and backward are libvitgrid kernels (vg_focal_r_fwd / vg_focal_r_bwd)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops


class _FocalR(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, beta, gamma, mse):
        _lib.require_device()
        pred_c, target_c = pred.float().contiguous(), target.float().contiguous()
        ctx.save_for_backward(pred_c, target_c)
        ctx.cfg = (beta, gamma, mse)
        with torch.cuda.device(pred_c.device):             # kernels launch on the tensors' device, whatever the current one is
            return ops.focal_r_forward(pred_c, target_c, beta, gamma, mse)

    @staticmethod
    def backward(ctx, grad_out):
        pred, target = ctx.saved_tensors
        beta, gamma, mse = ctx.cfg
        with torch.cuda.device(pred.device):
            g = ops.focal_r_backward(pred, target, 1.0, beta, gamma, mse)
        return g * grad_out, None, None, None, None


def focal_r_loss(pred, target, beta: float = 0.2, gamma: float = 1.0, mse: bool = False):
    return _FocalR.apply(pred, target, beta, gamma, mse)


class FocalRLoss(nn.Module):
    def __init__(self, beta: float = 0.2, gamma: float = 1.0, mse: bool = False):
        super().__init__()
        self.beta, self.gamma, self.mse = beta, gamma, mse

    def forward(self, pred, target):
        return focal_r_loss(pred, target, self.beta, self.gamma, self.mse)
