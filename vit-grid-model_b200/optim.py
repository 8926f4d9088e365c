"""Fused AdamW over the model's flat parameter / gradient buffers (vg_adamw_step): one kernel per step instead of one
per tensor.  Parameters are re-pointed at slices of one flat fp32 buffer laid out like train.GradBuffer, so the
gradient buffer the backward kernels fill (already all-reduced in a data-parallel run) is consumed in place."""
from __future__ import annotations

import torch

from . import ops_train as ot


class FlatAdamW:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.G = model.grad_buffer()
        params = dict(model.named_parameters())
        self.flat = torch.empty_like(self.G.flat)
        for name, gview in self.G.views.items():
            p = params[name]
            off = gview.storage_offset()
            dst = self.flat[off:off + p.numel()].view(p.shape)
            dst.copy_(p.data)
            p.data = dst                                  # the module now reads its weights from the flat buffer
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.step_count = 0

    def zero_grad(self, set_to_none: bool = True):
        for p in self.model.parameters():
            p.grad = None

    @torch.no_grad()
    def step(self):
        """uses the flat gradient buffer written by the last backward"""
        self.step_count += 1
        ot.adamw_step(self.flat, self.G.flat, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps,
                      self.weight_decay, self.step_count)
        for p in self.model.parameters():                 # packed kernel-layout weights are re-derived on the next forward
            p._version  # noqa: B018  (in-place kernel writes do not bump versions; invalidate explicitly below)
        self.model._packed_key = None
        self.model.vit._packed_key = None
