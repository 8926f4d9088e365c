"""Fused AdamW over the model's flat parameter / gradient buffers (vg_adamw_step): one kernel per step instead of one
per tensor.  Parameters are re-pointed at slices of one flat fp32 buffer laid out like train.GradBuffer.

Gradient contract (the same as torch.optim): ``step()`` consumes ``p.grad``.  The network's backward hands autograd views of
one flat snapshot with GradBuffer's layout (train.MetNet3TrainFn), and autograd accumulates later backward passes into those
views in place, so in the common cases -- one backward per step, gradient accumulation over micro-batches,
``clip_grad_norm_`` / in-place gradient scaling, hooks that edit ``p.grad`` in place -- ``p.grad`` is still that flat tensor and
is read by the fused kernel directly.  If anything replaced a ``p.grad`` by another tensor (or set some to None), the
gradients are gathered into a flat scratch first (slow path, one copy per parameter; None counts as zero)."""
from __future__ import annotations

import torch

from . import ops_train as ot


class FlatAdamW:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.G = model.grad_buffer()
        self.params = dict(model.named_parameters())
        self.flat = torch.empty_like(self.G.flat)
        for name, gview in self.G.views.items():
            p = self.params[name]
            off = gview.storage_offset()
            dst = self.flat[off:off + p.numel()].view(p.shape)
            dst.copy_(p.data)
            p.data = dst                                  # the module now reads its weights from the flat buffer
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.step_count = 0
        self._gather = None

    def zero_grad(self, set_to_none: bool = True):
        for p in self.model.parameters():
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def _grad_flat(self):
        """the flat gradient the next step consumes (see the module docstring)"""
        if all(p.grad is None for p in self.params.values()):
            raise RuntimeError("FlatAdamW.step(): no gradients -- call loss.backward() first")
        flat = self.G.flat_of(self.params)
        if flat is not None:
            return flat
        if self._gather is None:
            self._gather = torch.zeros_like(self.G.flat)
        views = self.G.views_of(self._gather)
        for n, p in self.params.items():
            if p.grad is None:
                views[n].zero_()
            else:
                views[n].copy_(p.grad)
        return self._gather

    @torch.no_grad()
    def clip_grad_norm_(self, max_norm: float, eps: float = 1e-6):
        """torch.nn.utils.clip_grad_norm_ on the flat gradient (one reduction + one scale); returns the total norm"""
        g = self._grad_flat()
        total = torch.linalg.vector_norm(g)
        g.mul_(torch.clamp(max_norm / (total + eps), max=1.0))
        if g is self._gather:                              # slow path: write the clipped values back
            for n, v in self.G.views_of(g).items():
                if self.params[n].grad is not None:
                    self.params[n].grad.copy_(v)
        return total

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0):
        """one AdamW update of every parameter from ``p.grad``; `grad_scale` multiplies the gradients inside the kernel
        (loss-scale removal / averaging over accumulated micro-batches) at no extra pass"""
        g = self._grad_flat()
        self.step_count += 1
        with torch.cuda.device(self.flat.device):
            ot.adamw_step(self.flat, g, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps,
                          self.weight_decay, self.step_count, gscale=grad_scale)
        self.model.invalidate_packed()                    # kernel-layout weight copies are re-derived on the next forward

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.m.clone(), "exp_avg_sq": self.v.clone(),
                "layout": [(n, v.storage_offset(), v.numel()) for n, v in self.G.views.items()],
                "hyper": {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay}}

    def load_state_dict(self, sd):
        layout = [(n, v.storage_offset(), v.numel()) for n, v in self.G.views.items()]
        if [tuple(t) for t in sd["layout"]] != layout:
            raise ValueError("FlatAdamW.load_state_dict: the checkpoint's flat layout does not match this model")
        self.step_count = int(sd["step"])
        self.m.copy_(sd["exp_avg"])
        self.v.copy_(sd["exp_avg_sq"])
        h = sd.get("hyper", {})
        self.lr, self.betas = h.get("lr", self.lr), tuple(h.get("betas", self.betas))
        self.eps, self.weight_decay = h.get("eps", self.eps), h.get("weight_decay", self.weight_decay)
