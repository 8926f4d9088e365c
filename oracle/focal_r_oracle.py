"""Focal-R regression loss oracle (test infrastructure).

PARITY UNPINNED: the reference names "Focal-R loss" only in README.md:16 and
contains no implementation (no loss, optimizer or backward anywhere in
/root/reference/src).  This restates the published form from Yang et al. 2021,
"Delving into Deep Imbalanced Regression" (Focal-R, L1 variant):

    loss = mean( |e| * (2*sigmoid(beta*|e|) - 1)**gamma ),   e = pred - target

with beta = 0.2, gamma = 1 (the paper's defaults).  ``mse=True`` gives the
squared-error variant (e**2 in place of |e|).
"""
import torch


def focal_r(pred: torch.Tensor, target: torch.Tensor, beta: float = 0.2, gamma: float = 1.0,
            mse: bool = False) -> torch.Tensor:
    e = (pred - target).abs()
    w = (2.0 * torch.sigmoid(beta * e) - 1.0) ** gamma
    base = e * e if mse else e
    return (base * w).mean()


def focal_r_grad(pred: torch.Tensor, target: torch.Tensor, beta: float = 0.2, gamma: float = 1.0,
                 mse: bool = False) -> torch.Tensor:
    """d loss / d pred by autograd (fp64 for a clean reference)."""
    p = pred.detach().double().requires_grad_(True)
    focal_r(p, target.double(), beta, gamma, mse).backward()
    return p.grad.to(pred.dtype)
