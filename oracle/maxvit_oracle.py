"""Functional fp32 restatement of /root/reference/src/maxvit.py (test infrastructure).

Written against explicit index maps instead of einops so that it is an
independent statement of the algorithm; the index maps themselves are checked
bit-exactly against einops-generated golden vectors (tests/golden/index_*.pt,
produced from the reference's own expressions by tests/golden/make_golden.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# integer index maps (bit-exact contract)
# ----------------------------------------------------------------------------
def rel_pos_indices(window: int, num_reg: int) -> torch.Tensor:
    """(num_reg+w*w, num_reg+w*w) int64 table index (maxvit.py:156-168).

    Token i = a*w+b at window position (a, b).  idx = (ai-aj+w-1)*(2w-1) + (bi-bj+w-1);
    every row/column that belongs to a register token shares index (2w-1)^2.
    """
    s = num_reg + window * window
    idx = torch.full((s, s), (2 * window - 1) ** 2, dtype=torch.int64)
    t = torch.arange(window * window)
    a, b = t // window, t % window
    da = a[:, None] - a[None, :] + window - 1
    db = b[:, None] - b[None, :] + window - 1
    idx[num_reg:, num_reg:] = da * (2 * window - 1) + db
    return idx


def block_pixel_index(H: int, W: int, w: int) -> torch.Tensor:
    """(X*Y, w*w) flat pixel id h*W+w of token t=a*w+b in window (x,y): pixel (x*w+a, y*w+b).
    Windows ordered x-major (maxvit.py:298, 306-307)."""
    X, Y = H // w, W // w
    x = torch.arange(X)[:, None, None, None]
    y = torch.arange(Y)[None, :, None, None]
    a = torch.arange(w)[None, None, :, None]
    b = torch.arange(w)[None, None, None, :]
    return ((x * w + a) * W + (y * w + b)).reshape(X * Y, w * w)


def grid_pixel_index(H: int, W: int, w: int) -> torch.Tensor:
    """(X*Y, w*w) flat pixel id of token (a,b) in window (x,y): pixel (a*X+x, b*Y+y)
    (dilated / grid partition, maxvit.py:322)."""
    X, Y = H // w, W // w
    x = torch.arange(X)[:, None, None, None]
    y = torch.arange(Y)[None, :, None, None]
    a = torch.arange(w)[None, None, :, None]
    b = torch.arange(w)[None, None, None, :]
    return ((a * X + x) * W + (b * Y + y)).reshape(X * Y, w * w)


# ----------------------------------------------------------------------------
# float path
# ----------------------------------------------------------------------------
def film(cond: torch.Tensor, sd, p: str):
    """FiLM MLP Linear->SiLU->Linear, split (gamma, beta) (maxvit.py:130-135,184)."""
    h = F.silu(F.linear(cond, sd[p + "film.0.weight"], sd[p + "film.0.bias"]))
    gb = F.linear(h, sd[p + "film.2.weight"], sd[p + "film.2.bias"])
    d = gb.shape[-1] // 2
    return gb[:, :d], gb[:, d:]


def attention(x: torch.Tensor, cond: torch.Tensor, sd, p: str, *, heads: int, window: int, num_reg: int,
              prob_mask: torch.Tensor | None = None, out_mask: torch.Tensor | None = None):
    """maxvit.py:170-219.  x (Nw, S, D) with Nw = N*windows (field-major), cond (N, cond_dim).
    Returns to_out(attn) WITHOUT the residual.  Dropout is identity (eval / p=0) unless explicit masks are given:
    prob_mask (Nw, heads, S, S) multiplies the softmax output (nn.Dropout of self.attend, :146), out_mask (Nw, S, D) the
    to_out output (:151); both already contain the 1/(1-p) scale."""
    Nw, S, D = x.shape
    N = cond.shape[0]
    if (p + "film.0.weight") in sd:
        x = F.layer_norm(x, (D,))                                     # no affine when cond_dim is set (:137)
        gamma, beta = film(cond, sd, p)
        rep = Nw // N
        x = x * gamma.repeat_interleave(rep, 0)[:, None, :] + beta.repeat_interleave(rep, 0)[:, None, :]
    else:                                                             # cond_dim=None: LayerNorm with affine, no FiLM (:128-137)
        x = F.layer_norm(x, (D,), sd[p + "norm.weight"], sd[p + "norm.bias"])
    qkv = F.linear(x, sd[p + "to_qkv.weight"])
    inner = qkv.shape[-1] // 3
    dh = inner // heads
    q, k, v = (t.reshape(Nw, S, heads, dh).permute(0, 2, 1, 3) for t in qkv.split(inner, dim=-1))
    q = F.normalize(q, dim=-1) * math.sqrt(dh) * sd[p + "q_norm.gamma"]       # RMSNorm (:18-30)
    k = F.normalize(k, dim=-1) * math.sqrt(dh) * sd[p + "k_norm.gamma"]
    sim = q @ k.transpose(-1, -2)                                              # no extra scale (:203)
    bias = sd[p + "rel_pos_bias.weight"][rel_pos_indices(window, num_reg)]     # (S,S,heads)
    sim = sim + bias.permute(2, 0, 1)
    attn = sim.softmax(dim=-1)
    if prob_mask is not None:
        attn = attn * prob_mask
    out = attn @ v
    out = out.permute(0, 2, 1, 3).reshape(Nw, S, inner)
    out = F.linear(out, sd[p + "to_out.0.weight"])
    return out if out_mask is None else out * out_mask


def mbconv(x: torch.Tensor, sd, p: str, *, residual: bool, training: bool = False, eps: float = 1e-5):
    """maxvit.py:75-102: 1x1 -> BN -> GELU -> DW3x3 -> BN -> GELU -> SE -> 1x1 -> BN (+x)."""
    def bn(t, i):
        return F.batch_norm(t, None if training else sd[p + f"{i}.running_mean"],
                            None if training else sd[p + f"{i}.running_var"],
                            sd[p + f"{i}.weight"], sd[p + f"{i}.bias"], training, 0.0, eps)
    h = F.gelu(bn(F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"]), 1))
    h = F.conv2d(h, sd[p + "3.weight"], sd[p + "3.bias"], padding=1, groups=h.shape[1])
    h = F.gelu(bn(h, 4))
    gate = torch.sigmoid(F.linear(F.relu(F.linear(h.mean(dim=(2, 3)), sd[p + "6.gate.1.weight"])),
                                  sd[p + "6.gate.3.weight"]))
    h = h * gate[:, :, None, None]
    h = bn(F.conv2d(h, sd[p + "7.weight"], sd[p + "7.bias"]), 8)
    return h + x if residual else h


def maxvit_forward(x: torch.Tensor, cond: torch.Tensor, sd, *, prefix: str = "", depth: int, heads: int,
                   window: int, num_reg: int, training: bool = False, return_registers: bool = False, drop_masks=None):
    """maxvit.py:289-341 for a single-stage MaxViT.  x (N,D,H,W), cond (N,cond_dim).
    drop_masks: optional {(layer, 1|2): (prob_mask, out_mask)} explicit dropout masks for block (1) / grid (2) attention."""
    drop_masks = drop_masks or {}
    N, D, H, W = x.shape
    assert H % window == 0 and W % window == 0
    nwin = (H // window) * (W // window)
    bidx = block_pixel_index(H, W, window).reshape(-1)
    gidx = grid_pixel_index(H, W, window).reshape(-1)
    regs_out = None
    for li in range(depth):
        # MBConvResidual (key level ".fn.") wraps every layer that is not the first of its stage (maxvit.py:99-100, 270);
        # with a tuple depth the first layer of a stage also changes the width
        residual = (f"{prefix}layers.{li}.0.fn.0.weight") in sd
        mb = f"{prefix}layers.{li}.0." + ("fn." if residual else "")
        x = mbconv(x, sd, mb, residual=residual, training=training)
        D = x.shape[1]
        flat = x.reshape(N, D, H * W).permute(0, 2, 1)                        # (N, HW, D)
        # block attention on (registers ++ window tokens); residual covers registers too (:310)
        tok = flat[:, bidx].reshape(N * nwin, window * window, D)
        reg = sd[f"{prefix}register_tokens.{li}"][None].expand(N * nwin, num_reg, D)
        seq = torch.cat([reg, tok], dim=1)
        pm, om = drop_masks.get((li, 1), (None, None))
        seq = attention(seq, cond, sd, f"{prefix}layers.{li}.1.", heads=heads, window=window, num_reg=num_reg,
                        prob_mask=pm, out_mask=om) + seq
        flat = torch.empty_like(flat)
        flat[:, bidx] = seq[:, num_reg:].reshape(N, nwin * window * window, D).to(flat.dtype)
        # grid attention: registers = mean over windows of block-attention register outputs (:326-327)
        reg = seq[:, :num_reg].reshape(N, nwin, num_reg, D).mean(dim=1)
        tok = flat[:, gidx].reshape(N * nwin, window * window, D)
        seq = torch.cat([reg.repeat_interleave(nwin, 0), tok], dim=1)
        pm, om = drop_masks.get((li, 2), (None, None))
        seq = attention(seq, cond, sd, f"{prefix}layers.{li}.2.", heads=heads, window=window, num_reg=num_reg,
                        prob_mask=pm, out_mask=om) + seq
        regs_out = seq[:, :num_reg]
        flat = torch.empty_like(flat)
        flat[:, gidx] = seq[:, num_reg:].reshape(N, nwin * window * window, D).to(flat.dtype)
        x = flat.permute(0, 2, 1).reshape(N, D, H, W)
    return (x, regs_out) if return_registers else x
