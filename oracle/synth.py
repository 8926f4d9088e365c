"""Synthetic configs, state dicts and CMAQ-shaped inputs (test infrastructure).

The reference ships neither its checkpoint (/root/reference/.MISSING_LARGE_BLOBS)
nor its data, so parity runs on seeded synthetic tensors of the reference's
shapes.  Weights are filled per *key name* (crc32-seeded), not
per constructor order, so the reference module, the oracle and the CUDA module
all receive bit-identical parameters from ``make_state_dict`` with no
dependency on how any of them builds its layers.

Key names / shapes follow the reference constructors:
/root/reference/src/metnet3.py:192-321 and /root/reference/src/maxvit.py:224-287
(SURVEY.md appendix A).  ``tests/golden/make_golden.py`` proves the spec by
loading it into the real reference with ``strict=True``.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from dataclasses import dataclass, asdict

import torch


@dataclass(frozen=True)
class GridConfig:
    """Constructor arguments of the reference ``MetNet3`` that shape the path."""
    T: int = 25            # window_size (time steps in the input)
    C: int = 24            # n_variables (4 CMAQ runs x 6 species)
    H: int = 82
    W: int = 67
    dim: int = 128         # n_start_channels
    L: int = 12            # end_lead_time
    lead_emb: int = 2
    time_emb: int = 1
    resnet_depth: int = 2
    vit_depth: int = 1
    heads: int = 32
    dim_head: int = 32
    window: int = 7
    expansion: float = 4
    shrink: float = 0.25
    num_reg: int = 4
    pm25_mean: float = 20.0
    pm25_std: float = 15.0

    # derived -----------------------------------------------------------
    @property
    def c_in(self) -> int:
        return self.T * self.C + self.lead_emb + 3 * self.time_emb

    @property
    def pads(self):
        """(left, right, top, bottom) as metnet3.py:324-333 computes them."""
        ph = (14 - self.H) % 14
        pw = (14 - self.W) % 14
        return (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2)

    @property
    def HP(self) -> int:
        return self.H + (14 - self.H) % 14

    @property
    def WP(self) -> int:
        return self.W + (14 - self.W) % 14

    def metnet3_kwargs(self) -> dict:
        """kwargs for the reference-compatible ``MetNet3`` constructor."""
        return dict(
            input_size_sample=(self.T, self.C, self.H, self.W),
            n_start_channels=self.dim, end_lead_time=self.L,
            pm25_boundaries=[15, 35, 75], pm10_boundaries=[15, 35, 75],
            pm25_mean=self.pm25_mean, pm25_std=self.pm25_std,
            lead_time_emb_dim=self.lead_emb, model_time_emb_dim=self.time_emb,
            resnet_block_depth=self.resnet_depth, vit_block_depth=self.vit_depth,
            n_heads=self.heads, dim_head=self.dim_head, vit_window_size=self.window,
            mbconv_expansion_rate=self.expansion, mbconv_shrinkage_rate=self.shrink,
            num_register_tokens=self.num_reg,
        )

    def to_dict(self) -> dict:
        return asdict(self)


# BASELINE.json configs[0..2]: the repo's 12hr model (evaluation_vit.py:105-106)
CFG_12HR = GridConfig()
# a few-second CPU case with every structural feature (depth 2 => one residual MBConv)
CFG_TINY = GridConfig(T=2, C=24, H=26, W=25, dim=16, L=3, vit_depth=2, heads=2, dim_head=8)
# GPU-kernel-shaped but small: full channel/head geometry on a 26x25 domain
CFG_SMALL128 = GridConfig(T=3, C=24, H=26, W=25, dim=128, L=2)
# BASELINE configs[4] "MetNet-3-style at larger dim/heads" (n_start_channels 512, dim_head 64) on a small domain,
# and an intermediate width with depth 2
CFG_WIDE512 = GridConfig(T=2, C=24, H=26, W=25, dim=512, L=2, heads=32, dim_head=64)
CFG_WIDE256 = GridConfig(T=3, C=24, H=26, W=25, dim=256, L=2, vit_depth=2, heads=4, dim_head=64)
# MetNet3_with_stn_imgs (metnet3.py:518-759): 25 variables, the last one a station image
CFG_STN_SMALL128 = GridConfig(T=3, C=25, H=26, W=25, dim=128, L=2)


def maxvit_spec(dim: int, depth: int, cond_dim: int, heads: int, dim_head: int, window: int,
                expansion: float, shrink: float, num_reg: int, prefix: str = "") -> "OrderedDict[str, tuple]":
    """name -> (shape, kind) for a single-stage MaxViT (maxvit.py:262-287)."""
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    hidden = int(expansion * dim)
    se = int(hidden * shrink)
    inner = heads * dim_head
    for li in range(depth):
        mb = f"{prefix}layers.{li}.0." + ("" if li == 0 else "fn.")   # MBConvResidual wraps .fn
        spec[mb + "0.weight"] = ((hidden, dim, 1, 1), "w")
        spec[mb + "0.bias"] = ((hidden,), "b")
        for bn, ch in (("1", hidden), ("4", hidden), ("8", dim)):
            spec[mb + bn + ".weight"] = ((ch,), "pos")
            spec[mb + bn + ".bias"] = ((ch,), "b")
            spec[mb + bn + ".running_mean"] = ((ch,), "b")
            spec[mb + bn + ".running_var"] = ((ch,), "pos")
            spec[mb + bn + ".num_batches_tracked"] = ((), "count")
        spec[mb + "3.weight"] = ((hidden, 1, 3, 3), "w")
        spec[mb + "3.bias"] = ((hidden,), "b")
        spec[mb + "6.gate.1.weight"] = ((se, hidden), "w")
        spec[mb + "6.gate.3.weight"] = ((hidden, se), "w")
        spec[mb + "7.weight"] = ((dim, hidden, 1, 1), "w")
        spec[mb + "7.bias"] = ((dim,), "b")
        for ai in (1, 2):
            at = f"{prefix}layers.{li}.{ai}."
            spec[at + "film.0.weight"] = ((2 * dim, cond_dim), "w1")
            spec[at + "film.0.bias"] = ((2 * dim,), "b")
            spec[at + "film.2.weight"] = ((2 * dim, 2 * dim), "w")
            spec[at + "film.2.bias"] = ((2 * dim,), "film_b")
            spec[at + "to_qkv.weight"] = ((3 * inner, dim), "w")
            spec[at + "q_norm.gamma"] = ((heads, 1, dim_head), "pos")
            spec[at + "k_norm.gamma"] = ((heads, 1, dim_head), "pos")
            spec[at + "to_out.0.weight"] = ((dim, inner), "w")
            spec[at + "rel_pos_bias.weight"] = (((2 * window - 1) ** 2 + 1, heads), "w1")
        spec[f"{prefix}register_tokens.{li}"] = ((num_reg, dim), "w1")
    return spec


def maxvit_stage_dims(dim: int, depth) -> list:
    """(dim_in, dim_out, first_of_stage) of every layer, as maxvit.py:240-262 builds them.  int depth: one stage at constant
    width.  Tuple depth (d0, d1, ...): widths dim, 2 dim, 4 dim, ...; ``zip(dim_pairs, depth)`` pairs stage k = (2^k dim ->
    2^(k+1) dim) with depth d_k and silently drops the last entry (quirk Q9), so len(depth) - 1 stages exist."""
    if isinstance(depth, int):
        return [(dim, dim, i == 0) for i in range(depth)]
    dims = [(2 ** i) * dim for i in range(len(depth))]
    pairs = list(zip(dims[:-1], dims[1:])) if len(depth) > 1 else [(dim, dim)]
    out = []
    for (d_in, d_out), n in zip(pairs, depth):
        for i in range(n):
            out.append((d_in if i == 0 else d_out, d_out, i == 0))
    return out


def maxvit_multistage_spec(dim: int, depth, cond_dim: int, heads: int, dim_head: int, window: int, expansion: float,
                           shrink: float, num_reg: int) -> "OrderedDict[str, tuple]":
    """name -> (shape, kind) of a MaxViT built with a tuple ``depth`` (maxvit.py:240-287)"""
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    inner = heads * dim_head
    for li, (d_in, d, first) in enumerate(maxvit_stage_dims(dim, depth)):
        hidden = int(expansion * d)
        se = int(hidden * shrink)
        mb = f"layers.{li}.0." + ("" if (first or d_in != d) else "fn.")
        spec[mb + "0.weight"] = ((hidden, d_in, 1, 1), "w")
        spec[mb + "0.bias"] = ((hidden,), "b")
        for bn, ch in (("1", hidden), ("4", hidden), ("8", d)):
            spec[mb + bn + ".weight"] = ((ch,), "pos")
            spec[mb + bn + ".bias"] = ((ch,), "b")
            spec[mb + bn + ".running_mean"] = ((ch,), "b")
            spec[mb + bn + ".running_var"] = ((ch,), "pos")
            spec[mb + bn + ".num_batches_tracked"] = ((), "count")
        spec[mb + "3.weight"] = ((hidden, 1, 3, 3), "w")
        spec[mb + "3.bias"] = ((hidden,), "b")
        spec[mb + "6.gate.1.weight"] = ((se, hidden), "w")
        spec[mb + "6.gate.3.weight"] = ((hidden, se), "w")
        spec[mb + "7.weight"] = ((d, hidden, 1, 1), "w")
        spec[mb + "7.bias"] = ((d,), "b")
        for ai in (1, 2):
            at = f"layers.{li}.{ai}."
            spec[at + "film.0.weight"] = ((2 * d, cond_dim), "w1")
            spec[at + "film.0.bias"] = ((2 * d,), "b")
            spec[at + "film.2.weight"] = ((2 * d, 2 * d), "w")
            spec[at + "film.2.bias"] = ((2 * d,), "film_b")
            spec[at + "to_qkv.weight"] = ((3 * inner, d), "w")
            spec[at + "q_norm.gamma"] = ((heads, 1, dim_head), "pos")
            spec[at + "k_norm.gamma"] = ((heads, 1, dim_head), "pos")
            spec[at + "to_out.0.weight"] = ((d, inner), "w")
            spec[at + "rel_pos_bias.weight"] = (((2 * window - 1) ** 2 + 1, heads), "w1")
        spec[f"register_tokens.{li}"] = ((num_reg, d), "w1")
    return spec


def attention_spec(dim: int, cond_dim, heads: int, dim_head: int, window: int) -> "OrderedDict[str, tuple]":
    """stand-alone Attention (maxvit.py:106-168); cond_dim None: LayerNorm carries an affine and there is no FiLM MLP"""
    inner = heads * dim_head
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    if cond_dim is None:
        spec["norm.weight"] = ((dim,), "pos")
        spec["norm.bias"] = ((dim,), "b")
    else:
        spec["film.0.weight"] = ((2 * dim, cond_dim), "w1")
        spec["film.0.bias"] = ((2 * dim,), "b")
        spec["film.2.weight"] = ((2 * dim, 2 * dim), "w")
        spec["film.2.bias"] = ((2 * dim,), "film_b")
    spec["to_qkv.weight"] = ((3 * inner, dim), "w")
    spec["q_norm.gamma"] = ((heads, 1, dim_head), "pos")
    spec["k_norm.gamma"] = ((heads, 1, dim_head), "pos")
    spec["to_out.0.weight"] = ((dim, inner), "w")
    spec["rel_pos_bias.weight"] = (((2 * window - 1) ** 2 + 1, heads), "w1")
    return spec


def metnet3_spec(cfg: GridConfig) -> "OrderedDict[str, tuple]":
    """name -> (shape, kind) of every persistent tensor of the reference MetNet3."""
    d = cfg.dim
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    spec["pm25_boundaries"] = ((3,), "boundaries")
    spec["condition_lead_time.weight"] = ((cfg.L + 1, cfg.lead_emb), "w1")
    for i, n in enumerate((13, 32, 25)):
        spec[f"condition_model_time.{i}.weight"] = ((n, cfg.time_emb), "w1")
    for name, c0 in (("resnet1", cfg.c_in), ("resnet2", d)):
        cin = c0
        for bi in range(cfg.resnet_depth):
            p = f"{name}.blocks.{bi}."
            spec[p + "mlp.1.weight"] = ((2 * d, cfg.lead_emb), "w1")
            spec[p + "mlp.1.bias"] = ((2 * d,), "b")
            spec[p + "block1.proj.weight"] = ((d, cin, 3, 3), "w")
            spec[p + "block1.proj.bias"] = ((d,), "b")
            spec[p + "block1.norm.g"] = ((1, d, 1, 1), "pos")
            spec[p + "block1.norm.b"] = ((1, d, 1, 1), "b")
            spec[p + "block2.proj.weight"] = ((d, d, 3, 3), "w")
            spec[p + "block2.proj.bias"] = ((d,), "b")
            spec[p + "block2.norm.g"] = ((1, d, 1, 1), "pos")
            spec[p + "block2.norm.b"] = ((1, d, 1, 1), "b")
            if cin != d:
                spec[p + "res_conv.weight"] = ((d, cin, 1, 1), "w")
                spec[p + "res_conv.bias"] = ((d,), "b")
            cin = d
    spec.update(maxvit_spec(d, cfg.vit_depth, cfg.lead_emb, cfg.heads, cfg.dim_head, cfg.window,
                            cfg.expansion, cfg.shrink, cfg.num_reg, prefix="vit."))
    spec["up.weight"] = ((d, d, 2, 2), "w_up")
    spec["up.bias"] = ((d,), "b")
    spec["classifier_pm25.weight"] = ((1, d, 1, 1), "w")
    spec["classifier_pm25.bias"] = ((1,), "b")
    return spec


def _fill(name: str, shape: tuple, kind: str, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    if kind == "count":
        return torch.zeros((), dtype=torch.int64)
    if kind == "boundaries":
        return torch.tensor([15.0, 35.0, 75.0])
    if kind == "w":            # fan-in scaled so activations stay O(1)
        fan_in = max(1, math.prod(shape[1:]))
        return torch.randn(shape, generator=g) / math.sqrt(fan_in)
    if kind == "w_up":         # ConvTranspose2d (in, out, 2, 2): each output sees `in` taps
        return torch.randn(shape, generator=g) / math.sqrt(shape[0])
    if kind == "w1":
        return torch.randn(shape, generator=g)
    if kind == "b":
        return 0.1 * torch.randn(shape, generator=g)
    if kind == "pos":          # BN weight / running_var, LN g, q/k gamma: U(0.5, 1.5)
        return 0.5 + torch.rand(shape, generator=g)
    if kind == "film_b":       # first half is FiLM gamma (used raw, maxvit.py:187) -> centre on 1
        t = 0.1 * torch.randn(shape, generator=g)
        t[: shape[0] // 2] += 1.0
        return t
    raise KeyError(kind)


def make_state_dict(spec, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((k, _fill(k, shp, kind, seed)) for k, (shp, kind) in spec.items())


def make_inputs(cfg: GridConfig, B: int, seed: int = 1234, n_ts: int | None = None):
    """Synthetic CMAQ tensors: x (B,T,C,H,W) f32, timestamps (B,n_ts,4) f32,
    target (B,L,H,W) f32.  Species != PM2.5 are z-scored N(0,1); PM2.5 channels
    (c % 6 == 4, i.e. 4/10/16/22, metnet3.py:362) are raw ug/m3, lognormal(ln 18, 0.6)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cfg.T, cfg.C, cfg.H, cfg.W, generator=g)
    pm = torch.exp(math.log(18.0) + 0.6 * torch.randn(B, cfg.T, cfg.C // 6, cfg.H, cfg.W, generator=g))
    x[:, :, 4::6] = pm.clamp_(0.0, 300.0)
    n_ts = max(cfg.T, 7) if n_ts is None else n_ts        # metnet3.py:405 reads time index 6
    h0 = torch.randint(0, 24 * 360, (B, 1), generator=g) + torch.arange(n_ts)[None, :]
    ts = torch.stack([torch.full_like(h0, 2023), 1 + (h0 // 720) % 12, 1 + (h0 // 24) % 30, h0 % 24], dim=-1)
    target = torch.exp(math.log(18.0) + 0.6 * torch.randn(B, cfg.L, cfg.H, cfg.W, generator=g)).clamp_(0.0, 300.0)
    if cfg.C >= 25:
        # MetNet3_with_stn_imgs: channel 24 is the station-observation image in raw ug/m3 (metnet3.py:701); drawn last
        # so the streams of the 24-channel cases are unchanged
        x[:, :, 24] = torch.exp(math.log(18.0) + 0.6 * torch.randn(B, cfg.T, cfg.H, cfg.W, generator=g)).clamp_(0.0, 300.0)
    return x, ts.float(), target
