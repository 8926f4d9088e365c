"""Functional fp32 restatement of MetNet3 (/root/reference/src/metnet3.py:86-430).

Test infrastructure only (see oracle/__init__.py).  Also used as the timed CPU
baseline ("port") by bench.py, because the Python reference itself cannot travel
to the GPU box.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .maxvit_oracle import maxvit_forward
from .synth import GridConfig


def chan_layer_norm(x, g, b, eps: float = 1e-5):
    """metnet3.py:94-104: biased variance, clamp(min=eps).rsqrt() (NOT var+eps)."""
    var = x.var(dim=1, unbiased=False, keepdim=True)
    mean = x.mean(dim=1, keepdim=True)
    return (x - mean) * var.clamp(min=eps).rsqrt() * g + b


def resnet_block(x, cond, sd, p: str):
    """metnet3.py:129-162.  FiLM (scale+1, shift) only in block1; cond MLP is ReLU->Linear."""
    ss = F.linear(F.relu(cond), sd[p + "mlp.1.weight"], sd[p + "mlp.1.bias"])
    scale, shift = ss.chunk(2, dim=1)
    h = F.conv2d(x, sd[p + "block1.proj.weight"], sd[p + "block1.proj.bias"], padding=1)
    h = chan_layer_norm(h, sd[p + "block1.norm.g"], sd[p + "block1.norm.b"])
    h = F.relu(h * (scale[:, :, None, None] + 1) + shift[:, :, None, None])
    h = F.conv2d(h, sd[p + "block2.proj.weight"], sd[p + "block2.proj.bias"], padding=1)
    h = F.relu(chan_layer_norm(h, sd[p + "block2.norm.g"], sd[p + "block2.norm.b"]))
    if (p + "res_conv.weight") in sd:
        x = F.conv2d(x, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"])
    return h + x


def time_embedding(timestamps, sd, cfg: GridConfig):
    """metnet3.py:389-408 for the WHOLE batch -> (t_emb (N, lead_emb + 3*time_emb), cond (N, lead_emb)), N = B*L.
    The scramble of quirk Q1 couples every field to the timestamps of other samples, so it is always computed on the
    full batch, even when the rest of the network is evaluated in sample chunks."""
    B, L = timestamps.shape[0], cfg.L
    N = B * L
    ts = timestamps[:, 6, :].repeat_interleave(L, dim=0)
    lead = torch.arange(1, L + 1).repeat(B)
    cond = sd["condition_lead_time.weight"][lead]
    mt = ts[:, 1:4].int().long()                                   # month, day, hour
    flat = torch.cat([sd[f"condition_model_time.{i}.weight"][mt[:, i]] for i in range(3)], dim=0)
    return torch.cat([cond, flat.reshape(N, -1)], dim=1), cond


def prepare_input(x, timestamps, sd, cfg: GridConfig, stn_imgs: bool = False, samples=None):
    """metnet3.py:354-416 -> (net_in (N,c_in,HP,WP), cond (N,lead_emb)), N = B*L.

    * PM2.5 channels {4,10,16,22} of every time step standardised (:361-380)
    * every sample replicated L times, lead-minor: n = b*L + l (:383, :407)
    * zero pad to multiples of 14 (:384)
    * time channels: lead embedding (lead_emb) ++ the *scrambled* model-time embedding:
      the three (N,1) embeddings are concatenated on dim 0 and viewed as (N,3), so row n
      holds flat[3n:3n+3] of [month_0..month_{N-1}, day_0.., hour_0..] (:395-401, quirk Q1)
    * timestamps are taken at hard-coded time index 6 (:405, quirk Q2)
    samples = (b0, b1): only the fields of samples [b0, b1) (time embedding still from the whole batch).
    """
    L = cfg.L
    t_emb, cond = time_embedding(timestamps, sd, cfg)
    if samples is not None:
        b0, b1 = samples
        x, t_emb, cond = x[b0:b1], t_emb[b0 * L:b1 * L], cond[b0 * L:b1 * L]
    B = x.shape[0]
    x = x.clone()
    pm = torch.tensor([4, 10, 16, 22] + ([24] if stn_imgs else []))   # MetNet3_with_stn_imgs also standardises the
    x[:, :, pm] = (x[:, :, pm] - cfg.pm25_mean) / cfg.pm25_std          # station image, channel 24 (metnet3.py:701)
    x = x.repeat_interleave(L, dim=0)
    x = F.pad(x, cfg.pads, value=0.0)
    N, HP, WP = B * L, x.shape[-2], x.shape[-1]
    x = x.reshape(N, -1, HP, WP)
    x = torch.cat([x, t_emb[:, :, None, None].expand(-1, -1, HP, WP)], dim=1)
    return x, cond


def metnet3_forward(x, timestamps, sd, cfg: GridConfig, *, training: bool = False, return_features: bool = False,
                    stn_imgs: bool = False, sample_chunk: int | None = None):
    """MetNet3.forward (metnet3.py:339-430): (B,T,C,H,W), (B,*,4) -> (B,L,H,W) fp32.
    stn_imgs=True: MetNet3_with_stn_imgs.forward (metnet3.py:666-759), identical but for the extra normalised channel
    (the reference ALSO writes that normalisation back into the caller's tensor, :701; the oracle is functional).
    sample_chunk (eval mode only): evaluate the network `sample_chunk` samples at a time to bound host memory at large B --
    nothing after the time embedding couples the fields of a batch in eval mode, so the result is the same
    (tests/test_oracle.py::test_sample_chunking_is_exact)."""
    B = x.shape[0]
    if sample_chunk is not None and not training and not return_features and sample_chunk < B:
        return torch.cat([_forward_samples(x, timestamps, sd, cfg, (b0, min(B, b0 + sample_chunk)), False, False, stn_imgs)
                          for b0 in range(0, B, sample_chunk)], dim=0)
    return _forward_samples(x, timestamps, sd, cfg, None, training, return_features, stn_imgs)


def _forward_samples(x, timestamps, sd, cfg, samples, training, return_features, stn_imgs):
    h, cond = prepare_input(x, timestamps, sd, cfg, stn_imgs, samples)
    B = h.shape[0] // cfg.L
    for bi in range(cfg.resnet_depth):
        h = resnet_block(h, cond, sd, f"resnet1.blocks.{bi}.")
    feats = {"resnet1": h}
    h = F.max_pool2d(h, 2, 2)
    h = maxvit_forward(h, cond, sd, prefix="vit.", depth=cfg.vit_depth, heads=cfg.heads,
                       window=cfg.window, num_reg=cfg.num_reg, training=training)
    feats["vit"] = h
    h = F.conv_transpose2d(h, sd["up.weight"], sd["up.bias"], stride=2)
    feats["up"] = h
    for bi in range(cfg.resnet_depth):
        h = resnet_block(h, cond, sd, f"resnet2.blocks.{bi}.")
    feats["resnet2"] = h
    pl, pr, pt, pb = cfg.pads
    h = h[..., pt:h.shape[-2] - pb, pl:h.shape[-1] - pr]           # unpad (:335-337), pads > 0 here
    out = F.conv2d(h, sd["classifier_pm25.weight"], sd["classifier_pm25.bias"])
    out = out.squeeze(1).reshape(B, cfg.L, cfg.H, cfg.W) * cfg.pm25_std + cfg.pm25_mean
    return (out, feats) if return_features else out
