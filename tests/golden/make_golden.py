"""Generate the golden fixtures from the REAL reference (/root/reference/src).

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own ``maxvit`` / ``metnet3`` modules (with the two
shims SURVEY.md §8c describes: a stub ``ipdb`` module, and a no-op
``Tensor.cuda`` because metnet3.py:356-359 hard-codes ``.cuda()``), loads the
name-seeded synthetic weights of ``oracle.synth`` with ``strict=True`` (which
also proves the state-dict contract), runs them on the seeded synthetic inputs
and stores the reference's outputs.  Index fixtures are produced with the
reference's own einops expressions on ``arange`` tensors.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
torch.Tensor.cuda = lambda self, *a, **k: self          # CPU-only container

import maxvit as ref_maxvit            # noqa: E402  (the reference)
import metnet3 as ref_metnet3          # noqa: E402
from einops import rearrange           # noqa: E402

from oracle import synth               # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)


def save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def index_fixtures():
    out = {}
    for (w, r) in ((7, 4), (8, 1), (4, 2)):
        att = ref_maxvit.Attention(dim=32, cond_dim=2, heads=2, dim_head=16, window_size=w, num_registers=r)
        out[f"rel_pos_w{w}_r{r}"] = att.rel_pos_indices.clone()
    for (H, W, w) in ((42, 35, 7), (14, 14, 7), (28, 21, 7), (259, 259, 7), (16, 24, 8)):
        pix = torch.arange(H * W).reshape(1, 1, H, W)
        blk = rearrange(pix, 'b d (x w1) (y w2) -> b x y w1 w2 d', w1=w, w2=w)      # maxvit.py:298
        grd = rearrange(pix, 'b d (w1 x) (w2 y) -> b x y w1 w2 d', w1=w, w2=w)      # maxvit.py:322
        out[f"block_{H}x{W}_w{w}"] = blk.reshape(-1, w * w).to(torch.int32)
        out[f"grid_{H}x{W}_w{w}"] = grd.reshape(-1, w * w).to(torch.int32)
    save("index_golden.pt", out)


def attention_fixture():
    dim, heads, dh, w, r, N, nwin = 32, 4, 8, 7, 4, 2, 3
    att = ref_maxvit.Attention(dim=dim, cond_dim=2, heads=heads, dim_head=dh, dropout=0.1,
                               window_size=w, num_registers=r).eval()
    spec = {k[len("layers.0.1."):]: v for k, v in
            synth.maxvit_spec(dim, 1, 2, heads, dh, w, 4, 0.25, r).items() if k.startswith("layers.0.1.")}
    sd = synth.make_state_dict(spec, seed=3)
    att.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N * nwin, r + w * w, dim, generator=g)
    cond = torch.randn(N, 2, generator=g)
    with torch.no_grad():
        y = att(x, cond)
    save("attention_small.pt", dict(dim=dim, heads=heads, dim_head=dh, window=w, num_reg=r, seed=3,
                                    x=x, cond=cond, y=y))


def maxvit_fixture():
    dim, depth, heads, dh, w, r, N, H, W = 16, 2, 2, 8, 7, 4, 3, 14, 21
    m = ref_maxvit.MaxViT(dim=dim, depth=depth, cond_dim=2, heads=heads, dim_head=dh, vit_window_size=w,
                          num_register_tokens=r).eval()
    sd = synth.make_state_dict(synth.maxvit_spec(dim, depth, 2, heads, dh, w, 4, 0.25, r), seed=5)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(N, dim, H, W, generator=g)
    cond = torch.randn(N, 2, generator=g)
    with torch.no_grad():
        y = m(x, cond)
    save("maxvit_small.pt", dict(dim=dim, depth=depth, heads=heads, dim_head=dh, window=w, num_reg=r, seed=5,
                                 x=x, cond=cond, y=y))


def maxvit_multistage_fixture():
    """tuple depth (maxvit.py:240-262): depth (2, 1) from width 64 builds ONE stage 64 -> 128 of two layers (quirk Q9)"""
    dim, depth, heads, dh, w, r, N, H, W = 64, (2, 1), 4, 32, 7, 4, 2, 14, 21
    m = ref_maxvit.MaxViT(dim=dim, depth=depth, cond_dim=2, heads=heads, dim_head=dh, vit_window_size=w,
                          num_register_tokens=r).eval()
    sd = synth.make_state_dict(synth.maxvit_multistage_spec(dim, depth, 2, heads, dh, w, 4, 0.25, r), seed=6)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(13)
    x = torch.randn(N, dim, H, W, generator=g)
    cond = torch.randn(N, 2, generator=g)
    with torch.no_grad():
        y = m(x, cond)
    assert y.shape == (N, 2 * dim, H, W)
    save("maxvit_multistage.pt", dict(dim=dim, depth=depth, heads=heads, dim_head=dh, window=w, num_reg=r, seed=6,
                                      x=x, cond=cond, y=y, keys=list(m.state_dict().keys())))


def attention_standalone_fixture():
    """stand-alone Attention.forward (maxvit.py:170-219) at a kernel-sized width, with FiLM conditioning and with
    cond_dim=None (LayerNorm affine, no FiLM: maxvit.py:128-137)"""
    dim, heads, dh, w, r, N, nwin = 128, 4, 32, 7, 4, 2, 3
    out = dict(dim=dim, heads=heads, dim_head=dh, window=w, num_reg=r)
    g = torch.Generator().manual_seed(14)
    x = torch.randn(N * nwin, r + w * w, dim, generator=g)
    cond = torch.randn(N, 2, generator=g)
    for name, cd in (("film", 2), ("nocond", None)):
        att = ref_maxvit.Attention(dim=dim, cond_dim=cd, heads=heads, dim_head=dh, dropout=0.1, window_size=w, num_registers=r).eval()
        sd = synth.make_state_dict(synth.attention_spec(dim, cd, heads, dh, w), seed=8)
        att.load_state_dict(sd, strict=True)
        with torch.no_grad():
            out[name] = att(x, cond)          # cond is required positionally even without FiLM (maxvit.py:172,177)
    out.update(x=x, cond=cond, seed=8)
    save("attention_standalone.pt", out)


def metnet3_fixture(name, cfg, B, wseed, iseed):
    m = ref_metnet3.MetNet3(**cfg.metnet3_kwargs()).eval()
    sd = synth.make_state_dict(synth.metnet3_spec(cfg), seed=wseed)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    x, ts, _ = synth.make_inputs(cfg, B, seed=iseed)
    with torch.no_grad():
        y = m(x, timestamps=ts)
    assert y.shape == (B, cfg.L, cfg.H, cfg.W)
    n_params = sum(p.numel() for p in m.parameters())
    save(name, dict(cfg=cfg.to_dict(), B=B, weight_seed=wseed, input_seed=iseed, y=y.contiguous(),
                    n_params=n_params, keys=list(m.state_dict().keys())))


def metnet3_stn_fixture(name, cfg, B, wseed, iseed):
    """MetNet3_with_stn_imgs (metnet3.py:518-759).  Also records that the reference normalises channel 24 of the CALLER's
    tensor in place (:701 runs before the clone at :702)."""
    m = ref_metnet3.MetNet3_with_stn_imgs(**cfg.metnet3_kwargs()).eval()
    sd = synth.make_state_dict(synth.metnet3_spec(cfg), seed=wseed)
    m.load_state_dict(sd, strict=True)
    x, ts, _ = synth.make_inputs(cfg, B, seed=iseed)
    x0 = x.clone()
    with torch.no_grad():
        y = m(x, timestamps=ts)
    mutated = torch.allclose(x[:, :, 24], (x0[:, :, 24] - cfg.pm25_mean) / cfg.pm25_std) and torch.equal(x[:, :, :24], x0[:, :, :24])
    assert mutated
    save(name, dict(cfg=cfg.to_dict(), B=B, weight_seed=wseed, input_seed=iseed, y=y.contiguous(), input_mutated=mutated,
                    keys=list(m.state_dict().keys())))


def sample_index(numel, k=48, seed=0):
    g = torch.Generator().manual_seed(seed + numel)
    return torch.randperm(numel, generator=g)[:min(k, numel)].clone()


def metnet3_train_fixture(name, cfg, B, wseed, iseed):
    """train() mode forward + Focal-R + backward through the REAL reference (dropout 0: its RNG stream cannot be matched).
    Stores the loss, the prediction, and per parameter the gradient norm plus 48 sampled entries; BatchNorm running
    statistics after the step."""
    from oracle.focal_r_oracle import focal_r
    m = ref_metnet3.MetNet3(**cfg.metnet3_kwargs(), dropout=0.0).train()
    sd = synth.make_state_dict(synth.metnet3_spec(cfg), seed=wseed)
    m.load_state_dict(sd, strict=True)
    x, ts, target = synth.make_inputs(cfg, B, seed=iseed)
    pred = m(x, timestamps=ts)
    loss = focal_r(pred, target)
    loss.backward()
    grads = {}
    for k, p in m.named_parameters():
        g = p.grad.detach().reshape(-1)
        idx = sample_index(g.numel())
        grads[k] = dict(norm=g.norm().item(), absmax=g.abs().max().item(), idx=idx, val=g[idx].clone())
    bn = {k: v.clone() for k, v in m.state_dict().items() if "running_" in k or "num_batches" in k}
    save(name, dict(cfg=cfg.to_dict(), B=B, weight_seed=wseed, input_seed=iseed, loss=loss.item(), pred=pred.detach().contiguous(),
                    grads=grads, bn=bn))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "r2":      # the fixtures added in round 2 only
        maxvit_multistage_fixture()
        attention_standalone_fixture()
        sys.exit(0)
    maxvit_multistage_fixture()
    attention_standalone_fixture()
    index_fixtures()
    attention_fixture()
    maxvit_fixture()
    metnet3_fixture("metnet3_tiny.pt", synth.CFG_TINY, 2, 0, 1234)
    metnet3_fixture("metnet3_small128.pt", synth.CFG_SMALL128, 2, 0, 1234)
    metnet3_fixture("metnet3_12hr_b1.pt", synth.CFG_12HR, 1, 0, 1234)
    metnet3_train_fixture("metnet3_small128_train.pt", synth.CFG_SMALL128, 3, 0, 4321)
    metnet3_stn_fixture("metnet3_stn_small128.pt", synth.CFG_STN_SMALL128, 2, 0, 1234)
    metnet3_fixture("metnet3_wide256.pt", synth.CFG_WIDE256, 2, 0, 1234)
    metnet3_fixture("metnet3_wide512.pt", synth.CFG_WIDE512, 2, 0, 1234)
