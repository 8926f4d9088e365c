"""Round-2 parity gates (GPU): the cases VERDICT r01 asked to be held by tests instead of prose.

 * the window / grid partition map evaluated ON THE DEVICE, bit-exact against the einops-generated golden tables
   (42x35, 259x259 = the 512x512 domain of BASELINE configs[3], 16x24 / w8)
 * BASELINE configs[1] at its full size (B=64 -> 768 fields), bf16 mode vs the CPU oracle, max rel err < 1e-2
 * the 512x512 geometry (X = Y = 37 windows per side, maxvit.py:322) through the whole network, depth 2, fp32 + bf16
 * BASELINE configs[4] at full depth (512 channels, 32 x 64 heads, depth 4): the default precision meets 1e-2
 * packed bf16 host batches (half the H2D bytes) give bit-identical predictions
 * an out-of-range timestamp raises like the reference's nn.Embedding (IndexError) instead of reading out of bounds
"""
import pytest
import torch

from oracle import synth
from oracle.metnet3_oracle import metnet3_forward

pytestmark = pytest.mark.gpu


def build(cfg, seed, precision, cls=None):
    """precision None: the constructor's default for this width"""
    from vit_grid_model_b200 import MetNet3
    m = (cls or MetNet3)(**cfg.metnet3_kwargs())
    m.load_state_dict(synth.make_state_dict(synth.metnet3_spec(cfg), seed=seed), strict=True)
    m = m.cuda().eval()
    return m if precision is None else m.set_precision(precision)


def rel_err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


# ------------------------------------------------------------------------------------------ index work: bit-exact
@pytest.mark.parametrize("key", ["42x35_w7", "14x14_w7", "28x21_w7", "259x259_w7", "16x24_w8"])
@pytest.mark.parametrize("mode", ["block", "grid"])
def test_partition_map_on_device_bit_exact(golden, key, mode):
    """vg_attn_partition_debug runs attn_token_pixel -- the one function every gather / scatter kernel addresses tokens
    through -- on the device for every token row; golden = einops rearrange of an arange (maxvit.py:298 / :322)"""
    from vit_grid_model_b200 import _lib
    ref = golden("index_golden.pt")[f"{mode}_{key}"]                # (nwin, win*win) pixel index inside one field
    hw, w = key.split("_w")
    H, W = (int(t) for t in hw.split("x"))
    win, R, N = int(w), 4, 3
    nwin, S = ref.shape[0], R + win * win
    out = torch.full((N * nwin * S,), -7, dtype=torch.int64, device="cuda")
    _lib.call("vg_attn_partition_debug", N, H, W, win, R, int(mode == "grid"), out.data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    out = out.cpu().view(N, nwin, S)
    assert (out[:, :, :R] == -1).all()
    want = ref.to(torch.int64)[None] + (torch.arange(N, dtype=torch.int64) * H * W)[:, None, None]
    assert torch.equal(out[:, :, R:], want)


@pytest.mark.parametrize("mode", ["block", "grid"])
@pytest.mark.parametrize("hw", [(42, 35), (259, 259)])
def test_fused_attention_moves_rows_through_the_golden_map(golden, hw, mode):
    """coverage through the production kernel: with zero to_out weights the fused attention must return its residual input
    unchanged at every pixel (every pixel is read and written exactly once, at its own address) and hand the register
    tokens through -- bit-exact.  259 x 259 is the 512 x 512 domain (X = Y = 37)."""
    from vit_grid_model_b200 import ops
    H, W = hw
    N, C, heads, dh = 2, 128, 32, 32
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, H, W, C, generator=g).cuda()
    reg = torch.randn(4, C, generator=g).cuda()
    film = torch.cat([torch.ones(N, C), torch.zeros(N, C)], 1).cuda()
    wqkv = (torch.randn(heads * 3 * dh, C, generator=g) / C ** 0.5).half().cuda()
    wout = torch.zeros(heads, C, dh).cuda()
    tab = ops.pack_head_tables(torch.zeros(170, heads).cuda(), torch.ones(heads * dh).cuda(), torch.ones(heads * dh).cuda())
    y, reg_out = ops.attn_fused(x, reg, film, wqkv, wout, tab, 7, 4, mode == "grid", True, heads, dh)
    assert torch.equal(y, x)
    assert torch.equal(reg_out, reg[None].expand(reg_out.shape[0], 4, C))


# ------------------------------------------------------------------------------------------ full-size configurations
@pytest.mark.slow
def test_config1_b64_bf16_vs_cpu_oracle():
    """BASELINE configs[1] exactly: B=64 (768 fields), bf16 mode, against the CPU oracle on all 4.2 M outputs.
    north_star: max rel err (max|pred - ref| / max|ref|) < 1e-2."""
    cfg = synth.CFG_12HR
    m = build(cfg, 0, "bf16")
    x, ts, _ = synth.make_inputs(cfg, 64, seed=5)
    with torch.no_grad():
        y = m(x.cuda(), timestamps=ts.cuda()).cpu()
        ref = metnet3_forward(x, ts, synth.make_state_dict(synth.metnet3_spec(cfg), seed=0), cfg, sample_chunk=4)
    assert y.shape == ref.shape == (64, 12, 82, 67)
    e = rel_err(y, ref)
    assert e < 1e-2, e
    # secondary statement, per predicted grid: relative L2 error of every one of the 768 grids
    per_grid = ((y - ref).flatten(2).norm(dim=2) / ref.flatten(2).norm(dim=2)).max().item()
    assert per_grid < 1e-2, per_grid
    print(f"B=64 bf16 vs CPU oracle: max rel err {e:.3e}, worst per-grid relative L2 {per_grid:.3e}")


@pytest.mark.slow
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_config3_geometry_512_vs_cpu_oracle(precision, tol):
    """BASELINE configs[3] geometry: 512 x 512 domain (padded 518 x 518, 259 x 259 at half resolution, X = Y = 37 windows per
    side: the grid partition gathers with stride 37, maxvit.py:322), MaxViT depth 2, one field, vs the CPU oracle.
    T = 7 time steps (the minimum the hard-coded time index 6 allows) keeps the oracle at ~20 s."""
    cfg = synth.GridConfig(T=7, C=24, H=512, W=512, dim=128, L=1, vit_depth=2)
    m = build(cfg, 4, precision)
    x, ts, _ = synth.make_inputs(cfg, 1, seed=17)
    with torch.no_grad():
        y = m(x.cuda(), timestamps=ts.cuda()).cpu()
        ref = metnet3_forward(x, ts, synth.make_state_dict(synth.metnet3_spec(cfg), seed=4), cfg)
    assert y.shape == ref.shape == (1, 1, 512, 512)
    e = rel_err(y, ref)
    assert e < tol, e


@pytest.mark.slow
def test_config4_full_depth_default_precision_vs_cpu_oracle():
    """BASELINE configs[4]: 512 channels, 32 heads x dim_head 64, MaxViT depth 4, 82 x 67 domain, one sample (12 fields).
    The DEFAULT precision of a wide network must meet the north-star 1e-2 against the CPU oracle.  Measured on B200
    (tools/config5_parity.py): bf16 3.6e-2, tf32 everywhere 1.4e-2, tf32 + exact-fp32 QKV projection 1.07e-2, tf32 convolutions +
    exact-fp32 SIMT MaxViT 4.3e-3, tf32 convolutions + 3xTF32 MaxViT projections and attention core (the default, 'tf32_conv') 4.4e-3,
    exact fp32 2.2e-5 -- four stacked MaxViT layers with un-scaled +-32 gamma^2 logits amplify the
    operand rounding of every projection of the block.  Wide networks therefore default to 'tf32_conv'; the faster
    reduced-precision modes are opt-in there.  The exact-fp32 mode is held to the fp32 tolerance on the same case."""
    cfg = synth.GridConfig(dim=512, heads=32, dim_head=64, vit_depth=4)
    m = build(cfg, 0, None)
    assert m.precision == "tf32_conv"
    x, ts, _ = synth.make_inputs(cfg, 1, seed=3)
    with torch.no_grad():
        y = m(x.cuda(), timestamps=ts.cuda()).cpu()
        ref = metnet3_forward(x, ts, synth.make_state_dict(synth.metnet3_spec(cfg), seed=0), cfg)
        y32 = m.set_precision("fp32")(x.cuda(), timestamps=ts.cuda()).cpu()
    e = rel_err(y, ref)
    assert e < 7e-3, e                                     # north_star bf16-mode tolerance 1e-2, with margin
    assert rel_err(y32, ref) < 1e-4
    assert build(synth.CFG_12HR, 0, None).precision == "bf16"


@pytest.mark.parametrize("precision", ["tf32", "tf32_conv"])
@pytest.mark.parametrize("name", ["metnet3_small128.pt", "metnet3_12hr_b1.pt", "metnet3_wide256.pt", "metnet3_wide512.pt"])
def test_tf32_mode_golden_from_reference(golden, name, precision):
    """set_precision('tf32'): fp32 storage, every contraction on tcgen05 kind::tf32 -- 3e-3 on the goldens (bf16 mode: 7e-3);
    'tf32_conv': tf32 convolutions, the MaxViT block in fp32 with 3xTF32 split projections and attention core (the default of wide
    networks, here also on the 128-channel goldens)"""
    f = golden(name)
    cfg = synth.GridConfig(**f["cfg"])
    m = build(cfg, f["weight_seed"], precision)
    x, ts, _ = synth.make_inputs(cfg, f["B"], seed=f["input_seed"])
    with torch.no_grad():
        y = m(x.cuda(), timestamps=ts.cuda())
    assert rel_err(y, f["y"]) < 6e-3


# ------------------------------------------------------------------------------------------ callers either side
def test_packed_host_batches_bit_identical():
    """pipeline.pack_host (bf16, PM2.5 channels standardised before the rounding) -> HostPipeline: half the host->device
    bytes, the same predictions bit for bit as the fp32 tensor"""
    from vit_grid_model_b200 import HostPipeline, pack_host
    cfg = synth.CFG_SMALL128
    m = build(cfg, 2, "bf16")
    batches32, batches16 = [], []
    for k in range(3):
        x, ts, _ = synth.make_inputs(cfg, 2 + (k % 2), seed=200 + k)
        batches32.append((x.pin_memory(), ts.pin_memory()))
        xp = pack_host(x, cfg.pm25_mean, cfg.pm25_std)
        assert xp.dtype == torch.bfloat16 and xp.is_pinned() and xp.numel() * 2 * 2 == x.numel() * 4
        batches16.append((xp, ts.pin_memory()))
    a = [y.clone() for y in HostPipeline(m).run(batches32)]
    b = [y.clone() for y in HostPipeline(m).run(batches16)]
    assert len(a) == len(b) == 3
    for ya, yb in zip(a, b):
        assert torch.equal(ya, yb)
    with pytest.raises(NotImplementedError):
        m.set_precision("fp32").forward_packed(batches16[0][0].cuda(), batches16[0][1].cuda())


def test_out_of_range_timestamp_raises_like_embedding():
    """month 13 / day 32 / hour 25 / -1: the reference's nn.Embedding raises IndexError (metnet3.py:392); host timestamps are
    checked before the launch, device timestamps poison the field with NaN and raise at the next call"""
    from vit_grid_model_b200 import _lib
    cfg = synth.CFG_SMALL128
    m = build(cfg, 0, "bf16")
    x, ts, _ = synth.make_inputs(cfg, 2, seed=9)
    bad = ts.clone()
    bad[1, 6, 1] = 13.0
    with pytest.raises(IndexError):
        m(x.cuda(), timestamps=bad)                       # host tensor: validated up front
    with torch.no_grad():
        y = m(x.cuda(), timestamps=bad.cuda())            # device tensor: no host sync, the kernel guards the lookup
    torch.cuda.synchronize()
    assert torch.isnan(y).any()
    assert _lib.load().vg_device_error(0) & 1
    with pytest.raises(IndexError):
        m(x.cuda(), timestamps=ts.cuda())                 # surfaced (and cleared) at the next call
    with torch.no_grad():
        y2 = m(x.cuda(), timestamps=ts.cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(y2).all() and _lib.load().vg_device_error(0) == 0
    neg = ts.clone()
    neg[0, 6, 3] = -1.0
    with pytest.raises(IndexError):
        m(x.cuda(), timestamps=neg)
