"""The oracle (oracle/) against the golden fixtures generated from the real reference."""
import pytest
import torch

from oracle import synth
from oracle.maxvit_oracle import (attention, block_pixel_index, grid_pixel_index, maxvit_forward,
                                  rel_pos_indices)
from oracle.metnet3_oracle import metnet3_forward
from oracle.focal_r_oracle import focal_r, focal_r_grad


def test_rel_pos_indices_bit_exact(golden):
    g = golden("index_golden.pt")
    for (w, r) in ((7, 4), (8, 1), (4, 2)):
        assert torch.equal(rel_pos_indices(w, r), g[f"rel_pos_w{w}_r{r}"])


@pytest.mark.parametrize("H,W,w", [(42, 35, 7), (14, 14, 7), (28, 21, 7), (259, 259, 7), (16, 24, 8)])
def test_partition_indices_bit_exact(golden, H, W, w):
    g = golden("index_golden.pt")
    assert torch.equal(block_pixel_index(H, W, w).int(), g[f"block_{H}x{W}_w{w}"])
    assert torch.equal(grid_pixel_index(H, W, w).int(), g[f"grid_{H}x{W}_w{w}"])


def test_attention_matches_reference(golden):
    f = golden("attention_small.pt")
    spec = {k[len("layers.0.1."):]: v for k, v in
            synth.maxvit_spec(f["dim"], 1, 2, f["heads"], f["dim_head"], f["window"], 4, 0.25, f["num_reg"]).items()
            if k.startswith("layers.0.1.")}
    sd = synth.make_state_dict(spec, seed=f["seed"])
    y = attention(f["x"], f["cond"], sd, "", heads=f["heads"], window=f["window"], num_reg=f["num_reg"])
    torch.testing.assert_close(y, f["y"], rtol=1e-5, atol=1e-5)


def test_maxvit_matches_reference(golden):
    f = golden("maxvit_small.pt")
    sd = synth.make_state_dict(synth.maxvit_spec(f["dim"], f["depth"], 2, f["heads"], f["dim_head"],
                                                 f["window"], 4, 0.25, f["num_reg"]), seed=f["seed"])
    y = maxvit_forward(f["x"], f["cond"], sd, depth=f["depth"], heads=f["heads"], window=f["window"],
                       num_reg=f["num_reg"])
    torch.testing.assert_close(y, f["y"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["metnet3_tiny.pt", "metnet3_small128.pt", "metnet3_12hr_b1.pt", "metnet3_wide256.pt",
                                  "metnet3_wide512.pt"])
def test_metnet3_matches_reference(golden, name):
    f = golden(name)
    cfg = synth.GridConfig(**f["cfg"])
    spec = synth.metnet3_spec(cfg)
    assert set(spec.keys()) == set(f["keys"])          # state-dict contract
    sd = synth.make_state_dict(spec, seed=f["weight_seed"])
    x, ts, _ = synth.make_inputs(cfg, f["B"], seed=f["input_seed"])
    y = metnet3_forward(x, ts, sd, cfg)
    rel = ((y - f["y"]).abs().max() / f["y"].abs().max()).item()
    assert rel < 1e-4, rel


def test_metnet3_stn_imgs_matches_reference(golden):
    """MetNet3_with_stn_imgs (metnet3.py:518-759): golden from the real reference class"""
    f = golden("metnet3_stn_small128.pt")
    cfg = synth.GridConfig(**f["cfg"])
    spec = synth.metnet3_spec(cfg)
    assert set(spec.keys()) == set(f["keys"]) and f["input_mutated"]
    sd = synth.make_state_dict(spec, seed=f["weight_seed"])
    x, ts, _ = synth.make_inputs(cfg, f["B"], seed=f["input_seed"])
    y = metnet3_forward(x, ts, sd, cfg, stn_imgs=True)
    rel = ((y - f["y"]).abs().max() / f["y"].abs().max()).item()
    assert rel < 1e-4, rel
    y_plain = metnet3_forward(x, ts, sd, cfg)                  # without the station-image normalisation: different
    assert ((y_plain - f["y"]).abs().max() / f["y"].abs().max()).item() > 1e-3


def test_12hr_param_count(golden):
    f = golden("metnet3_12hr_b1.pt")
    assert f["n_params"] == 3_346_145
    assert len(f["keys"]) == 93


def test_focal_r_properties():
    g = torch.Generator().manual_seed(0)
    p, t = torch.rand(4, 3, 9, 7, generator=g) * 50, torch.rand(4, 3, 9, 7, generator=g) * 50
    assert focal_r(p, p).item() == 0.0
    l1 = (p - t).abs().mean()
    assert 0 < focal_r(p, t) < l1                      # weight (2*sigmoid-1) in (0,1)
    # analytic gradient of |e|*(2s(b|e|)-1): sign(e)*[(2s-1) + 2 b |e| s (1-s)] / numel
    e = p - t
    s = torch.sigmoid(0.2 * e.abs())
    ga = torch.sign(e) * ((2 * s - 1) + 2 * 0.2 * e.abs() * s * (1 - s)) / e.numel()
    torch.testing.assert_close(focal_r_grad(p, t), ga, rtol=1e-5, atol=1e-8)


def test_sample_chunking_is_exact():
    """the oracle evaluated in sample chunks (full-batch time embedding, quirk Q1) == the oracle on the whole batch"""
    cfg = synth.CFG_TINY
    sd = synth.make_state_dict(synth.metnet3_spec(cfg), seed=1)
    x, ts, _ = synth.make_inputs(cfg, 5, seed=8)
    y = metnet3_forward(x, ts, sd, cfg)
    y2 = metnet3_forward(x, ts, sd, cfg, sample_chunk=2)
    torch.testing.assert_close(y2, y, rtol=1e-6, atol=1e-6)
    # and the scramble really couples samples: chunking the INPUT (timestamps included) changes the answer
    y3 = torch.cat([metnet3_forward(x[b:b + 1], ts[b:b + 1], sd, cfg) for b in range(5)])
    assert (y3 - y).abs().max() > 1e-4


def test_multistage_maxvit_matches_reference(golden):
    """tuple depth (maxvit.py:240-262, quirk Q9): golden from the real reference"""
    f = golden("maxvit_multistage.pt")
    spec = synth.maxvit_multistage_spec(f["dim"], f["depth"], 2, f["heads"], f["dim_head"], f["window"], 4, 0.25, f["num_reg"])
    assert list(spec.keys()) == [k for k in f["keys"]] or set(spec.keys()) == set(f["keys"])
    sd = synth.make_state_dict(spec, seed=f["seed"])
    n_layers = len(synth.maxvit_stage_dims(f["dim"], f["depth"]))
    assert n_layers == 2                                  # depth (2, 1): the trailing entry is dropped by zip()
    y = maxvit_forward(f["x"], f["cond"], sd, depth=n_layers, heads=f["heads"], window=f["window"], num_reg=f["num_reg"])
    assert y.shape == f["y"].shape == (2, 128, 14, 21)
    torch.testing.assert_close(y, f["y"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("variant,cond_dim", [("film", 2), ("nocond", None)])
def test_standalone_attention_matches_reference(golden, variant, cond_dim):
    f = golden("attention_standalone.pt")
    sd = synth.make_state_dict(synth.attention_spec(f["dim"], cond_dim, f["heads"], f["dim_head"], f["window"]), seed=f["seed"])
    y = attention(f["x"], f["cond"], sd, "", heads=f["heads"], window=f["window"], num_reg=f["num_reg"])
    torch.testing.assert_close(y, f[variant], rtol=1e-4, atol=1e-4)
