"""Diagnostics: stage-by-stage difference of the train-mode forward between fp32 and mixed precision (GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth                                    # noqa: E402
from vit_grid_model_b200 import MetNet3                     # noqa: E402
from vit_grid_model_b200.train import metnet3_train_forward  # noqa: E402

cfg = synth.CFG_SMALL128
B = 3
x, ts, target = synth.make_inputs(cfg, B, seed=4321)
out = {}
for precision in ("fp32", "bf16"):
    m = MetNet3(**cfg.metnet3_kwargs(), dropout=0.0)
    m.load_state_dict(synth.make_state_dict(synth.metnet3_spec(cfg), seed=0), strict=True)
    m = m.cuda().train().set_precision(precision)
    with torch.no_grad():
        pred, S = metnet3_train_forward(m, x.cuda(), ts.cuda().float())
    v = S["vit"][0]
    out[precision] = dict(h1=S["stem"]["h1"], h_enc=S["h_enc"], mb_h0=v["h0"], mb_h1=v["h1"], mb_h2=v["h2"], mb_h3=v["h3"], mb_h4=v["h4"],
                          mb_y0=v["y0"], battn_in=v["battn"]["x"], battn_tok=v["battn"]["tokens"], battn_qkv=v["battn"]["qkv"],
                          battn_att=v["battn"]["att"], gattn_in=v["gattn"]["x"], low_out=S["low_out"], dec0_in=S["dec"][0]["x"],
                          dec0_t1=S["dec"][0]["t1"], dec1_in=S["dec"][1]["x"], h_last=S["h_last"], pred=pred)
for k in out["fp32"]:
    a, b = out["fp32"][k].float(), out["bf16"][k].float()
    d = (a - b)
    print(f"{k:12s} rms rel {d.pow(2).mean().sqrt().item() / a.pow(2).mean().sqrt().item():.3e}   max rel {d.abs().max().item() / a.abs().max().item():.3e}"
          f"   mean diff {d.mean().item():+.3e} (mean {a.mean().item():+.3e})")
