"""Two train-mode MaxViT forwards on the same input: which saved tensor differs first? (race / uninitialised-read detector)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from vit_grid_model_b200 import MaxViT
from vit_grid_model_b200 import train as tr

dim, depth, heads, dh, w, r, N, H, W, p = 128, 1, 32, 32, 7, 4, 2, 14, 21, float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
sd = synth.make_state_dict(synth.maxvit_spec(dim, depth, 2, heads, dh, w, 4, 0.25, r), seed=7)
m = MaxViT(dim=dim, depth=depth, cond_dim=2, heads=heads, dim_head=dh, vit_window_size=w, num_register_tokens=r, dropout=p)
m.load_state_dict(sd, strict=True)
m = m.cuda().train().set_precision("bf16")
g = torch.Generator().manual_seed(1)
x = torch.randn(N, H, W, dim, generator=g).cuda()
cond = torch.randn(N, 2, generator=g).cuda()
outs = []
with torch.no_grad():
    for k in range(3):
        y, saved = tr.maxvit_train_forward(m, x, cond, seed=99)
        torch.cuda.synchronize()
        sv = saved[0]
        outs.append(dict(h0=sv["h0"], h1=sv["h1"], h2=sv["h2"], h3=sv["h3"], gate=sv["gate"], h4=sv["h4"], y0=sv["y0"],
                         battn_in=sv["battn"]["x"], film_b=sv["battn"]["film"], gattn_in=sv["gattn"]["x"], reg_g=sv["gattn"]["reg_in"], y=y))
for k in outs[0]:
    d1 = (outs[1][k] - outs[0][k]).abs().max().item()
    d2 = (outs[2][k] - outs[1][k]).abs().max().item()
    print(f"{k:10s} run1-run0 {d1:.3e}   run2-run1 {d2:.3e}")
