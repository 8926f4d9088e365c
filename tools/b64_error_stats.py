import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from vit_grid_model_b200 import MetNet3
cfg = synth.CFG_12HR
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = MetNet3(**cfg.metnet3_kwargs()); m.load_state_dict(synth.make_state_dict(synth.metnet3_spec(cfg), seed=0)); m = m.cuda().eval()
x, ts, _ = synth.make_inputs(cfg, B, seed=5)
with torch.no_grad():
    y = m.set_precision("bf16")(x.cuda(), timestamps=ts.cuda())
    y32 = m.set_precision("fp32")(x.cuda(), timestamps=ts.cuda())
e = (y - y32)
print("global max/max", (e.abs().max() / y32.abs().max()).item(), "rms/rms", (e.pow(2).mean().sqrt() / y32.pow(2).mean().sqrt()).item())
print("ref stats: mean", y32.mean().item(), "std", y32.std().item(), "absmax", y32.abs().max().item())
pg = e.flatten(2).norm(dim=2) / y32.flatten(2).norm(dim=2)        # (B, L)
print("per-grid rel L2: mean", pg.mean().item(), "max", pg.max().item())
print("per-lead mean rel L2:", [round(v, 5) for v in pg.mean(0).tolist()])
print("per-lead ref rms:", [round(v, 2) for v in y32.flatten(2).pow(2).mean(2).sqrt().mean(0).tolist()])
print("per-lead err rms:", [round(v, 4) for v in e.flatten(2).pow(2).mean(2).sqrt().mean(0).tolist()])
pm = e.flatten(2).abs().amax(2) / y32.flatten(2).abs().amax(2)
print("per-grid max/max: mean", pm.mean().item(), "max", pm.max().item())
