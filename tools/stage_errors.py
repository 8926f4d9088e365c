"""GPU diagnostic: per-stage relative error of the bf16 path against the fp32 path and the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from oracle.metnet3_oracle import metnet3_forward
from vit_grid_model_b200 import MetNet3

cfg = synth.CFG_SMALL128
sd = synth.make_state_dict(synth.metnet3_spec(cfg), seed=0)
x, ts, _ = synth.make_inputs(cfg, 2, seed=1234)
ref, feats = metnet3_forward(x, ts, sd, cfg, return_features=True)
m = MetNet3(**cfg.metnet3_kwargs()); m.load_state_dict(sd); m = m.cuda().eval()
caps = {}
for prec in ("fp32", "bf16"):
    m.set_precision(prec); m._capture = {}
    with torch.no_grad():
        y = m(x.cuda(), timestamps=ts.cuda())
    caps[prec] = dict(m._capture); caps[prec]["out"] = y
    m._capture = None; m.vit._capture = None
rel = lambda a, b: ((a.float().cpu() - b.float().cpu()).abs().max() / b.float().cpu().abs().max()).item()
rms = lambda a, b: ((a.float().cpu() - b.float().cpu()).pow(2).mean().sqrt() / b.float().cpu().pow(2).mean().sqrt()).item()
for k in caps["fp32"]:
    a, b = caps["bf16"][k], caps["fp32"][k]
    extra = f"  fp32-vs-oracle {rel(b, feats[k]):.2e}" if k in feats else ""
    print(f"{k:12s} bf16-vs-fp32 max {rel(a, b):.3e} rms {rms(a, b):.3e}{extra}")
print("out vs oracle: fp32", rel(caps["fp32"]["out"], ref), "bf16", rel(caps["bf16"]["out"], ref))
