import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from vit_grid_model_b200 import MetNet3
for name in ("metnet3_small128.pt", "metnet3_wide256.pt", "metnet3_wide512.pt", "metnet3_12hr_b1.pt"):
    f = torch.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden") + "/" + name, weights_only=False)
    cfg = synth.GridConfig(**f["cfg"])
    m = MetNet3(**cfg.metnet3_kwargs())
    m.load_state_dict(synth.make_state_dict(synth.metnet3_spec(cfg), seed=f["weight_seed"]), strict=True)
    m = m.cuda().eval().set_precision("tf32")
    x, ts, _ = synth.make_inputs(cfg, f["B"], seed=f["input_seed"])
    with torch.no_grad():
        y = m(x.cuda(), timestamps=ts.cuda()).cpu()
    print(name, "tf32 rel err", ((y - f["y"]).abs().max() / f["y"].abs().max()).item())
