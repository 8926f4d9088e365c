"""One process, two GPUs: the model runs on cuda:1 while cuda:0 is the current device (the reference's nn.DataParallel usage);
predictions must equal the cuda:0 run bit for bit (per-device function attributes / SM counts, device guards)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from vit_grid_model_b200 import MetNet3

cfg = synth.CFG_12HR
sd = synth.make_state_dict(synth.metnet3_spec(cfg), seed=0)
x, ts, _ = synth.make_inputs(cfg, 2, seed=1)
outs = []
for dev in ("cuda:0", "cuda:1", "cuda:0"):
    m = MetNet3(**cfg.metnet3_kwargs())
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    torch.cuda.set_device(0)
    with torch.no_grad():
        y = m(x.to(dev), timestamps=ts.to(dev))
    torch.cuda.synchronize(dev)
    outs.append(y.cpu())
    print(dev, "ok", float(y.abs().mean()))
assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
# training step on the second device
m = MetNet3(**cfg.metnet3_kwargs()); m.load_state_dict(sd, strict=True); m = m.to("cuda:1").train()
from vit_grid_model_b200 import focal_r_loss
x1, ts1, tg = synth.make_inputs(cfg, 2, seed=3)
loss = focal_r_loss(m(x1.to("cuda:1"), timestamps=ts1.to("cuda:1")), tg.to("cuda:1"))
loss.backward()
torch.cuda.synchronize("cuda:1")
print("train on cuda:1 ok, loss", loss.item(), "grad finite", all(torch.isfinite(p.grad).all().item() for p in m.parameters() if p.grad is not None))
print("two-device check passed")
