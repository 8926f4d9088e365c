"""Host time vs GPU time of one training step (is the step launch-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from vit_grid_model_b200 import MetNet3, FlatAdamW, focal_r_loss

cfg = synth.CFG_12HR
Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
model = MetNet3(**cfg.metnet3_kwargs())
model.load_state_dict(synth.make_state_dict(synth.metnet3_spec(cfg), seed=0), strict=True)
model = model.to(dev).train().set_precision("bf16")
opt = FlatAdamW(model, lr=1e-5)
x, ts, target = synth.make_inputs(cfg, Bt, seed=4321)
x, ts, target = x.to(dev), ts.to(dev), target.to(dev)

def step():
    opt.zero_grad()
    loss = focal_r_loss(model(x, timestamps=ts), target)
    loss.backward()
    opt.step()
    return loss

for _ in range(3):
    step()
torch.cuda.synchronize()
for n in (1, 4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        step()
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"steps={n}: host enqueue {1e3 * (t1 - t0) / n:.2f} ms/step, GPU {e0.elapsed_time(e1) / n:.2f} ms/step, wall {1e3 * (t2 - t0) / n:.2f} ms/step")
