"""BASELINE configs[3]: enlarged 512x512 domain, 2x depth MaxViT -- inference and one training step (functional check + timing)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from vit_grid_model_b200 import MetNet3, FlatAdamW, focal_r_loss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = synth.GridConfig(H=512, W=512, vit_depth=2)
print("padded", cfg.HP, cfg.WP, "low-res", cfg.HP // 2, cfg.WP // 2, "fields", B * cfg.L)
m = MetNet3(**cfg.metnet3_kwargs())
m.load_state_dict(synth.make_state_dict(synth.metnet3_spec(cfg), seed=0), strict=True)
m = m.cuda().eval()
x, ts, target = synth.make_inputs(cfg, B, seed=1)
x, ts, target = x.cuda(), ts.cuda(), target.cuda()
with torch.no_grad():
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y = m(x, timestamps=ts)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"inference: {dt * 1e3:.1f} ms for {B * cfg.L} fields -> {B * cfg.L / dt:.1f} fields/s; finite={torch.isfinite(y).all().item()} "
      f"mean={y.mean().item():.3f}; {B * cfg.L * 1382e9 / dt / 1e12:.0f} TFLOP/s (reference-graph FLOPs)")
print("max mem GB", torch.cuda.max_memory_allocated() / 2**30)
m.train()
opt = FlatAdamW(m, lr=1e-5)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.zero_grad()
    loss = focal_r_loss(m(x, timestamps=ts), target)
    loss.backward()
    opt.step()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"train step: {dt * 1e3:.1f} ms for {B * cfg.L} fields -> {B * cfg.L / dt:.1f} fields/s; loss={loss.item():.4f}")
print("max mem GB", torch.cuda.max_memory_allocated() / 2**30)
