"""BASELINE configs[4]: MetNet-3-style backbone, n_start_channels 512, 32 heads x dim_head 64, MaxViT depth 4,
82x67 domain -- inference timing on one GPU (the training kernels exist for 128 channels only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from vit_grid_model_b200 import MetNet3

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = synth.GridConfig(dim=512, heads=32, dim_head=64, vit_depth=4)
m = MetNet3(**cfg.metnet3_kwargs())
print("params", sum(p.numel() for p in m.parameters()))
m.load_state_dict(synth.make_state_dict(synth.metnet3_spec(cfg), seed=0), strict=True)
m = m.cuda().eval()
x, ts, _ = synth.make_inputs(cfg, B, seed=1)
x, ts = x.cuda(), ts.cuda()
with torch.no_grad():
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y = m(x, timestamps=ts)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n = B * cfg.L
    print(f"inference: {dt * 1e3:.1f} ms for {n} fields -> {n / dt:.1f} fields/s; finite={torch.isfinite(y).all().item()} "
          f"mean={y.mean().item():.3f}; {n * 370.9e9 / dt / 1e12:.0f} TFLOP/s (reference-graph FLOPs, 370.9 GF/field)")
    print("max mem GB", torch.cuda.max_memory_allocated() / 2**30)
    m.set_precision("fp32")
    xs, tss = x[:1], ts[:1]
    y32 = m(xs, timestamps=tss)
    m.set_precision("bf16")
    yb = m(xs, timestamps=tss)
    print("bf16-mode vs exact-fp32 path, rel err:", ((yb - y32).abs().max() / y32.abs().max()).item())
