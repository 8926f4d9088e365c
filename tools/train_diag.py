"""Diagnostics: per-parameter gradient error of the CUDA training step vs the CPU oracle (GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import synth                                    # noqa: E402
from test_train_oracle import oracle_train_step             # noqa: E402
from vit_grid_model_b200 import MetNet3, focal_r_loss       # noqa: E402

cfg = synth.CFG_SMALL128
B, wseed, iseed = 3, 0, 4321
sd_o, pred_o, loss_o = oracle_train_step(cfg, B, wseed, iseed)
for precision in sys.argv[1:] or ["fp32", "bf16"]:
    m = MetNet3(**cfg.metnet3_kwargs(), dropout=0.0)
    m.load_state_dict(synth.make_state_dict(synth.metnet3_spec(cfg), seed=wseed), strict=True)
    m = m.cuda().train().set_precision(precision)
    x, ts, target = synth.make_inputs(cfg, B, seed=iseed)
    pred = m(x.cuda(), timestamps=ts.cuda())
    loss = focal_r_loss(pred, target.cuda())
    loss.backward()
    torch.cuda.synchronize()
    e = (pred.detach().cpu() - pred_o)
    print(f"== {precision}: loss {loss.item():.6f} vs {loss_o.item():.6f}; pred max err {(e.abs().max() / pred_o.abs().max()).item():.3e} "
          f"rms {(e.pow(2).mean().sqrt() / pred_o.pow(2).mean().sqrt()).item():.3e}")
    rows = []
    for k, p in m.named_parameters():
        g, ref = p.grad.detach().float().cpu(), sd_o[k].grad
        emax = ((g - ref).abs().max() / max(ref.abs().max().item(), 1e-2)).item()
        el2 = ((g - ref).norm() / max(ref.norm().item(), 1e-2)).item()
        rows.append((emax, el2, k))
    for emax, el2, k in sorted(rows, reverse=True)[:25]:
        print(f"  {k:45s} max {emax:.3e}  l2 {el2:.3e}")
    for k in ("resnet1.blocks.0.block2.norm.b", "resnet1.blocks.0.block2.proj.bias", "resnet1.blocks.0.block2.norm.g"):
        g, ref = dict(m.named_parameters())[k].grad.detach().float().cpu().reshape(-1), sd_o[k].grad.reshape(-1)
        d = (g - ref).abs()
        top = d.topk(6)
        print(k, "worst idx", top.indices.tolist(), "err", [f"{v:.2e}" for v in top.values.tolist()], "ref", [f"{ref[i]:.3f}" for i in top.indices.tolist()],
              "median err", f"{d.median():.2e}")
