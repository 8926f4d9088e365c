"""Run the fused attention kernel repeatedly on the same input and compare the outputs bit for bit (race detector)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_grid_model_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
H, W, C, heads, dh, w, R = 42, 35, 128, 32, 32, 7, 4
g = torch.Generator().manual_seed(0)
x = torch.randn(N, H, W, C, generator=g).cuda()
reg = torch.randn(R, C, generator=g).cuda()
film = torch.randn(N, 2 * C, generator=g).cuda()
wqkv = (torch.randn(heads * 96, C, generator=g) / 11.3).cuda()
wout = (torch.randn(heads, C, dh, generator=g) / 32).cuda()
qg, kg = (0.5 + torch.rand(heads * dh, generator=g)).cuda(), (0.5 + torch.rand(heads * dh, generator=g)).cuda()
bias = torch.randn(170, heads, generator=g).cuda()
tab = ops.pack_head_tables(bias, qg, kg)
for drop in ((0, 0, 0), (1234, 3, 26)):
    for grid_mode in (False, True):
        ref = None
        bad = 0
        for k in range(8):
            y, r = ops.attn_fused(x, reg, film, wqkv, wout, tab, w, R, grid_mode, True, heads, dh, drop=drop)
            torch.cuda.synchronize()
            if ref is None:
                ref = (y.clone(), r.clone())
            elif not (torch.equal(y, ref[0]) and torch.equal(r, ref[1])):
                bad += 1
                d = (y - ref[0]).abs()
                print(f"  run {k}: max diff {d.max().item():.3e}, {int((d > 0).sum())} elements differ of {d.numel()}")
        print(f"drop={drop} grid={grid_mode}: {bad} of 7 repeats differ")
