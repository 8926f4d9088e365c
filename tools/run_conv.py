"""Run the 3x3-conv implicit-GEMM kernel (+LN/FiLM/ReLU/residual epilogue) alone at the bench shape (for ncu / timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_grid_model_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 768
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
HP, WP, C = 84, 70, 128
g = torch.Generator().manual_seed(0)
x = ops.pg_from_nchw(torch.randn(N, C, HP, WP, generator=g).cuda(), torch.bfloat16)
w = (torch.randn(C, 9 * C, generator=g) / 34).cuda().bfloat16()
b, ga, be = torch.zeros(C).cuda(), torch.ones(C).cuda(), torch.zeros(C).cuda()
film = torch.randn(N, 2 * C, generator=g).cuda() * 0.1
res = torch.randn(x.shape[0], C, generator=g, dtype=torch.float32).cuda() if N <= 96 else torch.zeros(x.shape[0], C, device="cuda")
out = torch.empty_like(x)
flops = 2.0 * N * HP * WP * C * 9 * C
for _ in range(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv3x3_ln(x, w, b, ga, be, 1e-5, film, res, out, N, HP, WP)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"N={N} ms={ms:.3f} algorithmic TFLOP/s={flops / ms / 1e9:.1f}")
# mainloop-only reference point: same shifted GEMM with the plain store epilogue
o2 = torch.empty_like(x)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.gemm(x, w, ntaps=9, tap_shift=ops.conv_tap_shifts(WP), out=o2)
    e1.record()
    torch.cuda.synchronize()
    print(f"plain-store epilogue: ms={e0.elapsed_time(e1):.3f} TFLOP/s={flops / e0.elapsed_time(e1) / 1e9:.1f}")
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv3x3_ln(x, w, b, ga, be, 1e-5, None, None, out, N, HP, WP)
    e1.record()
    torch.cuda.synchronize()
    print(f"LN epilogue, no film/res: ms={e0.elapsed_time(e1):.3f} TFLOP/s={flops / e0.elapsed_time(e1) / 1e9:.1f}")
# fp32 residual (skip connection) variants, as the network runs them
resf = torch.randn(x.shape[0], C, generator=torch.Generator(device="cuda").manual_seed(1), dtype=torch.float32, device="cuda")
copy = torch.empty(x.shape[0], C, dtype=torch.float32, device="cuda")
for label, oc in (("fp32 residual", None), ("fp32 residual + fp32 copy", copy)):
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv3x3_ln(x, w, b, ga, be, 1e-5, film, resf, out, N, HP, WP, out_copy=oc)
        e1.record()
        torch.cuda.synchronize()
        print(f"{label}: ms={e0.elapsed_time(e1):.3f} TFLOP/s={flops / e0.elapsed_time(e1) / 1e9:.1f}")
