"""The MBConv expand GEMM (+ folded BN + GELU, fp16 out) and the same GEMM without the activation, alone at bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_grid_model_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1128960
g = torch.Generator().manual_seed(0)
K, N = 128, 512
A = torch.randn(M, K, generator=g).cuda()
W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
sc, sh = torch.ones(N).cuda(), torch.zeros(N).cuda()
for act in (1, 0):
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(A, W, scale=sc, shift=sh, act=act, out=out, tf32=True, out_dtype=torch.float16)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    print(f"expand M={M} act={act}: {ms:.3f} ms")
