"""Run the plain-store tcgen05 GEMMs of the MaxViT block alone at bench shapes (for ncu / timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_grid_model_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1128960
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g = torch.Generator().manual_seed(0)
for (K, N, act) in ((128, 512, 1), (128, 512, 0), (512, 128, 0)):
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    sc, sh = torch.ones(N).cuda(), torch.zeros(N).cuda()
    out = torch.empty(M, N, device="cuda")
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(A, W, scale=sc, shift=sh, act=act, out=out, tf32=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    print(f"M={M} K={K} N={N} act={act}: {ms:.3f} ms  write {M * N * 4 / ms / 1e6:.0f} GB/s  read {M * K * 4 / ms / 1e6:.0f} GB/s")
    del A, out
