"""3xTF32 projection GEMM against a float64 product: error of the split path, the tf32 path and the exact-fp32 SIMT path, and their times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_grid_model_b200 import ops

torch.manual_seed(0)
for M, K, N in ((19080, 512, 6144), (4096, 128, 384), (1000, 2048, 512)):
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    ref = (A.double() @ W.double().t())
    scale = ref.abs().max().item()
    for name, kw in (("simt", {}), ("x3", {"x3": True}), ("tf32", {"tf32": True})):
        for _ in range(2):
            out = ops.gemm(A, W, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = ops.gemm(A, W, **kw)
        e1.record(); torch.cuda.synchronize()
        err = (out.double() - ref).abs().max().item() / scale
        print(f"M={M} K={K} N={N} {name:5s} max err / max |ref| = {err:.3e}   {e0.elapsed_time(e1) / 5:.3f} ms")
