"""The depthwise 3x3 + BN + GELU kernel (fp16 hidden tensor) alone at bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_grid_model_b200 import _lib
N, H, W, C = (int(sys.argv[1]) if len(sys.argv) > 1 else 768), 42, 35, 512
x = torch.randn(N, H, W, C, device="cuda").half()
w9 = torch.randn(9, C, device="cuda") * 0.3
sc, sh = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
out = torch.empty_like(x)
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call("vg_dw3x3_fwd", 3, x.data_ptr(), w9.data_ptr(), sc.data_ptr(), sh.data_ptr(), 1, out.data_ptr(), 0, N, H, W, C, torch.cuda.current_stream().cuda_stream)
    e1.record(); torch.cuda.synchronize()
print(f"dw3x3 fp16 N={N}: {e0.elapsed_time(e1):.3f} ms")
