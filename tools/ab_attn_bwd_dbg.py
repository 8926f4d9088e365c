"""clock stamps of the tcgen05 attention backward (CTA 0): VG_ABTC_DBG=1 python tools/ab_attn_bwd_dbg.py N"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_grid_model_b200 import ops_train as ot
N = int(sys.argv[1]) if len(sys.argv) > 1 else 192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 26
Hl, Wl, win, R, heads, dh = 42, 35, 7, 4, 32, 32
S, nwin, inner = R + win * win, (Hl // win) * (Wl // win), heads * dh
rows = N * nwin * S
qkv = torch.randn(rows, 3 * inner, device="cuda").to(torch.bfloat16)
datt = torch.randn(rows, inner, device="cuda").to(torch.bfloat16)
qg, kg = torch.ones(inner, device="cuda"), torch.ones(inner, device="cuda")
bt = torch.randn(170, heads, device="cuda")
dqg, dkg, dbt = torch.zeros_like(qg), torch.zeros_like(kg), torch.zeros_like(bt)
os.environ["VG_ATTN_BWD_TC"] = "1"
for _ in range(2):
    ot.attn_core_bwd(qkv, datt, qg, kg, bt, N, Hl, Wl, win, R, heads, dh, dqg, dkg, dbt, tf32=True, want_att=True, drop=(1, 1, T))
torch.cuda.synchronize()
