"""Run the fused attention kernel alone on the 12hr-model map (for ncu / timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_grid_model_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W, C, heads, dh, w, R = 42, 35, 128, 32, 32, 7, 4
g = torch.Generator().manual_seed(0)
x = torch.randn(N, H, W, C, generator=g).cuda()
reg = torch.randn(R, C, generator=g).cuda()
film = torch.randn(N, 2 * C, generator=g).cuda()
wqkv = (torch.randn(heads * 96, C, generator=g) / 11.3).cuda()
wout = (torch.randn(heads, C, dh, generator=g) / 32).cuda()
qg, kg = torch.ones(heads * dh).cuda(), torch.ones(heads * dh).cuda()
bias = torch.randn(170, heads, generator=g).cuda()
tab = ops.pack_head_tables(bias, qg, kg)
for _ in range(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y, r = ops.attn_fused(x, reg, film, wqkv, wout, tab, w, R, False, True, heads, dh)
    e1.record()
    torch.cuda.synchronize()
    print(f"N={N} windows={N*30} ms={e0.elapsed_time(e1):.3f}")

# per-phase clock stamps of one compute thread (CTA 0, first tile)
dbg = torch.zeros(2 * heads * 8, dtype=torch.int64, device="cuda")
os.environ["VG_ATTN_DBG"] = hex(dbg.data_ptr())
ops.attn_fused(x, reg, film, wqkv, wout, tab, w, R, False, True, heads, dh)
torch.cuda.synchronize()
dall = dbg.cpu().view(2, heads, 8)
d = dall[0]
names = ["wait s_done", "S load + s_free", "bias + max (+ hoisted waits)", "pair_sync (max)", "exp2 + sum", "pair_sync (sum) + normalise + pack + P store issue", "wait P store + p_ready"]
delta = (d[:, 1:] - d[:, :-1]).float()
print("head period (cycles):", (d[1:, 0] - d[:-1, 0]).float()[2:].mean().item())
for i, nm in enumerate(names):
    print(f"  {nm:24s} {delta[2:, i].mean().item():8.0f}")
print("  next-head gap           ", (d[1:, 0] - d[:-1, 7]).float()[2:].mean().item())

m = dall[1]
mn = ["wait WO / WQ (hoisted)", "wait p_ready", "issue PV", "wait pv_done + issue out", "wait s_free, qk_ready + issue S(h+2)", "issue QKV(h+3)", "-"]
md = (m[:, 1:] - m[:, :-1]).float()
print("MMA warp, head period:", (m[1:, 0] - m[:-1, 0]).float()[2:-2].mean().item())
for i, nm in enumerate(mn):
    print(f"  {nm:32s} {md[2:-2, i].mean().item():8.0f}")
