"""A/B of the two attention-core backward kernels on bf16 tensors: tcgen05 (vg_attn_bwd_tc.cu) against mma.sync (vg_bwd_vit.cu),
plus an fp32 autograd reference of the same math on the same bf16 inputs.  usage: python tools/ab_attn_bwd.py [N] [drop_thresh]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_grid_model_b200 import ops_train as ot


def run(tc, qkv, datt, qg, kg, bt, N, Hl, Wl, win, R, heads, dh, drop):
    os.environ["VG_ATTN_BWD_TC"] = "1" if tc else "0"
    dqg, dkg, dbt = torch.zeros_like(qg), torch.zeros_like(kg), torch.zeros_like(bt)
    dqkv, att = ot.attn_core_bwd(qkv, datt, qg, kg, bt, N, Hl, Wl, win, R, heads, dh, dqg, dkg, dbt, tf32=True, want_att=True, drop=drop)
    torch.cuda.synchronize()
    return dqkv, att, dqg, dkg, dbt


def reference(qkv, datt, qg, kg, bt, N, nwin, S, win, R, heads, dh):
    """fp32 autograd of maxvit.py:189-213 on one head layout [rows][3*inner] (no dropout)"""
    inner = heads * dh
    x = qkv.float().view(N * nwin, S, 3, heads, dh).permute(2, 0, 3, 1, 4).detach().requires_grad_(True)   # (3, W, h, S, d)
    qg_ = qg.view(heads, 1, dh).detach().requires_grad_(True)
    kg_ = kg.view(heads, 1, dh).detach().requires_grad_(True)
    bt_ = bt.detach().requires_grad_(True)
    q, k, v = x[0], x[1], x[2]
    l2 = lambda t: torch.nn.functional.normalize(t, dim=-1)
    qh = l2(q) * (dh ** 0.5) * qg_
    kh = l2(k) * (dh ** 0.5) * kg_
    sim = qh @ kh.transpose(-1, -2)
    W2 = 2 * win - 1
    nb = W2 * W2 + 1
    idx = torch.full((S, S), nb - 1, dtype=torch.long)
    for i in range(R, S):
        for j in range(R, S):
            a, b = divmod(i - R, win)
            c, d = divmod(j - R, win)
            idx[i, j] = (a - c + win - 1) * W2 + (b - d + win - 1)
    bias = bt_[idx.cuda()].permute(2, 0, 1)       # (h, S, S)
    att = (sim + bias).softmax(-1) @ v
    out = att.permute(0, 2, 1, 3).reshape(N * nwin * S, inner)
    out.backward(datt.float())
    dqkv = x.grad.permute(1, 3, 0, 2, 4).reshape(N * nwin * S, 3 * inner)
    return dqkv, out.detach(), qg_.grad.reshape(-1), kg_.grad.reshape(-1), bt_.grad


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    Hl, Wl, win, R, heads, dh = (int(os.environ.get("AB_HL", 42)), int(os.environ.get("AB_WL", 35)), 7, 4, 32, 32)
    S, nwin, inner = R + win * win, (Hl // win) * (Wl // win), heads * dh
    rows = N * nwin * S
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv = torch.randn(rows, 3 * inner, device="cuda", generator=g).to(torch.bfloat16)
    datt = torch.randn(rows, inner, device="cuda", generator=g).to(torch.bfloat16)
    qg = 1 + 0.2 * torch.randn(inner, device="cuda", generator=g)
    kg = 1 + 0.2 * torch.randn(inner, device="cuda", generator=g)
    bt = 0.5 * torch.randn((2 * win - 1) ** 2 + 1, heads, device="cuda", generator=g)
    drop = (1234, 1, T)
    a = run(False, qkv, datt, qg, kg, bt, N, Hl, Wl, win, R, heads, dh, drop)
    b = run(True, qkv, datt, qg, kg, bt, N, Hl, Wl, win, R, heads, dh, drop)
    names = ["dqkv", "att", "dqgamma", "dkgamma", "dbias"]
    for nm, x, y in zip(names, a, b):
        print(f"tc vs mma.sync  {nm:8s} rel {rel(y, x):.3e}  finite {bool(torch.isfinite(y.float()).all())}")
    for w, nm in enumerate(("dq", "dk", "dv")):
        print(f"   {nm}: rel {rel(b[0].view(rows, 3, inner)[:, w], a[0].view(rows, 3, inner)[:, w]):.3e}")
    if T == 0 and N <= 4:
        r = reference(qkv, datt, qg, kg, bt, N, nwin, S, win, R, heads, dh)
        for nm, x, y, z in zip(names, a, b, r):
            print(f"vs fp32 autograd {nm:8s} mma.sync {rel(x, z):.3e}   tc {rel(y, z):.3e}")
    for tc in (False, True):
        for _ in range(2):
            run(tc, qkv, datt, qg, kg, bt, N, Hl, Wl, win, R, heads, dh, drop)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dqg, dkg, dbt = torch.zeros_like(qg), torch.zeros_like(kg), torch.zeros_like(bt)
        e0.record()
        for _ in range(5):
            ot.attn_core_bwd(qkv, datt, qg, kg, bt, N, Hl, Wl, win, R, heads, dh, dqg, dkg, dbt, tf32=True, want_att=True, drop=drop)
        e1.record()
        torch.cuda.synchronize()
        print(f"{'tcgen05 ' if tc else 'mma.sync'}: {e0.elapsed_time(e1) / 5:.3f} ms  ({N} fields)")


if __name__ == "__main__":
    main()
