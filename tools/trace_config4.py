"""Per-entry-point time of one BASELINE configs[4] inference step (512 channels, 32 x 64 heads, depth 4), default precision."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from vit_grid_model_b200 import MetNet3, _lib

cfg = synth.GridConfig(dim=512, heads=32, dim_head=64, vit_depth=4)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m = MetNet3(**cfg.metnet3_kwargs())
m.load_state_dict(synth.make_state_dict(synth.metnet3_spec(cfg), seed=0), strict=True)
m = m.cuda().eval()
if len(sys.argv) > 2:
    m.set_precision(sys.argv[2])
x, ts, _ = synth.make_inputs(cfg, B, seed=1)
x, ts = x.cuda(), ts.cuda()
with torch.no_grad():
    for _ in range(2):
        m(x, timestamps=ts)
    torch.cuda.synchronize()
    _lib.TRACE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m(x, timestamps=ts); e1.record()
    torch.cuda.synchronize()
tr, _lib.TRACE = _lib.TRACE, None
agg = collections.defaultdict(lambda: [0, 0.0])
for name, tag, a, b in tr:
    k = f"{name}[{tag}]" if name in ("vg_gemm_fwd",) else name
    agg[k][0] += 1; agg[k][1] += a.elapsed_time(b)
print(f"precision {m.precision}: step {e0.elapsed_time(e1):.2f} ms for {B * cfg.L} fields")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"  {k:70s} n={n:3d} {t:8.3f} ms")
