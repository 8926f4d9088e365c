"""Evaluation statistics: the fused device kernel + host derivation (EvalMetrics) against the literal CPU
restatement of the reference's test loop (oracle/eval_oracle.py, /root/reference/src/evaluation_vit.py:239-455).
Counts must be exact; floating-point sums within 1e-5 relative (fp32 per-thread partial sums, fp64 above)."""
import numpy as np
import pytest
import torch

from oracle import eval_oracle as eo


def make_batch(B, L, P, seed, i64=True):
    g = torch.Generator().manual_seed(seed)
    truth = torch.exp(np.log(18.0) + 0.7 * torch.randn(B, L, P, generator=g)).clamp_(0, 400)
    truth[torch.rand(B, L, P, generator=g) < 0.03] = 0.0                     # exercises the nonzero mask (:311)
    preds = truth * (1 + 0.3 * torch.randn(B, L, P, generator=g)) - 2.0      # some negative -> clamped (:254)
    preds[0, 0, :7] = torch.tensor([15.0, 35.0, 75.0, -1.0, -3.0, 15.000001, 74.99999])   # class edges: (lo, hi]
    cls = torch.from_numpy(eo.assign_class(truth.numpy(), eo.RANGE_4CLASS, eo.CLASS_FOUR))
    cls[torch.rand(B, L, P, generator=g) < 0.1] = -1                         # unlabeled cells (dataset.py:8-14)
    cls = cls.to(torch.int64 if i64 else torch.int32)
    last = torch.exp(np.log(18.0) + 0.7 * torch.randn(B, P, generator=g)).clamp_(0, 400)
    sim21 = truth * (1 + 0.5 * torch.randn(B, L, P, generator=g)).abs()
    simavg = truth * (1 + 0.4 * torch.randn(B, L, P, generator=g)).abs()
    return preds, truth, cls, last, sim21, simavg


def test_oracle_identities():
    """the restated accumulators satisfy the identities any confusion table must"""
    B, L, P = 3, 4, 211
    s = eo.new_state(L)
    preds, truth, cls, last, sim21, simavg = make_batch(B, L, P, 0)
    eo.update(s, preds, truth, cls, last, sim21, simavg)
    labeled = int((cls >= 0).sum())
    for pre in ("", "per_", "sim_21h_", "sim_avg_"):
        assert sum(s[f"{pre}total_{a}{b}"] for a in "abcd" for b in "1234") == labeled
    for pre in ("_", "p_", "sim_21h_", "sim_avg_"):
        tot = s[pre + "TP"] + s[pre + "TN"] + s[pre + "FP"] + s[pre + "FN"]
        per_lead = np.array([int((cls[:, j] >= 0).sum()) for j in range(L)] * 3)
        # TP and FN carry no `> -1` term but an unlabeled cell (-1) never satisfies `cp > i-1`, so the four cells partition the labeled ones
        assert np.array_equal(tot, per_lead)
    assert (preds >= 0).all() and s["valid_entry_cnt"] == B * L * P
    f = eo.finish(s)
    assert -1.0 <= f["r"] <= 1.0 and f["nme"] >= abs(f["nmb"])


@pytest.mark.gpu
@pytest.mark.parametrize("i64", [True, False])
def test_eval_metrics_kernel_matches_oracle(i64):
    from vit_grid_model_b200 import EvalMetrics
    L = 12
    ev = EvalMetrics(L)
    s = eo.new_state(L)
    for step, (B, P) in enumerate(((5, 82 * 67), (2, 82 * 67), (1, 333))):      # ragged last batches
        batch = make_batch(B, L, P, 10 + step, i64)
        dev = [t.cuda() for t in batch]
        ev.update(*dev)
        eo.update(s, *batch)
        assert torch.equal(dev[0].cpu(), batch[0])                             # both clamped preds in place, identically
    f = eo.finish(s)
    out = ev.result()
    exact = [k for k in s if "total_" in k or (k[:3] in ("_TP", "_TN", "_FP", "_FN") and len(k) == 4)] + ["val_acc", "valid_entry_cnt", "valid_nonzero_entry_cnt"]
    for k in exact:
        assert out[k] == s[k], k
    for pre in ("_", "p_", "sim_21h_", "sim_avg_"):
        for k in ("TP", "TN", "FP", "FN"):
            assert np.array_equal(out[pre + k], s[pre + k]), pre + k
        for k in ("RMSE_np", "MAE_np"):
            np.testing.assert_allclose(out[pre + k], s[pre + k], rtol=1e-5)
    assert np.array_equal(out["valid_count"], s["valid_count"])
    for k in s:
        if k.startswith("valid_diff") or k.startswith("valid_norm") or k == "val_loss_sum":
            # the reference's own fp32 torch sums carry ~1e-6 of rounding; signed relative-error sums cancel, so compare
            # them on the scale of their absolute counterpart
            scale = abs(s[k.replace("norm_diff_sum", "norm_diff_abs_sum")]) if "norm_diff_sum" in k else abs(s[k])
            assert abs(out[k] - s[k]) <= 2e-5 * scale, (k, out[k], s[k])
    for k, v in f.items():
        assert abs(out[k] - v) <= 1e-5 * max(1.0, abs(v)), (k, out[k], v)
    # bit-reproducible: a second accumulator fed the same batches gives identical tables
    ev2 = EvalMetrics(L)
    for step, (B, P) in enumerate(((5, 82 * 67), (2, 82 * 67), (1, 333))):
        ev2.update(*[t.cuda() for t in make_batch(B, L, P, 10 + step, i64)])
    for a, b in zip(ev.tables(), ev2.tables()):
        assert np.array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.gpu
def test_eval_metrics_errors():
    from vit_grid_model_b200 import EvalMetrics, VitGridError
    ev = EvalMetrics(4)
    batch = make_batch(2, 4, 50, 0)
    with pytest.raises(VitGridError):
        ev.update(*batch)                                                      # CPU tensors: no fallback
    dev = [t.cuda() for t in batch]
    with pytest.raises(ValueError):
        ev.update(dev[0][:, :3].contiguous(), *dev[1:])                        # wrong output_dim
    with pytest.raises(ValueError):
        ev.update(dev[0], dev[1], dev[2].float(), *dev[3:])                    # classes must be integer
