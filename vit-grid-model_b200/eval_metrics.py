"""Device-side evaluation statistics: the accumulators of the reference's test loop
(/root/reference/src/evaluation_vit.py:239-455) without its ~200 ``.item()`` synchronisations per batch.

    ev = EvalMetrics(output_dim=12)
    for batch in loader:
        preds = model(x, timestamps=ts).reshape(B, 12, -1)
        ev.update(preds, truth, truth_classes, last_obs, sim_21h, sim_avg)     # one kernel, no host sync
    stats = ev.result()                                                        # one device->host copy

``result()`` returns the reference's variables under the reference's names (``total_a1`` .. ``sim_avg_total_d4``,
``_TP1`` .., the ``(3*output_dim,)`` arrays ``_TP`` / ``p_TP`` / ``sim_21h_RMSE_np`` / ``valid_count`` .., the error sums
``valid_diff_sum*`` / ``valid_norm_diff_*``, ``val_loss_sum``, ``val_acc``) plus the post-loop ``nmb*`` / ``nme*`` / ``r*``
(:507-523, :572-576), all derived on the host from three device tables (vg_eval_metrics in include/vitgrid.h).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_METHODS = (("", "_", ""), ("per_", "p_", "_p"), ("sim_21h_", "sim_21h_", "_sim_21h"), ("sim_avg_", "sim_avg_", "_sim_avg"))


class EvalMetrics:
    def __init__(self, output_dim: int, boundaries=(15.0, 35.0, 75.0), device="cuda"):
        _lib.require_device()
        self.L = int(output_dim)
        self.bounds = tuple(float(b) for b in boundaries)
        if len(self.bounds) != 3:
            raise ValueError("three class boundaries (range_4class, evaluation_vit.py:194)")
        dev = torch.device(device)
        self.counts = torch.zeros(4, self.L, 4, 5, dtype=torch.int64, device=dev)
        self.sums = torch.zeros(4, self.L, 5, 2, dtype=torch.float64, device=dev)
        self.glob = torch.zeros(22, dtype=torch.float64, device=dev)
        self.nonzero = torch.zeros(1, dtype=torch.int64, device=dev)
        self.loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        self.entries = 0
        self.steps = 0
        self._work = None

    def reset(self):
        for t in (self.counts, self.sums, self.glob, self.nonzero, self.loss_sum):
            t.zero_()
        self.entries = self.steps = 0

    def update(self, preds, truth, truth_classes, last_obs, sim_21h, sim_avg, clamp_preds: bool = True):
        """preds / truth / sim_21h / sim_avg: (B, output_dim, P) fp32; truth_classes: same shape, int32 or int64, -1 = no
        label; last_obs: (B, P) fp32 (persistence).  ``preds`` is clamped at zero IN PLACE (evaluation_vit.py:254)."""
        B, L, P = preds.shape
        if L != self.L:
            raise ValueError(f"output_dim {L} != {self.L}")
        for name, t, shape, dts in (("preds", preds, (B, L, P), (torch.float32,)), ("truth", truth, (B, L, P), (torch.float32,)),
                                    ("truth_classes", truth_classes, (B, L, P), (torch.int32, torch.int64)),
                                    ("last_obs", last_obs, (B, P), (torch.float32,)), ("sim_21h", sim_21h, (B, L, P), (torch.float32,)),
                                    ("sim_avg", sim_avg, (B, L, P), (torch.float32,))):
            if not t.is_cuda:
                raise _lib.VitGridError(f"EvalMetrics.update: {name} must be a CUDA tensor (no CPU fallback)")
            if tuple(t.shape) != shape or t.dtype not in dts or not t.is_contiguous():
                raise ValueError(f"EvalMetrics.update: {name} must be a contiguous {shape} tensor of {dts}")
        need = _lib.load().vg_eval_metrics_workspace(B, L, P)
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty(need, dtype=torch.float64, device=preds.device)
        with torch.cuda.device(preds.device):              # launch on the tensors' device, whatever the current one is
            _lib.call("vg_eval_metrics", preds.data_ptr(), truth.data_ptr(), truth_classes.data_ptr(),
                      int(truth_classes.dtype == torch.int64), last_obs.data_ptr(), sim_21h.data_ptr(), sim_avg.data_ptr(), B, L, P,
                      self.bounds[0], self.bounds[1], self.bounds[2], int(clamp_preds), self.counts.data_ptr(), self.sums.data_ptr(),
                      self.glob.data_ptr(), self.nonzero.data_ptr(), self.loss_sum.data_ptr(), self._work.data_ptr(),
                      self._work.numel(), torch.cuda.current_stream().cuda_stream)
        self.entries += B * L * P
        self.steps += 1

    def tables(self):
        """(counts (4,L,4,5) int64, sums (4,L,5,2) f64, glob (22,) f64, nonzero, loss_sum) on the host: one sync"""
        packed = torch.cat([self.counts.double().flatten(), self.sums.flatten(), self.glob, self.nonzero.double(), self.loss_sum]).cpu().numpy()
        n1, n2 = self.counts.numel(), self.sums.numel()
        counts = np.rint(packed[:n1]).astype(np.int64).reshape(4, self.L, 4, 5)      # exact below 2^53
        sums = packed[n1:n1 + n2].reshape(4, self.L, 5, 2)
        glob = packed[n1 + n2:n1 + n2 + 22]
        return counts, sums, glob, int(round(packed[-2])), float(packed[-1])

    def result(self) -> dict:
        counts, sums, glob, nonzero, loss_sum = self.tables()
        L = self.L
        out = {"val_loss_sum": loss_sum, "val_step_cnt": self.steps, "valid_entry_cnt": self.entries,
               "valid_nonzero_entry_cnt": nonzero}
        conf = counts.sum(axis=1)                                        # (method, class of the method, truth class -1..3)
        out["val_acc"] = float(sum(conf[0, c, c + 1] for c in range(4)))                           # :340
        y_sum, y_sq = glob[20], glob[21]
        n = float(self.entries)
        for m, (pre_tot, pre_arr, suf) in enumerate(_METHODS):
            for ai, a in enumerate("abcd"):                                                         # :345-415
                for bi, b in enumerate("1234"):
                    out[f"{pre_tot}total_{a}{b}"] = float(conf[m, ai, bi + 1])
            out["valid_diff_sum" + suf] = float(sums[m, :, :, 0].sum())                             # :293-307
            out["valid_diff_squares_sum" + suf] = float(sums[m, :, :, 1].sum())
            out["valid_norm_diff_sum" + suf] = float(glob[m * 2])                                   # :314-326 (the reference
            out["valid_norm_diff_abs_sum" + suf] = float(glob[m * 2 + 1])                           #  skips _sim_avg; kept here)
            tp, tn, fp, fn = (np.zeros(3 * L) for _ in range(4))
            rmse, mae, cnt = np.zeros(3 * L), np.zeros(3 * L), np.zeros(3 * L)
            for i in range(1, 4):                                                                   # :432-463
                c = counts[m]                                            # (L, a, t) with t index = class + 1
                hi_a, lo_a = c[:, i:, :].sum(axis=1), c[:, :i, :].sum(axis=1)                       # method class > i-1 / < i
                sl = slice((i - 1) * L, i * L)
                tp[sl] = hi_a[:, i + 1:].sum(axis=1)                     # truth class > i-1
                fn[sl] = lo_a[:, i + 1:].sum(axis=1)
                tn[sl] = lo_a[:, 1:i + 1].sum(axis=1)                    # truth class < i and > -1
                fp[sl] = hi_a[:, 1:i + 1].sum(axis=1)
                rmse[sl] = sums[m, :, i + 1:, 1].sum(axis=1)
                mae[sl] = sums[m, :, i + 1:, 0].sum(axis=1)
                cnt[sl] = c[:, :, i + 1:].sum(axis=(1, 2))
            out[pre_arr + "TP"], out[pre_arr + "TN"], out[pre_arr + "FP"], out[pre_arr + "FN"] = tp, tn, fp, fn
            out[pre_arr + "RMSE_np"], out[pre_arr + "MAE_np"] = rmse, mae
            if m == 0:
                out["valid_count"] = cnt
                for i in range(1, 4):                                                               # :418-429: no `> -1` term
                    a_hi, a_lo = conf[0, i:, :].sum(axis=0), conf[0, :i, :].sum(axis=0)
                    out[f"_TP{i}"] = float(a_hi[i + 1:].sum())
                    out[f"_FN{i}"] = float(a_lo[i + 1:].sum())
                    if i == 1:                                           # `cur_preds == 0`
                        out["_TN1"], out["_FP1"] = float(a_lo[1]), float(a_hi[1])
                    else:                                                # `cur_preds < i` includes the unlabeled -1
                        out[f"_TN{i}"], out[f"_FP{i}"] = float(a_lo[:i + 1].sum()), float(a_hi[:i + 1].sum())
            # post-loop: normalised mean bias / error (percent) and Pearson r (:507-523, :552-576)
            v_sum, v_sq, vy = glob[8 + m * 3], glob[9 + m * 3], glob[10 + m * 3]
            out["nmb" + suf] = (v_sum - y_sum) / y_sum * 100
            out["nme" + suf] = out["valid_diff_sum" + suf] / y_sum * 100
            cov = vy - v_sum * y_sum / n
            out["r" + suf] = cov / (np.sqrt(v_sq - v_sum * v_sum / n) * np.sqrt(y_sq - y_sum * y_sum / n))
        return out
