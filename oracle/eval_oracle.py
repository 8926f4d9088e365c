"""CPU restatement of the reference's evaluation accumulators (TEST INFRASTRUCTURE ONLY -- nothing in the product
imports this file).

Follows /root/reference/src/evaluation_vit.py:239-455 statement by statement for ONE batch, with the reference's own
expressions (masked sums, ``np.select`` class assignment, per-lead / per-threshold loops) and the reference's variable
names as dictionary keys, plus the post-loop quantities of :507-576.  ``evaluation_vit.py`` itself cannot be imported here
(it needs ``xarray`` and private data), and its loop body is not a callable: **parity unpinned by execution** -- the
restatement is literal so it can be checked against the cited lines by eye.
"""
from __future__ import annotations

import numpy as np
import torch

RANGE_4CLASS = [(-1, 15), (15, 35), (35, 75), (75, np.inf)]      # evaluation_vit.py:194
CLASS_FOUR = [0, 1, 2, 3]                                        # :195


def assign_class(arr, range_, classes):
    """evaluation_vit.py:31-32 (default=0; dataset.py:8-9 is the default=-1 twin used for the loader's truth classes)"""
    return np.select([np.logical_and(arr > r[0], arr <= r[1]) for r in range_], classes, default=0)


def new_state(output_dim: int) -> dict:
    s = {k: 0.0 for k in (
        "val_loss_sum", "val_acc", "valid_entry_cnt", "valid_nonzero_entry_cnt",
        "valid_diff_sum", "valid_diff_sum_p", "valid_diff_sum_sim_21h", "valid_diff_sum_sim_avg",
        "valid_diff_squares_sum", "valid_diff_squares_sum_p", "valid_diff_squares_sum_sim_21h", "valid_diff_squares_sum_sim_avg",
        "valid_norm_diff_sum", "valid_norm_diff_sum_p", "valid_norm_diff_sum_sim_21h",
        "valid_norm_diff_abs_sum", "valid_norm_diff_abs_sum_p", "valid_norm_diff_abs_sum_sim_21h")}
    for pre in ("", "per_", "sim_21h_", "sim_avg_"):
        for a in "abcd":
            for b in "1234":
                s[f"{pre}total_{a}{b}"] = 0.0
    for k in ("_TP1", "_TN1", "_FP1", "_FN1", "_TP2", "_TN2", "_FP2", "_FN2", "_TP3", "_TN3", "_FP3", "_FN3"):
        s[k] = 0.0
    for pre in ("_", "p_", "sim_21h_", "sim_avg_"):
        for k in ("TP", "TN", "FP", "FN"):
            s[pre + k] = np.zeros(3 * output_dim)
        s[pre + "RMSE_np"] = np.zeros(3 * output_dim)
        s[pre + "MAE_np"] = np.zeros(3 * output_dim)
    s["valid_count"] = np.zeros(3 * output_dim)
    for k in ("valid_vals_gt", "valid_vals_model", "valid_vals_p", "valid_vals_sim_21h", "valid_vals_sim_avg"):
        s[k] = []
    return s


def update(s: dict, preds, pred_vals, pred_classes, last_PM, sim_21h_pm_vals, sim_avg_pm_vals) -> None:
    """One iteration of the loop body, from ``preds[preds < 0.] = 0.`` (:254) on.  All tensors (B, output_dim, P) except
    last_PM (B, P); pred_classes integer with -1 = unlabeled.  ``preds`` is clamped in place, as in the reference."""
    output_dim = preds.shape[1]
    preds[preds < 0.] = 0.                                                                     # :254
    last_PM = last_PM.reshape(preds.shape[0], 1, -1).repeat(1, output_dim, 1)                  # :241-243
    persistent_PM = torch.from_numpy(assign_class(last_PM.numpy(), RANGE_4CLASS, CLASS_FOUR))   # :245
    labels = torch.from_numpy(assign_class(preds.numpy(), RANGE_4CLASS, CLASS_FOUR))            # :259-261
    sim_21h_labels = torch.from_numpy(assign_class(sim_21h_pm_vals.numpy(), RANGE_4CLASS, CLASS_FOUR))   # :278-280
    sim_avg_labels = torch.from_numpy(assign_class(sim_avg_pm_vals.numpy(), RANGE_4CLASS, CLASS_FOUR))   # :281-283

    s["val_loss_sum"] += torch.nn.functional.mse_loss(preds, pred_vals).item()                 # :291 (criterion :140)
    for suf, vals in (("", preds), ("_p", last_PM), ("_sim_21h", sim_21h_pm_vals), ("_sim_avg", sim_avg_pm_vals)):   # :293-307
        d = vals - pred_vals
        s["valid_diff_sum" + suf] += torch.abs(d).sum().item()
        s["valid_diff_squares_sum" + suf] += (d ** 2).sum().item()
    s["valid_entry_cnt"] += torch.numel(preds)                                                 # :309
    nonzero_mask = (pred_vals > 0)                                                             # :311
    s["valid_nonzero_entry_cnt"] += nonzero_mask.sum().item()
    for suf, vals in (("", preds), ("_p", last_PM), ("_sim_21h", sim_21h_pm_vals)):            # :314-326
        nd = vals[nonzero_mask] - pred_vals[nonzero_mask]
        s["valid_norm_diff_sum" + suf] += (nd / pred_vals[nonzero_mask]).sum().item()
        s["valid_norm_diff_abs_sum" + suf] += torch.abs(nd / pred_vals[nonzero_mask]).sum().item()
    s["valid_vals_gt"] += pred_vals.tolist()                                                   # :328-332
    s["valid_vals_model"] += preds.tolist()
    s["valid_vals_p"] += last_PM.tolist()
    s["valid_vals_sim_21h"] += sim_21h_pm_vals.tolist()
    s["valid_vals_sim_avg"] += sim_avg_pm_vals.tolist()

    cur_preds = pred_classes                                                                   # :334-338 (sic: truth classes)
    s["val_acc"] += ((labels == cur_preds)).float().sum().item()                               # :340
    for pre, cur in (("", labels), ("per_", persistent_PM), ("sim_21h_", sim_21h_labels), ("sim_avg_", sim_avg_labels)):   # :345-415
        for ai, a in enumerate("abcd"):
            for bi, b in enumerate("1234"):
                s[f"{pre}total_{a}{b}"] += ((cur == ai) & (cur_preds == bi)).float().sum().item()
    cur_labels = labels
    s["_TP1"] += ((cur_labels > 0) & (cur_preds > 0)).float().sum().item()                     # :418-429
    s["_TN1"] += ((cur_labels == 0) & (cur_preds == 0)).float().sum().item()
    s["_FP1"] += ((cur_labels > 0) & (cur_preds == 0)).float().sum().item()
    s["_FN1"] += ((cur_labels == 0) & (cur_preds > 0)).float().sum().item()
    s["_TP2"] += ((cur_labels > 1) & (cur_preds > 1)).float().sum().item()
    s["_TN2"] += ((cur_labels < 2) & (cur_preds < 2)).float().sum().item()
    s["_FP2"] += ((cur_labels > 1) & (cur_preds < 2)).float().sum().item()
    s["_FN2"] += ((cur_labels < 2) & (cur_preds > 1)).float().sum().item()
    s["_TP3"] += ((cur_labels > 2) & (cur_preds > 2)).float().sum().item()
    s["_TN3"] += ((cur_labels < 3) & (cur_preds < 3)).float().sum().item()
    s["_FP3"] += ((cur_labels > 2) & (cur_preds < 3)).float().sum().item()
    s["_FN3"] += ((cur_labels < 3) & (cur_preds > 2)).float().sum().item()

    for i in range(1, 3 + 1):                                                                  # :432-463
        for j in range(output_dim):
            cp = pred_classes[:, j]
            k = (i - 1) * output_dim + j
            for pre, lab in (("_", labels), ("p_", persistent_PM), ("sim_21h_", sim_21h_labels), ("sim_avg_", sim_avg_labels)):
                cl = lab[:, j]
                s[pre + "TP"][k] += ((cl > i - 1) & (cp > i - 1)).sum().item()
                s[pre + "TN"][k] += ((cl < i) & (cp < i) & (cp > -1)).sum().item()
                s[pre + "FP"][k] += ((cl > i - 1) & (cp < i) & (cp > -1)).sum().item()
                s[pre + "FN"][k] += ((cl < i) & (cp > i - 1)).sum().item()
            sel = pred_classes[:, j] > i - 1
            for pre, vals in (("_", preds), ("p_", last_PM), ("sim_21h_", sim_21h_pm_vals), ("sim_avg_", sim_avg_pm_vals)):
                s[pre + "RMSE_np"][k] += ((vals[:, j][sel] - pred_vals[:, j][sel]) ** 2).sum().item()
                s[pre + "MAE_np"][k] += torch.abs(vals[:, j][sel] - pred_vals[:, j][sel]).sum().item()
            s["valid_count"][k] += sel.sum().item()


def finish(s: dict) -> dict:
    """post-loop scalars of :507-523 and :572-576 (normalised mean bias / error in percent, Pearson r)"""
    def flat(key):      # np.array(list) in the reference (same P every batch); flattened here so ragged test batches work too
        return np.concatenate([np.asarray(e, dtype=np.float64).ravel() for e in s[key]])
    gt = flat("valid_vals_gt")
    out = {}
    for suf, key in (("", "valid_vals_model"), ("_p", "valid_vals_p"), ("_sim_21h", "valid_vals_sim_21h"), ("_sim_avg", "valid_vals_sim_avg")):
        v = flat(key)
        out["nmb" + suf] = np.sum(v - gt) / np.sum(gt) * 100
        out["nme" + suf] = np.sum(np.abs(v - gt)) / np.sum(gt) * 100
        vc, gc = v - np.mean(v), gt - np.mean(gt)
        out["r" + suf] = np.sum(vc * gc) / (np.sqrt(np.sum(vc ** 2)) * np.sqrt(np.sum(gc ** 2)))
    return out
