"""Evaluation metrics of the reference on the device.

Mirrors `evaluation.evaluate` (/root/reference/evaluation.py:128-273): same argument meaning, same dictionary as the
one the reference writes to `eval.json` (`ns_to_dict(eval_dict)`).  The per-transform work (centre-of-rotation
correction, errors, threshold levels, set / range bucketing) and the reductions run in one CUDA kernel
(`an3d_evaluate`); the final means and RMS values are a 210-element division on the host.  The val/test split of the
reference (KITTI track ids or idx >= 1000 for the synthetic sets, evaluation.py:159-162) is passed in as `is_test`.
Track / velocity files (evaluation.py:49-118, 212-225) are outside this row."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib

RANGES = ("5m", "10m", "15m", "20m")


def _dev(x, device, cols: Optional[int]) -> torch.Tensor:
    t = torch.as_tensor(np.asarray(x.detach().cpu() if isinstance(x, torch.Tensor) else x, dtype=np.float64))
    t = t.reshape(-1, cols) if cols else t.reshape(-1)
    return t.contiguous().to(device)


def accumulate(all_pred_translations, all_pred_angles, all_gt_translations, all_gt_angles, all_pred_centers,
               all_gt_pc1centers, is_test=None, accept_inverted_angle: bool = False, device="cuda:0") -> np.ndarray:
    """Raw sums [3 sets (all, val, test)][5 ranges (all, 5, 10, 15, 20 m)][14] from the device kernel."""
    lib = _lib.load()
    pt, pc = _dev(all_pred_translations, device, 3), _dev(all_pred_centers, device, 3)
    gt, gc = _dev(all_gt_translations, device, 3), _dev(all_gt_pc1centers, device, 3)
    pa, ga = _dev(all_pred_angles, device, None), _dev(all_gt_angles, device, None)
    n = int(pt.shape[0])
    if not (pc.shape[0] == gt.shape[0] == gc.shape[0] == pa.shape[0] == ga.shape[0] == n):
        raise ValueError("evaluate: all arrays must describe the same number of transforms")
    flags = None
    if is_test is not None:
        flags = torch.as_tensor(np.asarray(is_test, dtype=np.uint8)).contiguous().to(device)
        if flags.numel() != n:
            raise ValueError("evaluate: is_test must have one entry per transform")
    if n == 0:                       # nothing to reduce (empty tensors have no device address to pass)
        return np.zeros((3, 5, 14))
    acc = torch.empty(210, dtype=torch.float64, device=device)
    stream = torch.cuda.current_stream(torch.device(device)).cuda_stream
    _lib.check(lib.an3d_evaluate(pt.data_ptr(), pa.data_ptr(), pc.data_ptr(), gt.data_ptr(), ga.data_ptr(), gc.data_ptr(),
                                 flags.data_ptr() if flags is not None else None, n, 1 if accept_inverted_angle else 0,
                                 acc.data_ptr(), stream), "an3d_evaluate")
    return acc.cpu().numpy().reshape(3, 5, 14)


def _node(v: np.ndarray) -> Dict:
    n = v[0] if v[0] != 0 else 1e-20        # evaluation.py:196-198: an empty bucket yields huge numbers on purpose
    return dict(corr_levels=(v[11:14] / n).tolist(), corr_levels_translation=(v[1:4] / n).tolist(),
                mean_dist_translation=float(v[4] / n), mean_sq_dist_translation=float(np.sqrt(v[5] / n)),
                corr_levels_angles=(v[6:9] / n).tolist(), mean_dist_angle=float(v[9] / n),
                mean_sq_dist_angle=float(np.sqrt(v[10] / n)), num=int(v[0]))


def _group(a: np.ndarray) -> Dict:
    d = _node(a[0])
    for r, key in enumerate(RANGES, start=1):
        d["eval_" + key] = _node(a[r])
    return d


def evaluate(all_pred_translations, all_pred_angles, all_gt_translations, all_gt_angles, all_pred_centers,
             all_gt_pc1centers, is_test=None, accept_inverted_angle: bool = False, mean_time: float = 0.0,
             device="cuda:0") -> Dict:
    """The dictionary `evaluation.evaluate(...)` writes to eval.json (argument order of evaluation.py:128 after
    `cfg, val_idxs`, which only serve the val/test split there)."""
    acc = accumulate(all_pred_translations, all_pred_angles, all_gt_translations, all_gt_angles, all_pred_centers,
                     all_gt_pc1centers, is_test, accept_inverted_angle, device)
    out = _group(acc[0])
    out["val"] = _group(acc[1])
    out["test"] = _group(acc[2])
    out["reg_eval"] = dict(fitness=0.0, inlier_rmse=0.0)
    out["mean_time"] = mean_time
    return out
