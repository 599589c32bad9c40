"""CPU restatement of the reference's evaluation metrics (row N3) -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/evaluation.py:
  eval_translation :16-23   xy distance, thresholds 0.02 / 0.1 / 0.2 m (strict <)
  angle_diff       :26-28   ((b - a + pi) mod 2 pi) - pi   (floor-mod)
  eval_angle       :31-40   degrees, optional min with the 180-degree flipped prediction, thresholds 1 / 5 / 10
  eval_transform   :43-46   element-wise min of both level vectors
  evaluate         :128-211 centre-of-rotation correction (pointcloud.py:309-318), accumulation over the sets
                            {both, val, test} x ranges {all, 5m, 10m, 15m, 20m} of |gt_pc1center|, final means / RMS
  output schema    :229-273 (the dict written to eval.json)
Pinned to the reference's own evaluate() executed on seeded inputs: tests/golden/reference_eval.json.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from . import rigid

RANGES = ("all", "5m", "10m", "15m", "20m")
RANGE_LIMIT = (np.inf, 5.0, 10.0, 15.0, 20.0)
FIELDS = ("num", "corr_levels_translation", "mean_dist_translation", "mean_sq_dist_translation", "corr_levels_angles",
          "mean_dist_angle", "mean_sq_dist_angle", "corr_levels")


def angle_diff(a, b):
    return np.mod(b - a + np.pi, 2.0 * np.pi) - np.pi


def per_transform(pred_t, pred_a, pred_c, gt_t, gt_a, gt_c1, accept_inverted_angle: bool):
    """Vectorised per-sample quantities: corrected translation, distances, level vectors [n,3]."""
    pred_a = np.asarray(pred_a, np.float64).reshape(-1)
    gt_a = np.asarray(gt_a, np.float64).reshape(-1)
    t_new = rigid.translate_transform_to_new_center_of_rotation(pred_t, pred_a[:, None], pred_c, gt_c1)
    dist_t = np.linalg.norm(t_new[:, :2] - np.asarray(gt_t, np.float64)[:, :2], axis=1)
    dist_a = np.abs(angle_diff(pred_a, gt_a)) / np.pi * 180.0
    if accept_inverted_angle:
        dist_a = np.minimum(dist_a, np.abs(angle_diff(pred_a + np.pi, gt_a)) / np.pi * 180.0)
    lev_t = (dist_t[:, None] < np.array([0.02, 0.1, 0.2])[None]).astype(np.float64)
    lev_a = (dist_a[:, None] < np.array([1.0, 5.0, 10.0])[None]).astype(np.float64)
    return t_new, dist_t, dist_a, lev_t, lev_a, np.minimum(lev_t, lev_a)


def accumulate(pred_t, pred_a, pred_c, gt_t, gt_a, gt_c1, is_test, accept_inverted_angle: bool) -> np.ndarray:
    """Raw sums [3 sets (both, val, test)][5 ranges][14]: num, 3 translation levels, sum d_t, sum d_t^2,
    3 angle levels, sum d_a, sum d_a^2, 3 joint levels -- what the device kernel produces."""
    _, dist_t, dist_a, lev_t, lev_a, lev = per_transform(pred_t, pred_a, pred_c, gt_t, gt_a, gt_c1, accept_inverted_angle)
    cd = np.linalg.norm(np.asarray(gt_c1, np.float64), axis=1)
    is_test = np.asarray(is_test, bool)
    acc = np.zeros((3, 5, 14))
    row = np.concatenate([np.ones((len(dist_t), 1)), lev_t, dist_t[:, None], dist_t[:, None] ** 2, lev_a, dist_a[:, None],
                          dist_a[:, None] ** 2, lev], axis=1)
    ok = dist_t <= 10000.0
    for s, sel in enumerate((np.ones_like(is_test), ~is_test, is_test)):
        for r, lim in enumerate(RANGE_LIMIT):
            m = ok & sel & ~(cd > lim)
            acc[s, r] = row[m].sum(axis=0)
    return acc


def finalize(acc: np.ndarray, mean_time: float = 0.0) -> Dict:
    """Sums -> the eval.json dictionary of evaluation.py:229-273."""
    def node(v):
        n = v[0] if v[0] != 0 else 1e-20
        return dict(corr_levels=(v[11:14] / n).tolist(), corr_levels_translation=(v[1:4] / n).tolist(),
                    mean_dist_translation=float(v[4] / n), mean_sq_dist_translation=float(np.sqrt(v[5] / n)),
                    corr_levels_angles=(v[6:9] / n).tolist(), mean_dist_angle=float(v[9] / n),
                    mean_sq_dist_angle=float(np.sqrt(v[10] / n)), num=int(v[0]))

    def group(a):
        d = node(a[0])
        for r, key in enumerate(RANGES[1:], start=1):
            d["eval_" + key] = node(a[r])
        return d
    out = group(acc[0])
    out["val"] = group(acc[1])
    out["test"] = group(acc[2])
    out["reg_eval"] = dict(fitness=0.0, inlier_rmse=0.0)
    out["mean_time"] = mean_time
    return out


def evaluate(pred_t, pred_a, pred_c, gt_t, gt_a, gt_c1, is_test, accept_inverted_angle=False, mean_time=0.0) -> Dict:
    return finalize(accumulate(pred_t, pred_a, pred_c, gt_t, gt_a, gt_c1, is_test, accept_inverted_angle), mean_time)
