"""CPU restatement of the z-constrained point-to-point ICP refinement (row N4) -- TEST INFRASTRUCTURE ONLY.

The reference refines the network's estimate with `o3.registration_icp(pc1, pc2, radius, init,
TransformationEstimationPointToPoint(with_constraint=True, with_scaling=False), ICPConvergenceCriteria(max_iteration=its))`
(/root/reference/icp.py:69-78, called from train.py:463-484 with radius 0.1 and the init built by get_mat_angle).
The `with_constraint` estimator belongs to the authors' Open3D fork, which is not vendored: PARITY UNPINNED.  What is
restated here is Open3D's documented ICP loop -- nearest neighbour of every transformed source point within `radius`,
closed-form update, stop after `its` iterations or when fitness and inlier RMSE both change by less than 1e-6 -- with
the rotation restricted to yaw: for matched pairs (p_i, q_i) with centroids pbar, qbar,
    theta = atan2( sum (px' qy' - py' qx'),  sum (px' qx' + py' qy') ),    t = qbar - Rz(theta) pbar
(the 2-D Kabsch solution; z only translates).  Validated on synthetic ground truth (tests/test_icp.py)."""
from __future__ import annotations

import numpy as np


def rot_z(theta: float) -> np.ndarray:
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def icp_yaw(source: np.ndarray, target: np.ndarray, init: np.ndarray, radius: float = 0.1, its: int = 30,
            rel_fitness: float = 1e-6, rel_rmse: float = 1e-6):
    """Returns (T [4,4], fitness, inlier_rmse, iterations)."""
    src = np.asarray(source, np.float64)[:, :3]
    tgt = np.asarray(target, np.float64)[:, :3]
    T = np.asarray(init, np.float64).copy()
    fitness, rmse, it_done = 0.0, 0.0, 0
    if len(src) == 0 or len(tgt) == 0:
        return T, fitness, rmse, it_done

    def correspond(Tm):
        p = src @ Tm[:3, :3].T + Tm[:3, 3]
        d2 = ((p[:, None, :] - tgt[None, :, :]) ** 2).sum(-1)
        j = d2.argmin(1)
        dmin = d2[np.arange(len(p)), j]
        ok = dmin <= radius * radius
        fit = ok.mean()
        rm = np.sqrt(dmin[ok].mean()) if ok.any() else 0.0
        return p, j, ok, fit, rm

    p, j, ok, fitness, rmse = correspond(T)
    for it in range(its):
        if not ok.any():
            break
        P, Q = p[ok], tgt[j[ok]]
        pb, qb = P.mean(0), Q.mean(0)
        Pc, Qc = P - pb, Q - qb
        theta = np.arctan2((Pc[:, 0] * Qc[:, 1] - Pc[:, 1] * Qc[:, 0]).sum(), (Pc[:, 0] * Qc[:, 0] + Pc[:, 1] * Qc[:, 1]).sum())
        R = rot_z(theta)
        U = np.eye(4)
        U[:3, :3] = R
        U[:3, 3] = qb - R @ pb
        T = U @ T
        it_done = it + 1
        prev_f, prev_r = fitness, rmse
        p, j, ok, fitness, rmse = correspond(T)
        if abs(prev_f - fitness) < rel_fitness and abs(prev_r - rmse) < rel_rmse:
            break
    return T, float(fitness), float(rmse), it_done
