"""NumPy float64 restatement of the z-axis rigid-transform utilities.

Oracle = test infrastructure (see oracle/__init__.py).  Citations into /root/reference.
The Rz sign convention is the one pinned by the reference's only known-answer vectors, the
`euler2mat` doctests (utils/eulerangles.py:152-154): euler2mat(z=pi/2) == [[0,-1,0],[1,0,0],[0,0,1]].
"""
from __future__ import annotations

import math

import numpy as np


def rot_z(theta: float) -> np.ndarray:
    """utils/eulerangles.py:172-178 (z block of euler2mat) == models/tp8.py:26-27 == scipy
    Rotation.from_rotvec([0,0,theta]).as_dcm() used at tp_utils/pointcloud.py:288."""
    c, s = math.cos(theta), math.sin(theta)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def get_mat_angle(translation=None, rotation=None, rotation_center=np.array([0.0, 0.0, 0.0])) -> np.ndarray:
    """tp_utils/pointcloud.py:279-289: 4x4  T(c + t) . Rz(theta) . T(-c)."""
    mat1, mat2, mat3 = np.eye(4), np.eye(4), np.eye(4)
    mat1[:3, 3] = -np.asarray(rotation_center, dtype=np.float64)
    mat3[:3, 3] = np.asarray(rotation_center, dtype=np.float64)
    if translation is not None:
        mat3[:3, 3] += np.asarray(translation, dtype=np.float64)
    if rotation is not None:
        mat2[:3, :3] = rot_z(float(rotation))
    return mat3 @ mat2 @ mat1


def transform_points(ps: np.ndarray, mats) -> np.ndarray:
    """tp_utils/pointcloud.py:292-298: homogeneous rows times M^T; a list composes in order."""
    if isinstance(mats, list):
        ps = np.array(ps, dtype=np.float64, copy=True)
        for mat in mats:
            ps[:, :4] = ps[:, :4] @ mat.T
        return ps
    return ps[:, :4] @ mats.T


def rigid_apply(points: np.ndarray, translation, angle: float, center) -> np.ndarray:
    """p' = Rz(angle) (p - c) + c + t for [n,3] points, via the two functions above."""
    hom = np.concatenate([np.asarray(points, np.float64), np.ones((points.shape[0], 1))], axis=1)
    return transform_points(hom, get_mat_angle(translation, angle, center))[:, :3]


def translate_transform_to_new_center_of_rotation(pred_translations, pred_angles, pred_centers, gt_pc1centers):
    """tp_utils/pointcloud.py:309-318: t' = -d + Rz(theta) d + t,  d = c_new - c_old."""
    out = np.zeros_like(np.asarray(pred_translations, np.float64))
    for i, (t, a, c_old, c_new) in enumerate(zip(pred_translations, pred_angles, pred_centers, gt_pc1centers)):
        shift = np.asarray(c_new, np.float64) - np.asarray(c_old, np.float64)
        out[i] = -shift + get_mat_angle(rotation=float(np.asarray(a).reshape(-1)[0]))[:3, :3] @ shift + t
    return out


def tf_transform_pcs(pcs, translations=None, angles=None, rotation_centers=None) -> np.ndarray:
    """a21: models/tp8.py:361-371 statement by statement, quirk Q6 included: `tf_translate_pcs` (:357-358) returns
    tile(translation) instead of pcs + translation, so every translate step REPLACES the cloud; the rotation is
    `tf.matmul(pcs, R)` with R = tf_get_rotation_matrix_z(a) = [[c,-s,0],[s,c,0],[0,0,1]] (:26-27).  [B,N,3] float64."""
    p = np.asarray(pcs, np.float64).copy()
    tile = lambda t: np.repeat(np.asarray(t, np.float64)[:, None, :], p.shape[1], axis=1)
    if rotation_centers is not None:
        p = tile(-np.asarray(rotation_centers, np.float64))
    if angles is not None:
        for b, a in enumerate(np.asarray(angles, np.float64).reshape(-1)):
            c, s = math.cos(a), math.sin(a)
            p[b] = p[b] @ np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    if translations is not None:
        p = tile(-np.asarray(translations, np.float64))
    if rotation_centers is not None:
        p = tile(np.asarray(rotation_centers, np.float64))
    return p


def loss_p2p(pcs1, pred_translations, pred_angles, pred_s2_pc1centers, translations, rel_angles, pc1_centers):
    """a21: models/tp8.py:383-397 statement by statement on top of `tf_transform_pcs`: (per_transform_loss, loss)."""
    a = tf_transform_pcs(pcs1, pred_translations, pred_angles, pred_s2_pc1centers)
    g = tf_transform_pcs(pcs1, translations, np.asarray(rel_angles).reshape(len(a), -1)[:, 0], pc1_centers)
    point_distances = np.linalg.norm(a - g, axis=1)          # tf.norm(..., axis=1): over the POINT axis -> [B,3]
    loss = float(np.mean(np.square(point_distances)))
    return loss / a.shape[0], loss
