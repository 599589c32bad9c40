"""Yaw-constrained point-to-point ICP refinement on the device.

Mirrors `icp.icp_p2point(..., with_constraint=True, radius=0.1, init=get_mat_angle(t, angle, centre), its=N)` as the
reference driver calls it per validation pair (/root/reference/icp.py:69-78, train.py:463-484), batched: one CTA per
pair, full (not resampled) clouds.  The estimator of the authors' Open3D fork is not vendored -- parity unpinned; see
oracle/icp_ref.py for the restated algorithm and tests/test_icp.py for the validation on synthetic ground truth."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def get_mat_angle(translation, angle: float, center) -> np.ndarray:
    """pointcloud.py:279-289: T(c + t) Rz(angle) T(-c)."""
    c, s = np.cos(angle), np.sin(angle)
    R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    m = np.eye(4)
    m[:3, :3] = R
    m[:3, 3] = np.asarray(center, float) + np.asarray(translation, float) - R @ np.asarray(center, float)
    return m


def refine(sources: Sequence[np.ndarray], targets: Sequence[np.ndarray], inits: np.ndarray, radius: float = 0.1, its: int = 30,
           device="cuda:0") -> Tuple[np.ndarray, np.ndarray]:
    """ICP for a batch of pairs.  sources[i] / targets[i]: [n_i, >=3] clouds; inits: [pairs,4,4].
    Returns (transforms [pairs,4,4] float64, stats [pairs,3] = fitness, inlier RMSE, iterations)."""
    lib = _lib.load()
    dev = torch.device(device)
    pairs = len(sources)
    if pairs == 0:
        return np.zeros((0, 4, 4)), np.zeros((0, 3))

    def pack(clouds: Sequence[np.ndarray]):
        counts = np.array([len(c) for c in clouds], np.int32)
        offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
        nz = [np.ascontiguousarray(np.asarray(c)[:, :3], dtype=np.float32) for c in clouds if len(c)]
        flat = np.concatenate(nz, axis=0) if nz else np.zeros((1, 3), np.float32)
        return torch.from_numpy(flat).to(dev), torch.from_numpy(offs).to(dev), torch.from_numpy(counts).to(dev)

    s, so, sn = pack(sources)
    t, to, tn = pack(targets)
    init = torch.from_numpy(np.ascontiguousarray(np.asarray(inits, np.float32).reshape(pairs, 16))).to(dev)
    out = torch.empty((pairs, 16), dtype=torch.float32, device=dev)
    stats = torch.empty((pairs, 3), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.an3d_icp_yaw(s.data_ptr(), so.data_ptr(), sn.data_ptr(), t.data_ptr(), to.data_ptr(), tn.data_ptr(),
                                init.data_ptr(), pairs, float(radius), int(its), out.data_ptr(), stats.data_ptr(), stream),
               "an3d_icp_yaw")
    return out.cpu().numpy().astype(np.float64).reshape(pairs, 4, 4), stats.cpu().numpy().astype(np.float64)


def to_translation_angle(transforms: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """train.py:470-480: the refined transform is in world space (rotation about the origin): translation = T[:3,3],
    yaw = atan2(T[1,0], T[0,0]); the stored centre of rotation becomes (0,0,0)."""
    return transforms[:, :3, 3].copy(), np.arctan2(transforms[:, 1, 0], transforms[:, 0, 0])
