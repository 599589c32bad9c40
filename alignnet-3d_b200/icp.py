"""Yaw-constrained point-to-point ICP refinement on the device.

Mirrors `icp.icp_p2point(..., with_constraint=True, radius=0.1, init=get_mat_angle(t, angle, centre), its=N)` as the
reference driver calls it per validation pair (/root/reference/icp.py:69-78, train.py:463-484), batched: one CTA per
pair, full (not resampled) clouds.  The estimator of the authors' Open3D fork is not vendored -- parity unpinned; see
oracle/icp_ref.py for the restated algorithm and tests/test_icp.py for the validation on synthetic ground truth."""
from __future__ import annotations

import json
import logging
import os
import time
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib

logger = logging.getLogger("tp")


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


def load_pointclouds(basepath: str, file_idx: int) -> Tuple[np.ndarray, np.ndarray]:
    """icp.py:42-47 (`load_pountclouds`): the FULL clouds of a pair, xyz columns."""
    ps1 = np.load(f"{basepath}/pointcloud1/{str(file_idx).zfill(8)}.npy")[:, :3]
    ps2 = np.load(f"{basepath}/pointcloud2/{str(file_idx).zfill(8)}.npy")[:, :3]
    return ps1, ps2


def get_centroid_init(ps1: np.ndarray, ps2: np.ndarray) -> np.ndarray:
    """icp.py:62-66: identity rotation, translation = difference of the centroids."""
    init = np.eye(4)
    init[:3, 3] = ps2.mean(axis=0) - ps1.mean(axis=0)
    return init


def evaluate(cfg, use_old_results: bool = False, device: str = "cuda:0", chunk: int = 256) -> Dict:
    """`evaluation.special.mode == 'icp'` (train.py:548-551 -> icp.py:150-225) for the variant this engine implements:
    `variant == 'p2point'` with `with_constraint` true (the reference's `icp_<dataset>_o3_p2p.json` configs,
    make_icp_configs.py:7) -- `icp_p2point(file_idx, cfg, radius=0.10)` from the centroid initialisation, 30 iterations,
    over the whole validation split, `chunk` pairs per launch.  Writes what the reference writes: `<logdir>/val/eval000000/
    {pred_translations, pred_angles, pred_s1_pc1centers}.npy` (centres = origin, icp.py:207: the ICP transform rotates
    about the origin) and `eval.json` / `eval_180.json`.  The other variants (global registration of the Open3D fork) are
    rejected by config.validate()."""
    from . import evaluation, provider
    sp = cfg.evaluation.special.icp
    if sp.variant != "p2point" or not bool(sp.with_constraint) or sp.has("refine"):
        raise ValueError("evaluation.special.icp: only variant='p2point' with with_constraint=true (no 'refine') is implemented")
    val_idxs = provider.get_data_files(f"{cfg.data.basepath}/split/val.txt")
    n = len(val_idxs)
    eval_dir = f"{cfg.logging.logdir}/val/eval{str(0).zfill(6)}"             # icp.py:154,176
    meta = [provider.load_meta(cfg.data.basepath, i) for i in val_idxs]       # the labels provider.load_batch returns (icp.py:175)
    gt_t = np.stack([m[0] for m in meta]).astype(np.float32).reshape(n, 3)
    gt_a = np.stack([np.ravel(m[1])[:1] for m in meta]).astype(np.float32).reshape(n, 1)
    gt_c1 = np.stack([m[2] for m in meta]).astype(np.float32).reshape(n, 3)
    total_time = 0.0
    if use_old_results and os.path.isfile(f"{eval_dir}/pred_translations.npy"):          # icp.py:177-180
        pred_t = np.load(f"{eval_dir}/pred_translations.npy")
        pred_a = np.load(f"{eval_dir}/pred_angles.npy")
        pred_c = np.load(f"{eval_dir}/pred_s1_pc1centers.npy")
    else:
        pred_t = np.empty((n, 3), np.float32)
        pred_a = np.empty((n, 1), np.float32)
        pred_c = np.zeros((n, 3), np.float32)                                           # icp.py:207
        for lo in range(0, n, chunk):
            pairs = [load_pointclouds(cfg.data.basepath, i) for i in val_idxs[lo:lo + chunk]]
            inits = np.stack([get_centroid_init(a, b) for a, b in pairs])
            torch.cuda.synchronize()
            t0 = time.time()
            tf, _ = refine([a for a, _ in pairs], [b for _, b in pairs], inits, radius=0.10, its=30, device=device)
            total_time += time.time() - t0
            t, a = to_translation_angle(tf)
            pred_t[lo:lo + len(pairs)], pred_a[lo:lo + len(pairs), 0] = t, a
        os.makedirs(eval_dir, exist_ok=True)
        np.save(f"{eval_dir}/pred_translations.npy", pred_t)
        np.save(f"{eval_dir}/pred_angles.npy", pred_a)
        np.save(f"{eval_dir}/pred_s1_pc1centers.npy", pred_c)
    from .train import _is_test
    is_test = _is_test(cfg, list(val_idxs))
    result = {}
    for inverted in (False, True):                                                       # icp.py:222-224
        d = evaluation.evaluate(pred_t, pred_a, gt_t, gt_a, pred_c, gt_c1, is_test, inverted, total_time / max(n, 1),
                                device=device)
        with open(f'{eval_dir}/eval{"_180" if inverted else ""}.json', "w") as fh:
            json.dump(d, fh)
        logger.info(d)
        result["eval_180" if inverted else "eval"] = d
    return result
