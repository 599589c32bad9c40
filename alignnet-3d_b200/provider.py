"""Data provider for the reference's on-disk format.

Mirrors `provider.load_batch` / `load_from_separate_files` / `jitter_point_cloud` (/root/reference/provider.py:60-71,
85-136) and the text-encoded arrays of `tp_utils/pointcloud.py:247-265` (`np.savetxt` strings inside the meta JSON):

    <basepath>/meta/%08d.json        translation, rel_angle, start_position, end_position, start_angle, end_angle
    <basepath>/pointcloud{1,2}/%08d.npy   ragged [n_i, >=3] clouds
    <basepath>/split/{train,val}.txt      one example index per line

File IO and the random draws stay on the host -- the draws use numpy's legacy global RNG with the reference's exact
call order (`np.random.choice(n, num_points, replace=True)` per cloud, `np.random.randn(B, N, 3)` per jitter), so a
seeded run selects the same points as the reference -- while the batch itself is assembled on the device: the ragged
clouds are uploaded once, `an3d_resample_gather` resamples them to [B, N, 3] and adds the clipped jitter.
`Prefetcher` overlaps the host side of the next batch with the device work on the current one."""
from __future__ import annotations

import io
import json
import os
import threading
from queue import Queue
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _lib

LABEL_KEYS = ("translations", "rel_angles", "pc1_centers", "pc2_centers", "pc1_angles", "pc2_angles")


def str_to_np(s: str) -> np.ndarray:
    """pointcloud.py:260-265 (plaintext branch): array stored as the text np.savetxt wrote."""
    return np.loadtxt(io.BytesIO(s.encode("ascii")))


def np_to_str(arr) -> str:
    """pointcloud.py:247-252 (plaintext branch): the text np.savetxt writes; `str_to_np` reads it back exactly (%.18e)."""
    out = io.BytesIO()
    np.savetxt(out, np.asarray(arr, np.float64))
    return out.getvalue().decode("ascii")


def save_example(basepath: str, idx: int, cloud1: np.ndarray, cloud2: np.ndarray, start_position, start_angle: float,
                 end_position, end_angle: float, translation, rel_angle: float, additional_meta: Optional[dict] = None) -> None:
    """The writer half of the on-disk format (Scene.save_pointclouds / save_meta, pointcloud.py:979-997):
    `pointcloud{1,2}/%08d.npy` and `meta/%08d.json` with the text-encoded arrays.  Creates the directories."""
    for d in ("meta", "pointcloud1", "pointcloud2", "split"):
        os.makedirs(os.path.join(basepath, d), exist_ok=True)
    name = str(idx).zfill(8)
    np.save(f"{basepath}/pointcloud1/{name}", np.asarray(cloud1))
    np.save(f"{basepath}/pointcloud2/{name}", np.asarray(cloud2))
    data = {"start_position": np_to_str(start_position), "start_angle": float(start_angle),
            "end_position": np_to_str(end_position), "end_angle": float(end_angle),
            "translation": np_to_str(translation), "rel_angle": float(rel_angle), **(additional_meta or {})}
    with open(f"{basepath}/meta/{name}.json", "w") as fh:
        json.dump(data, fh)


def save_split(basepath: str, name: str, indices: Sequence[int]) -> None:
    """`split/{train,val}.txt`: one example index per line (read by get_data_files, provider.py:74-75)."""
    os.makedirs(os.path.join(basepath, "split"), exist_ok=True)
    with open(f"{basepath}/split/{name}.txt", "w") as fh:
        fh.write("".join(f"{int(i)}\n" for i in indices))


def get_data_files(list_filename: str) -> List[int]:
    """provider.py:74-75."""
    return [int(line.rstrip()) for line in open(list_filename)]


def load_meta(basepath: str, idx: int):
    """The label part of provider.py:85-90."""
    data = json.load(open(f"{basepath}/meta/{str(idx).zfill(8)}.json", "r"))
    return (str_to_np(data["translation"]), data["rel_angle"], str_to_np(data["start_position"]),
            str_to_np(data["end_position"]), data["start_angle"], data["end_angle"])


def read_host_batch(basepath: str, indices: Sequence[int], num_points: int, jitter: bool = False,
                    jitter_sigma: float = 0.01, jitter_clip: float = 0.05) -> Dict[str, np.ndarray]:
    """Host half of load_batch: labels, the ragged clouds back to back, and the random draws in the reference's
    order (per example: choice for cloud 1, choice for cloud 2 -- provider.py:97-98; then, for training,
    randn for pcs1 and randn for pcs2 -- train.py:355-356)."""
    B = len(indices)
    labels = {k: np.empty((B, 3 if k in ("translations", "pc1_centers", "pc2_centers") else 1)) for k in LABEL_KEYS}
    clouds, offsets, sample_idx = [[], []], [np.zeros(B, np.int64), np.zeros(B, np.int64)], \
        [np.empty((B, num_points), np.int32), np.empty((B, num_points), np.int32)]
    rows = [0, 0]
    for i, ex in enumerate(indices):
        t, ra, c1, c2, a1, a2 = load_meta(basepath, ex)
        labels["translations"][i], labels["rel_angles"][i] = t, ra
        labels["pc1_centers"][i], labels["pc2_centers"][i] = c1, c2
        labels["pc1_angles"][i], labels["pc2_angles"][i] = a1, a2
        for w in (0, 1):
            pc = np.load(f"{basepath}/pointcloud{w + 1}/{str(ex).zfill(8)}.npy")
            offsets[w][i] = rows[w]
            if pc.shape[0] > 0:
                sample_idx[w][i] = np.random.choice(pc.shape[0], num_points, replace=True)
                clouds[w].append(np.ascontiguousarray(pc[:, :3], dtype=np.float32))
                rows[w] += pc.shape[0]
            else:                              # provider.py:97: an empty cloud becomes zeros (and draws nothing)
                sample_idx[w][i] = -1
    out = dict(labels)
    for w in (0, 1):
        out[f"points{w + 1}"] = np.concatenate(clouds[w], axis=0) if clouds[w] else np.zeros((0, 3), np.float32)
        out[f"offsets{w + 1}"] = offsets[w]
        out[f"sample_idx{w + 1}"] = sample_idx[w]
    if jitter:
        for w in (0, 1):                        # provider.py:60-71
            out[f"jitter{w + 1}"] = np.clip(jitter_sigma * np.random.randn(B, num_points, 3), -jitter_clip,
                                            jitter_clip).astype(np.float32)
    return out


def assemble_on_device(host: Dict[str, np.ndarray], device="cuda:0") -> Dict[str, torch.Tensor]:
    """Device half: upload the ragged clouds and draws, gather / jitter into the 8 feeds of tp8.placeholder_inputs."""
    lib = _lib.load()
    dev = torch.device(device)
    stream = torch.cuda.current_stream(dev).cuda_stream
    batch: Dict[str, torch.Tensor] = {}
    for w in (1, 2):
        pts = torch.from_numpy(host[f"points{w}"]).to(dev)
        off = torch.from_numpy(host[f"offsets{w}"]).to(dev)
        idx = torch.from_numpy(host[f"sample_idx{w}"]).to(dev)
        B, N = idx.shape
        jit = torch.from_numpy(host[f"jitter{w}"]).to(dev) if f"jitter{w}" in host else None
        out = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
        _lib.check(lib.an3d_resample_gather(pts.data_ptr() if pts.numel() else None, off.data_ptr(), idx.data_ptr(), B, N, 3,
                                            jit.data_ptr() if jit is not None else None, out.data_ptr(), stream),
                   "an3d_resample_gather")
        batch[f"pcs{w}"] = out
    for k in LABEL_KEYS:
        batch[k] = torch.from_numpy(np.ascontiguousarray(host[k], dtype=np.float32)).to(dev)
    return batch


def load_batch(basepath: str, indices: Sequence[int], num_points: int, jitter: bool = False, device="cuda:0"):
    """provider.load_batch (+ the jitter of train.py:355-356) -> dict of CUDA tensors keyed like the engine's batch."""
    return assemble_on_device(read_host_batch(basepath, indices, num_points, jitter), device)


class Prefetcher:
    """Iterates over batches of `indices`; a worker thread reads and draws batch k+1 (host) while batch k is in use.
    The worker is the only consumer of numpy's global RNG while it runs, so the draw order stays the reference's."""

    def __init__(self, basepath: str, indices: Sequence[int], batch_size: int, num_points: int, jitter: bool = False,
                 device="cuda:0", depth: int = 2):
        self.args = (basepath, num_points, jitter)
        self.device = device
        self.batches = [list(indices[i:i + batch_size]) for i in range(0, len(indices) - batch_size + 1, batch_size)]
        self.queue: Queue = Queue(maxsize=depth)
        self.thread = threading.Thread(target=self._work, daemon=True)
        self.thread.start()

    def _work(self):
        basepath, num_points, jitter = self.args
        try:
            for b in self.batches:
                self.queue.put(read_host_batch(basepath, b, num_points, jitter))
        except Exception as exc:            # surface IO errors in the consumer
            self.queue.put(exc)
        self.queue.put(None)

    def __iter__(self) -> Iterable[Dict[str, torch.Tensor]]:
        while True:
            item = self.queue.get()
            if item is None:
                return
            if isinstance(item, Exception):
                raise item
            yield assemble_on_device(item, self.device)
