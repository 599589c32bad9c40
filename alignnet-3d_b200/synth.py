"""Seeded synthetic point-cloud pairs shaped like the reference's SynthCars / SynthCarsPersons data.

The reference's datasets are external downloads (README.md:39-47), so benchmarks and tests use
this generator.  Distributions are read from the reference's dataset author code
(tp_utils/pointcloud.py): RandomTransform (:520-541) for start pose / motion, SyntheticScene
(:1056,1064) for object scale, range noise (:1134-1135); resampling with replacement to N points
follows provider.py:97 and the optional jitter provider.py:60-71.  Objects are boxes seen from
the sensor at the origin (only the faces turned towards the sensor are sampled), which gives
the partial-view, yaw-ambiguous clouds the network is built for.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def _box_partial_view(rng, n, length, width, height, center, yaw, noise_sigma):
    """Sample n points on the vertical faces of an oriented box that face the origin."""
    c, s = np.cos(yaw), np.sin(yaw)
    # outward normals of the four vertical faces in the object frame, half extents, tangent extents
    normals = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]], np.float64)
    half = np.array([length / 2, length / 2, width / 2, width / 2])
    tang = np.array([width, width, length, length])
    R = np.array([[c, -s], [s, c]])
    wn = normals @ R.T                                   # world-frame normals
    face_centers = center[None, :2] + wn * half[:, None]
    visible = np.einsum("ij,ij->i", wn, -face_centers) > 0
    if not visible.any():
        visible[:] = True
    area = tang * height * visible
    face = rng.choice(4, size=n, p=area / area.sum())
    u = rng.uniform(-0.5, 0.5, size=n) * tang[face]
    v = rng.uniform(-0.5, 0.5, size=n) * height
    local = normals[face] * half[face, None] + np.stack([-normals[face][:, 1], normals[face][:, 0]], 1) * u[:, None]
    xy = local @ R.T + center[None, :2]
    pts = np.concatenate([xy, (center[2] + v)[:, None]], axis=1)
    # range noise along the viewing ray (pointcloud.py:1134-1135)
    rngs = np.linalg.norm(pts, axis=1, keepdims=True)
    noise = np.clip(rng.normal(0.0, noise_sigma, size=(n, 1)), -0.05, 0.05)
    return pts * (1.0 + noise / np.maximum(rngs, 1e-6))


def make_batch(batch_size: int, num_points: int, seed: int = 1234, persons_prob: float = 0.0,
               jitter: bool = False, dtype=np.float32) -> Dict[str, np.ndarray]:
    """Returns the eight arrays of models/tp8.py:13-23 placeholder_inputs, keyed
    pcs1, pcs2, translations, rel_angles, pc1_centers, pc2_centers, pc1_angles, pc2_angles."""
    rng = np.random.Generator(np.random.PCG64(seed))
    B, N = batch_size, num_points
    out = {
        "pcs1": np.zeros((B, N, 3), np.float64), "pcs2": np.zeros((B, N, 3), np.float64),
        "translations": np.zeros((B, 3), np.float64), "rel_angles": np.zeros((B, 1), np.float64),
        "pc1_centers": np.zeros((B, 3), np.float64), "pc2_centers": np.zeros((B, 3), np.float64),
        "pc1_angles": np.zeros((B, 1), np.float64), "pc2_angles": np.zeros((B, 1), np.float64),
    }
    for b in range(B):
        if rng.uniform() < persons_prob:
            height = rng.uniform(1.6, 2.0)
            length, width = 0.3 * height, 0.3 * height
        else:
            length = 6.0
            width = length * rng.uniform(0.35, 0.45)
            height = length * rng.uniform(0.25, 0.35)
        start_angle = rng.uniform(-np.pi, np.pi)
        r, phi = rng.uniform(4.0, 20.0), rng.uniform(-np.pi, np.pi)
        start = np.array([r * np.sin(phi), r * np.cos(phi), 0.0])
        v, psi = rng.uniform(0.0, 1.0), rng.uniform(-np.pi, np.pi)
        trans = np.array([v * np.sin(psi), v * np.cos(psi), 0.0])
        rel = rng.uniform(-np.pi, np.pi) / 2.0
        end, end_angle = start + trans, start_angle + rel
        sigma = max(0.005, 0.05 * r / 80.0)
        out["pcs1"][b] = _box_partial_view(rng, N, length, width, height, start, start_angle, sigma)
        out["pcs2"][b] = _box_partial_view(rng, N, length, width, height, end, end_angle, sigma)
        out["translations"][b], out["rel_angles"][b, 0] = trans, rel
        out["pc1_centers"][b], out["pc2_centers"][b] = start, end
        out["pc1_angles"][b, 0], out["pc2_angles"][b, 0] = start_angle, end_angle
    if jitter:  # provider.py:60-71
        for k in ("pcs1", "pcs2"):
            out[k] += np.clip(0.01 * rng.standard_normal(out[k].shape), -0.05, 0.05)
    return {k: v.astype(dtype) for k, v in out.items()}


def make_batch_fast(batch_size: int, num_points: int, seed: int = 1234, dtype=np.float32) -> Dict[str, np.ndarray]:
    """Vectorised generator for large benchmark batches (same distributions for pose/motion; the
    object is an oriented box whose two sensor-facing vertical faces are sampled uniformly)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    B, N = batch_size, num_points
    length = np.full(B, 6.0)
    width = length * rng.uniform(0.35, 0.45, B)
    height = length * rng.uniform(0.25, 0.35, B)
    a1 = rng.uniform(-np.pi, np.pi, B)
    r, phi = rng.uniform(4.0, 20.0, B), rng.uniform(-np.pi, np.pi, B)
    start = np.stack([r * np.sin(phi), r * np.cos(phi), np.zeros(B)], 1)
    v, psi = rng.uniform(0.0, 1.0, B), rng.uniform(-np.pi, np.pi, B)
    trans = np.stack([v * np.sin(psi), v * np.cos(psi), np.zeros(B)], 1)
    rel = rng.uniform(-np.pi, np.pi, B) / 2.0
    end, a2 = start + trans, a1 + rel

    def cloud(center, yaw):
        c, s = np.cos(yaw), np.sin(yaw)
        to_sensor = -center[:, :2]
        lx = c * to_sensor[:, 0] + s * to_sensor[:, 1]          # sensor direction in the object frame
        ly = -s * to_sensor[:, 0] + c * to_sensor[:, 1]
        sx, sy = np.sign(lx) + (lx == 0), np.sign(ly) + (ly == 0)
        area_x, area_y = width * height, length * height         # faces with normal +-x / +-y
        pick_x = rng.uniform(size=(B, N)) < (area_x / (area_x + area_y))[:, None]
        u = rng.uniform(-0.5, 0.5, (B, N))
        ox = np.where(pick_x, (sx * length / 2)[:, None], u * length[:, None])
        oy = np.where(pick_x, u * width[:, None], (sy * width / 2)[:, None])
        oz = rng.uniform(-0.5, 0.5, (B, N)) * height[:, None]
        x = c[:, None] * ox - s[:, None] * oy + center[:, 0:1]
        y = s[:, None] * ox + c[:, None] * oy + center[:, 1:2]
        pts = np.stack([x, y, oz + center[:, 2:3]], -1)
        rn = np.linalg.norm(pts, axis=-1, keepdims=True)
        sig = np.maximum(0.005, 0.05 * r / 80.0)[:, None, None]
        noise = np.clip(rng.standard_normal((B, N, 1)) * sig, -0.05, 0.05)
        return pts * (1.0 + noise / np.maximum(rn, 1e-6))

    out = {
        "pcs1": cloud(start, a1), "pcs2": cloud(end, a2), "translations": trans, "rel_angles": rel[:, None],
        "pc1_centers": start, "pc2_centers": end, "pc1_angles": a1[:, None], "pc2_angles": a2[:, None],
    }
    return {k: np.ascontiguousarray(v, dtype=dtype) for k, v in out.items()}


def write_dataset(basepath: str, num_examples: int, seed: int = 1234, points_range=(150, 600), persons_prob: float = 0.0,
                  val_fraction: float = 0.2, max_rel_angle: float = np.pi / 2, max_speed: float = 1.0) -> None:
    """Writes `num_examples` synthetic pairs in the reference's on-disk format (provider.save_example: the layout
    `Scene.save_pointclouds` / `save_meta` produce, pointcloud.py:979-997), with RAGGED full clouds of
    `points_range` points each (the provider resamples them to cfg.model.num_points, provider.py:97), a fourth column
    like the reference's clouds carry, and `split/{train,val}.txt`.  The same distributions as `make_batch`;
    `max_rel_angle` / `max_speed` narrow the motion (e.g. for ICP tests, whose basin of convergence is small)."""
    from . import provider
    rng = np.random.Generator(np.random.PCG64(seed))
    for i in range(num_examples):
        if rng.uniform() < persons_prob:
            height = rng.uniform(1.6, 2.0)
            length, width = 0.3 * height, 0.3 * height
        else:
            length = 6.0
            width = length * rng.uniform(0.35, 0.45)
            height = length * rng.uniform(0.25, 0.35)
        start_angle = rng.uniform(-np.pi, np.pi)
        r, phi = rng.uniform(4.0, 20.0), rng.uniform(-np.pi, np.pi)
        start = np.array([r * np.sin(phi), r * np.cos(phi), 0.0])
        v, psi = rng.uniform(0.0, max_speed), rng.uniform(-np.pi, np.pi)
        trans = np.array([v * np.sin(psi), v * np.cos(psi), 0.0])
        rel = rng.uniform(-1.0, 1.0) * max_rel_angle
        end, end_angle = start + trans, start_angle + rel
        sigma = max(0.005, 0.05 * r / 80.0)
        n1, n2 = (int(rng.integers(points_range[0], points_range[1] + 1)) for _ in range(2))
        clouds = [_box_partial_view(rng, n, length, width, height, c, a, sigma)
                  for n, c, a in ((n1, start, start_angle), (n2, end, end_angle))]
        clouds = [np.concatenate([c, np.ones((len(c), 1))], axis=1).astype(np.float32) for c in clouds]
        provider.save_example(basepath, i, clouds[0], clouds[1], start, start_angle, end, end_angle, trans, rel)
    n_val = max(1, int(round(num_examples * val_fraction)))
    provider.save_split(basepath, "train", range(num_examples - n_val))
    provider.save_split(basepath, "val", range(num_examples - n_val, num_examples))
