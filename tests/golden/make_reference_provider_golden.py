"""Generates tests/golden/dataset_tiny/ (a six-example dataset in the reference's on-disk format, arrays encoded with
the reference's own `np_to_str`) and tests/golden/reference_provider.npz = the output of the reference's OWN,
unmodified `provider.load_batch` + `provider.jitter_point_cloud` on it with `np.random.seed(77)`.
Run from the repo root in the build container: python tests/golden/make_reference_provider_golden.py"""
import importlib
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_run as RR  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BASE = os.path.join(HERE, "dataset_tiny")
N_POINTS = 64


def main():
    pc = RR.load_pointcloud_module()
    sys.modules["pointcloud"] = pc                 # provider.py: `from pointcloud import str_to_np`
    sys.path.insert(0, RR.REFERENCE_ROOT)
    sys.modules.pop("provider", None)
    provider = importlib.import_module("provider")   # real module (imports config, which re-imports provider)
    config = importlib.import_module("config")
    rng = np.random.Generator(np.random.PCG64(5))
    shutil.rmtree(BASE, ignore_errors=True)
    for d in ("meta", "pointcloud1", "pointcloud2", "split"):
        os.makedirs(os.path.join(BASE, d))
    n_ex = 6
    sizes = [(120, 80), (300, 17), (64, 64), (0, 33), (5, 250), (200, 1)]     # one empty cloud, ragged, 4 columns
    for i in range(n_ex):
        start = rng.normal(size=3) * 8
        trans = rng.normal(size=3) * 0.5
        a0, ra = float(rng.uniform(-np.pi, np.pi)), float(rng.uniform(-0.7, 0.7))
        meta = dict(translation=pc.np_to_str(trans), rel_angle=ra, start_position=pc.np_to_str(start),
                    end_position=pc.np_to_str(start + trans), start_angle=a0, end_angle=a0 + ra)
        json.dump(meta, open(os.path.join(BASE, "meta", f"{i:08d}.json"), "w"))
        for w, n in enumerate(sizes[i]):
            cloud = (rng.normal(size=(n, 4)) * 2 + np.array([start[0], start[1], start[2], 0.0])).astype(np.float32)
            np.save(os.path.join(BASE, f"pointcloud{w + 1}", f"{i:08d}.npy"), cloud)
    with open(os.path.join(BASE, "split", "val.txt"), "w") as fh:
        fh.write("\n".join(str(i) for i in range(n_ex)) + "\n")
    config.dump_to_namespace(config.configGlobal, {"data": {"basepath": BASE, "num_channels": 3},
                                                   "model": {"num_points": N_POINTS}, "training": {"batch_size": n_ex}})
    idx = provider.getDataFiles(os.path.join(BASE, "split", "val.txt"))
    order = [3, 0, 5, 1, 4, 2]
    np.random.seed(77)
    pcs1, pcs2, t, ra, c1, c2, a1, a2 = provider.load_batch([idx[i] for i in order])
    pcs1j = provider.jitter_point_cloud(pcs1)
    pcs2j = provider.jitter_point_cloud(pcs2)
    np.savez_compressed(os.path.join(HERE, "reference_provider.npz"), order=np.array(order), num_points=N_POINTS, pcs1=pcs1,
                        pcs2=pcs2, translations=t, rel_angles=ra, pc1_centers=c1, pc2_centers=c2, pc1_angles=a1,
                        pc2_angles=a2, pcs1_jittered=pcs1j, pcs2_jittered=pcs2j)
    print("written", pcs1.shape, float(np.abs(pcs1).sum()))


if __name__ == "__main__":
    main()
