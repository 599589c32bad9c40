"""Generates tests/golden/reference_eval.{npz,json} by calling the reference's OWN evaluation.evaluate()
(/root/reference/evaluation.py, unmodified) on seeded synthetic predictions.  The function reads one meta JSON per
sample (only to pick the val/test split: for a 'Synth' base path the split is idx >= 1000), so a throw-away
directory with empty meta files is created.  Run from the repo root: python tests/golden/make_reference_eval_golden.py"""
import importlib
import json
import os
import sys
import tempfile
from argparse import Namespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_run as RR  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.modules["pointcloud"] = RR.load_pointcloud_module()          # evaluation.py: `from pointcloud import ...`
    sys.path.insert(0, RR.REFERENCE_ROOT)
    ev = importlib.import_module("evaluation")
    rng = np.random.Generator(np.random.PCG64(2024))
    n = 1500
    gt_c1 = rng.normal(size=(n, 3)) * np.array([9.0, 9.0, 0.5])
    gt_t = rng.normal(size=(n, 3)) * 0.5
    gt_a = rng.uniform(-np.pi, np.pi, size=(n, 1)) / 2
    scale = rng.choice([0.005, 0.05, 0.3], size=(n, 1))               # populate every threshold level
    pred_t = (gt_t + rng.normal(size=(n, 3)) * scale).astype(np.float32).astype(np.float64)
    pred_a = gt_a + rng.normal(size=(n, 1)) * np.deg2rad(rng.choice([0.5, 3.0, 8.0, 40.0], size=(n, 1)))
    flip = rng.uniform(size=(n, 1)) < 0.2
    pred_a = np.where(flip, pred_a + np.pi, pred_a).astype(np.float32).astype(np.float64)
    pred_c = (gt_c1 + rng.normal(size=(n, 3)) * 0.3).astype(np.float32).astype(np.float64)
    pred_t[7] = 1e6                                                   # > 10000 m: skipped by the reference (:168)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        base = os.path.join(tmp, "SynthCars")
        os.makedirs(os.path.join(base, "meta"))
        for i in range(n):
            with open(os.path.join(base, "meta", f"{i:08d}.json"), "w") as fh:
                fh.write("{}")
        cfg = Namespace(data=Namespace(basepath=base))
        for inv in (False, True):
            d = ev.evaluate(cfg, list(range(n)), pred_t, pred_a, gt_t, gt_a, pred_c, gt_c1, eval_dir=None,
                            accept_inverted_angle=inv, mean_time=0.0125)
            out["inverted" if inv else "plain"] = ev.ns_to_dict(d)
    np.savez_compressed(os.path.join(HERE, "reference_eval.npz"), pred_t=pred_t, pred_a=pred_a, pred_c=pred_c, gt_t=gt_t,
                        gt_a=gt_a, gt_c1=gt_c1, is_test=(np.arange(n) >= 1000))
    with open(os.path.join(HERE, "reference_eval.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("written; plain corr_levels", out["plain"]["corr_levels"], "inverted", out["inverted"]["corr_levels"])


if __name__ == "__main__":
    main()
