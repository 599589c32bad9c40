"""Trains the fp32 parity mode and the bf16 fast mode from the same initial parameters on the same seeded synthetic
stream and prints loss trajectories + evaluation metrics of both: the raw material of tests/test_gpu_convergence.py."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--points", type=int, default=200)
ap.add_argument("--lr", type=float, default=0.001)
ap.add_argument("--pool", type=int, default=16, help="distinct training batches cycled through")
ap.add_argument("--every", type=int, default=200, help="evaluate every this many steps")
ap.add_argument("--runs", default="fp32:0,bf16:0,fp32:1,bf16:1", help="precision:dropout-seed-offset list")
a = ap.parse_args()
ge.build()
from alignnet_b200 import engine, synth, evaluation

dev = lambda d: {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in d.items()}
train = [dev(synth.make_batch_fast(a.batch, a.points, seed=1000 + i)) for i in range(a.pool)]
val_host = synth.make_batch_fast(1024, a.points, seed=999)
val = dev(val_host)
out = {}


def evaluate(e):
    ep = e.forward(val["pcs1"], val["pcs2"], False)
    vloss = float(e.loss(val, ep)[0].cpu())
    ev = evaluation.evaluate(ep["pred_translations"], e.pred_angles(ep), val_host["translations"], val_host["rel_angles"],
                             ep["pred_s2_pc1centers"], val_host["pc1_centers"], accept_inverted_angle=True)
    return dict(val_loss=round(vloss, 5), t_mean=round(ev["mean_dist_translation"], 4), a_mean_deg=round(ev["mean_dist_angle"], 2),
                t_levels=[round(x, 3) for x in ev["corr_levels_translation"]], a_levels=[round(x, 3) for x in ev["corr_levels_angles"]])


for run in a.runs.split(","):
    prec, off = run.split(":")
    e = engine.Engine(engine.shipped_arch(), "cuda:0", prec, seed=3)
    losses, hist = [], []
    for t in range(a.steps):
        losses.append(float(e.train_step(train[t % a.pool], lr=a.lr, bn_decay=0.5, seed=t + 100000 * int(off))[0].cpu()))
        if (t + 1) % a.every == 0:
            hist.append(dict(step=t + 1, train_loss=round(float(np.mean(losses[-20:])), 5), **evaluate(e)))
            print(run, json.dumps(hist[-1]), flush=True)
    out[run] = hist
