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
a = ap.parse_args()
ge.build()
from alignnet_b200 import engine, synth, evaluation

dev = lambda d: {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in d.items()}
train = [dev(synth.make_batch_fast(a.batch, a.points, seed=1000 + i)) for i in range(a.pool)]
val_host = synth.make_batch_fast(1024, a.points, seed=999)
val = dev(val_host)
out = {}
for prec in ("fp32", "bf16"):
    e = engine.Engine(engine.shipped_arch(), "cuda:0", prec, seed=3)
    losses = []
    for t in range(a.steps):
        losses.append(float(e.train_step(train[t % a.pool], lr=a.lr, bn_decay=0.5, seed=t)[0].cpu()))
    ep = e.forward(val["pcs1"], val["pcs2"], False)
    vloss = float(e.loss(val, ep)[0].cpu())
    pa = e.pred_angles(ep).cpu().numpy()
    pt = ep["pred_translations"].cpu().numpy()
    terr = np.linalg.norm(pt[:, :2] - val_host["translations"][:, :2], axis=1)
    aerr = np.abs((pa - val_host["rel_angles"][:, 0] + np.pi) % (2 * np.pi) - np.pi)
    out[prec] = dict(loss_first=losses[:5], loss_last=float(np.mean(losses[-20:])), val_loss=vloss,
                     t_err_mean=float(terr.mean()), t_err_med=float(np.median(terr)), a_err_mean_deg=float(np.degrees(aerr.mean())),
                     a_err_med_deg=float(np.degrees(np.median(aerr))), t_lt_10cm=float((terr < 0.1).mean()),
                     t_lt_20cm=float((terr < 0.2).mean()), a_lt_5deg=float((aerr < np.radians(5)).mean()),
                     a_lt_10deg=float((aerr < np.radians(10)).mean()), traj=[float(np.mean(losses[i:i + 20])) for i in range(0, a.steps, 20)])
    print(prec, json.dumps(out[prec]))
