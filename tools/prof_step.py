"""Profiling target: a few eager (non-graph) steps of one BASELINE workload, nothing else -- what `ncu` wraps.

    ncu --set full --clock-control none --import-source on -k regex:conv_stack_fwd_kernel -s 4 -c 1 -o gpurun_out/x \
        python tools/prof_step.py --workload c3 --steps 1
Launch order of a training step's heavy kernels: conv_stats2 / conv_stack_fwd (S1 b0, S1 b1, S2 b0, S2 b1, EMB b0, EMB b1),
then the backward in reverse stage order.  AN3D_TWO_STREAMS=0 keeps the launch order deterministic under the profiler."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("AN3D_TWO_STREAMS", "0")
import torch
import __graft_entry__ as ge

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3")
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
ge.build()
from alignnet_b200 import engine, synth
from bench import WORKLOADS

wl = WORKLOADS[a.workload]
eng = engine.Engine(engine.default_arch() if wl.get("arch") == "default" else engine.shipped_arch(), "cuda:0", a.precision, seed=0)
batch = {k: torch.from_numpy(v).cuda() for k, v in synth.make_batch_fast(wl["B"], wl["N"], seed=1236).items()}
if not wl["train"]:
    for i in range(3):
        eng.forward(batch["pcs1"], batch["pcs2"], True, 0.5, None, seed=i)
for i in range(a.steps):
    if wl["train"]:
        eng.train_step(batch, lr=0.005, bn_decay=0.5)
    else:
        eng.forward(batch["pcs1"], batch["pcs2"], False)
torch.cuda.synchronize()
print("done")
