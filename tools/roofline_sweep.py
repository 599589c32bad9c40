"""Roofline sweep of the tp8 hot path on one B200: pairs/s and fraction of the measured
sustained bf16 tensor peak (whole-step convention: 2 MACs forward, 6 MACs forward+backward) over cloud size N and
per-GPU batch B.  CUDA-graph replay, inputs resident, L2 flushed between timed iterations.
Output is committed as profiles/r1b_roofline_sweep.txt."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
from alignnet_b200 import engine, synth  # noqa: E402

pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, n):
    tot = 0.0
    for _ in range(n):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


print(f"peak = {pk} TFLOP/s (measured sustained bf16)")
print("mode   N     B    ms/step   pairs/s   TFLOP/s(alg)  frac")
for N in (128, 200, 256, 512, 1024):
    for B in (256, 1024, 2048, 4096):
        if B * N > 4096 * 1024 // 2 and N == 1024 and B > 2048:
            continue
        host = synth.make_batch_fast(B, N, seed=7)
        batch = {k: torch.from_numpy(v).to(dev) for k, v in host.items()}
        macs = 509056.0 * N + 2571008.0
        for mode in ("eval", "train"):
            eng = engine.Engine(engine.shipped_arch(), "cuda:0", "bf16", seed=0)
            for i in range(4):
                eng.forward(batch["pcs1"], batch["pcs2"], True, 0.5, None, seed=i)
            if mode == "train":
                fn = lambda: eng.train_step_graph(batch, lr=0.005, bn_decay=0.5)   # noqa: E731
            else:
                fn = lambda: eng.forward_graph(batch["pcs1"], batch["pcs2"])        # noqa: E731
            for _ in range(3):
                fn()
            ms = timed(fn, 5)
            flops = (6.0 if mode == "train" else 2.0) * macs * B
            tf = flops / (ms * 1e-3) / 1e12
            print(f"{mode:5s} {N:5d} {B:5d} {ms:9.3f} {B / (ms * 1e-3):10.0f} {tf:10.1f}     {tf / pk:5.3f}", flush=True)
            del eng
        del batch
        torch.cuda.empty_cache()
