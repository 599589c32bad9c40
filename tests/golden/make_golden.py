"""Generates the committed golden vectors from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference ships no golden data for the network path
(SURVEY section 4), so these pin the oracle's own outputs on seeded inputs: kernel tests then do
not depend on oracle runtime, and an accidental change of the oracle is caught on CPU.

Parameters are NOT stored (8.7 MB); they are regenerated from the seed by oracle.arch.init_params
+ randomize_for_test, which is deterministic (numpy PCG64).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import arch as A, np_forward as NF, torch_ref as TR  # noqa: E402
from alignnet_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case(name, arch, B, N, seed, grads=True):
    params = A.init_params(arch, seed)
    state = A.init_state(arch)
    params, state = A.randomize_for_test(arch, params, state, seed + 1)
    batch = synth.make_batch(B, N, seed=seed + 2, persons_prob=0.2)
    rng = np.random.Generator(np.random.PCG64(seed + 3))
    widths = {"s1_b0": arch.s1_fc[-1], "s1_b1": arch.s1_fc[-1], "s2_b0": arch.s2_fc[-1], "s2_b1": arch.s2_fc[-1],
              "head": arch.head_fc[-1]}
    masks = {k: (rng.uniform(size=(B, w)) < 0.7).astype(np.float32) for k, w in widths.items()}
    out = {"B": B, "N": N, "seed": seed}
    for k, v in batch.items():
        out["in/" + k] = v
    for k, v in masks.items():
        out["mask/" + k] = v
    ep_eval, _ = NF.get_model(batch["pcs1"], batch["pcs2"], arch, params, state, False)
    for k, v in ep_eval.items():
        out["eval/" + k] = v.astype(np.float32)
    out["eval/pred_angles"] = NF.pred_angles(ep_eval, arch.num_bins)
    ep_tr, new_state = NF.get_model(batch["pcs1"], batch["pcs2"], arch, params, state, True, 0.5, masks)
    for k, v in ep_tr.items():
        out["train/" + k] = v.astype(np.float32)
    if grads:
        loss, ep64, g, st64 = TR.loss_and_grads(batch, arch, params, state, 0.5, masks)
        out["train/loss"] = np.float64(loss)
        for k, v in ep64.items():
            out["train64/" + k] = v
        for k, v in g.items():
            out["gradnorm/" + k] = np.float64(np.sqrt((v ** 2).sum()))
            if v.size <= 2048:
                out["grad/" + k] = v
        for k, v in st64.items():
            if v.size <= 256:
                out["state/" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written;", "loss" if grads else "", out.get("train/loss"))


if __name__ == "__main__":
    case("tiny_B4_N16", A.tiny_arch(), 4, 16, seed=11)
    case("shipped_B4_N16", A.Arch(), 4, 16, seed=21)
    case("shipped_B32_N200", A.Arch(), 32, 200, seed=31)
    case("default_B32_N64", A.default_arch(), 32, 64, seed=41)      # configs/default.json: five-layer conv stacks
