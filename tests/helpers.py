"""Shared helpers for the test-suite (oracle <-> engine plumbing)."""
import os

import numpy as np

from oracle import arch as A

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BATCH_KEYS = ("pcs1", "pcs2", "translations", "rel_angles", "pc1_centers", "pc2_centers", "pc1_angles", "pc2_angles")
MASK_KEYS = ("s1_b0", "s1_b1", "s2_b0", "s2_b1", "head")
OUTPUT_KEYS = ("pred_s1_pc1centers", "pred_s1_pc2centers", "pred_s2_pc1centers", "pred_s2_pc2centers",
               "pred_pc1angle_logits", "pred_pc2angle_logits", "pred_translations", "pred_remaining_angle_logits")


def golden_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    arch = A.tiny_arch() if name.startswith("tiny") else (A.default_arch() if name.startswith("default") else A.Arch())
    seed = int(g["seed"])
    params = A.init_params(arch, seed)
    state = A.init_state(arch)
    params, state = A.randomize_for_test(arch, params, state, seed + 1)
    batch = {k: g["in/" + k] for k in BATCH_KEYS}
    masks = {k: g["mask/" + k] for k in MASK_KEYS}
    return g, arch, params, state, batch, masks


def engine_arch(arch: A.Arch):
    """oracle Arch -> C-ABI an3d_arch."""
    from alignnet_b200 import engine
    return engine.make_arch(arch.num_bins, arch.s1_conv, arch.s1_fc, arch.s1_keep, arch.s2_conv, arch.s2_fc,
                            arch.s2_keep, arch.emb_conv, arch.head_fc, arch.head_keep, arch.angle_factor,
                            arch.early_stage_factor, arch.accept_inverted_angle)


def top2_margin(logits, nb):
    s = np.sort(logits[:, :nb], axis=1)
    return s[:, -1] - s[:, -2]
