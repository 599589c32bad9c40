"""NumPy fp32 restatement of the tp8 forward pass, op for op in the reference's order.

Oracle = test infrastructure (see oracle/__init__.py).  Every function cites the reference
lines (into /root/reference) it follows.  [TF-sem] marks TensorFlow-1.8 op semantics that are
not in the reference tree.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from .arch import Arch, BN_EPS, bn_names, stage_specs, weight_names

F32 = np.float32


def _bn(z: np.ndarray, spec, branch: int, params, state, new_state, training: bool, bn_decay, axes):
    """utils/tf_util.py:455-492 batch_norm_template.
    training: tf.nn.moments (mean, then mean of squared difference -> biased variance) [TF-sem];
              shadow <- shadow - (1-decay)*(shadow - stat), zero-initialised, no debias [TF-sem];
    eval:     use the shadows (tf_util.py:490).
    normalise: inv = gamma*rsqrt(var+eps); y = z*inv + (beta - mean*inv) [TF-sem], eps=1e-3 (:491)."""
    n = bn_names(spec, branch)
    gamma, beta = params[n["gamma"]], params[n["beta"]]
    if training:
        mean = z.mean(axis=axes, dtype=F32)
        var = np.square(z - mean, dtype=F32).mean(axis=axes, dtype=F32)
        d = F32(0.9 if bn_decay is None else bn_decay)
        one_minus = F32(1.0) - d
        new_state[n["moving_mean"]] = (state[n["moving_mean"]] - one_minus * (state[n["moving_mean"]] - mean)).astype(F32)
        new_state[n["moving_var"]] = (state[n["moving_var"]] - one_minus * (state[n["moving_var"]] - var)).astype(F32)
    else:
        mean, var = state[n["moving_mean"]], state[n["moving_var"]]
    inv = (gamma / np.sqrt(var + F32(BN_EPS), dtype=F32)).astype(F32)
    return (z * inv + (beta - mean * inv)).astype(F32)


def _conv_stack(p: np.ndarray, specs, branch, params, state, new_state, training, bn_decay, aux, tag):
    """models/tp8.py:49-59 _get_pointnet: conv(+bias)+BN+ReLU per layer (utils/tf_util.py:145-169),
    then max over the N axis (tf_util.py:350-373).  p: [B,N,3] -> [B,C_last]."""
    x = p
    for i, s in enumerate(specs):
        n = weight_names(s)
        z = (x @ params[n["weights"]] + params[n["biases"]]).astype(F32)      # conv2d + bias_add
        y = _bn(z, s, branch, params, state, new_state, training, bn_decay, axes=(0, 1))
        x = np.maximum(y, F32(0))                                              # tf.nn.relu
        if aux is not None:
            aux[f"{tag}/z{i + 1}"] = z
    return x.max(axis=1)


def _mlp(g: np.ndarray, specs, branch, params, state, new_state, training, bn_decay, keep, mask):
    """models/tp8.py:75-82 get_mlp: fc(+bias)+BN+ReLU for all but the last layer
    (utils/tf_util.py:330-347), dropout after the last hidden layer in training only
    (tp8.py:80-81; tf.nn.dropout = x/keep*mask [TF-sem]), linear output layer (tp8.py:82)."""
    x = g
    for s in specs[:-1]:
        n = weight_names(s)
        z = (x @ params[n["weights"]] + params[n["biases"]]).astype(F32)
        y = _bn(z, s, branch, params, state, new_state, training, bn_decay, axes=(0,))
        x = np.maximum(y, F32(0))
    if training and keep is not None and mask is not None:
        x = (x / F32(keep) * mask.astype(F32)).astype(F32)
    s = specs[-1]
    n = weight_names(s)
    return (x @ params[n["weights"]] + params[n["biases"]]).astype(F32)


def get_angles(logits: np.ndarray, nb: int) -> np.ndarray:
    """models/tp8.py:294-301 tf_get_angles + :202-212 tf_class2angle: first-max argmax,
    residual scaled by pi/nb, wrapped with floor-mod to [-pi, pi)."""
    k = np.argmax(logits[:, :nb], axis=1)
    res = logits[:, nb:] * (F32(np.pi) / F32(nb))
    r = res[np.arange(logits.shape[0]), k]
    apc = F32(2.0) * F32(np.pi) / F32(nb)
    a = k.astype(F32) * apc + r
    two_pi = F32(2.0) * F32(np.pi)
    return (np.mod(a + F32(np.pi), two_pi) - F32(np.pi)).astype(F32)


def rot_z_rows(p: np.ndarray, a: np.ndarray) -> np.ndarray:
    """models/tp8.py:125-127: (p) @ Rz(-a) with Rz(t)=[[c,-s,0],[s,c,0],[0,0,1]] (tp8.py:26-27),
    row-vector convention  ==  x' = x cos a - y sin a ; y' = x sin a + y cos a."""
    c, s = np.cos(-a, dtype=F32), np.sin(-a, dtype=F32)
    R = np.zeros((a.shape[0], 3, 3), F32)
    R[:, 0, 0], R[:, 0, 1], R[:, 1, 0], R[:, 1, 1], R[:, 2, 2] = c, -s, s, c, 1
    return np.matmul(p, R).astype(F32)


def embedding_net(pcs: np.ndarray, branch: int, arch: Arch, params, state, new_state, training, bn_decay,
                  masks: Optional[Dict[str, np.ndarray]], aux):
    """models/tp8.py:101-132 get_embedding_net for one siamese branch."""
    st = stage_specs(arch)
    tag = f"b{branch}"
    mu = pcs.mean(axis=1, dtype=F32)                                           # :104
    p0 = (pcs - mu[:, None, :]).astype(F32)                                    # :106
    g1 = _conv_stack(p0, st["s1_conv"], branch, params, state, new_state, training, bn_decay, aux, tag + "/s1")
    d1 = _mlp(g1, st["s1_fc"], branch, params, state, new_state, training, bn_decay, arch.s1_keep,
              None if masks is None else masks.get(f"s1_b{branch}"))
    c1 = (d1 + mu).astype(F32)                                                 # :109
    p1 = (pcs - c1[:, None, :]).astype(F32)                                    # :113
    g2 = _conv_stack(p1, st["s2_conv"], branch, params, state, new_state, training, bn_decay, aux, tag + "/s2")
    o2 = _mlp(g2, st["s2_fc"], branch, params, state, new_state, training, bn_decay, arch.s2_keep,
              None if masks is None else masks.get(f"s2_b{branch}"))
    c2 = (o2[:, :3] + c1).astype(F32)                                          # :117
    logits = o2[:, 3:]                                                         # :118
    p2 = (pcs - c2[:, None, :]).astype(F32)                                    # :122
    ang = get_angles(logits, arch.num_bins)                                    # :123
    q = rot_z_rows(p2, ang)                                                    # :125-127
    e = _conv_stack(q, st["emb_conv"], branch, params, state, new_state, training, bn_decay, aux, tag + "/emb")
    if aux is not None:
        aux[tag + "/g1"], aux[tag + "/g2"], aux[tag + "/e"], aux[tag + "/angle"] = g1, g2, e, ang
        aux[tag + "/mu"] = mu
    return e, mu, c1, c2, logits


def get_model(pcs1: np.ndarray, pcs2: np.ndarray, arch: Arch, params: Dict[str, np.ndarray],
              state: Dict[str, np.ndarray], is_training: bool, bn_decay: Optional[float] = None,
              masks: Optional[Dict[str, np.ndarray]] = None, return_aux: bool = False):
    """models/tp8.py:135-158 get_model.  Returns (end_points, new_state[, aux]).
    masks: dropout keep-masks {s1_b0,s1_b1,s2_b0,s2_b1,head} of shape [B, last hidden width]."""
    pcs1 = np.asarray(pcs1, F32)
    pcs2 = np.asarray(pcs2, F32)
    new_state = dict(state)
    aux = {} if return_aux else None
    e1, _, s1c1, s2c1, lg1 = embedding_net(pcs1, 0, arch, params, state, new_state, is_training, bn_decay, masks, aux)
    e2, _, s1c2, s2c2, lg2 = embedding_net(pcs2, 1, arch, params, state, new_state, is_training, bn_decay, masks, aux)
    feat = np.concatenate([e1, e2], axis=1)                                    # :144,153
    st = stage_specs(arch)
    o = _mlp(feat, st["head_fc"], 0, params, state, new_state, is_training, bn_decay, arch.head_keep,
             None if masks is None else masks.get("head"))
    end_points = {
        "pred_s1_pc1centers": s1c1, "pred_s1_pc2centers": s1c2,
        "pred_s2_pc1centers": s2c1, "pred_s2_pc2centers": s2c2,
        "pred_pc1angle_logits": lg1, "pred_pc2angle_logits": lg2,
        "pred_translations": (o[:, :3] + (s2c2 - s2c1)).astype(F32),           # :155
        "pred_remaining_angle_logits": o[:, 3:],                               # :156
    }
    if return_aux:
        return end_points, new_state, aux
    return end_points, new_state


def class2angle(pred_cls: int, residual: float, nb: int, to_label_format: bool = True) -> float:
    """models/tp8.py:229-238 (host decode; residual is NOT scaled by pi/nb -- quirk Q1)."""
    angle_per_class = 2 * np.pi / float(nb)
    angle = pred_cls * angle_per_class + residual
    if to_label_format and angle > np.pi:
        angle = angle - 2 * np.pi
    return angle


def classLogits2angle(logits: np.ndarray, nb: int) -> np.ndarray:
    """models/tp8.py:241-244."""
    class_logits, residuals = logits[:, :nb], logits[:, nb:]
    classes = np.argmax(class_logits, axis=1)
    return np.array([class2angle(c, r[c], nb) for c, r in zip(classes, residuals)])


def pred_angles(end_points: Dict[str, np.ndarray], nb: int) -> np.ndarray:
    """train.py:453-456: dec(pc2) - dec(pc1) + dec(remaining), no wrap (Q10)."""
    a1 = classLogits2angle(end_points["pred_pc1angle_logits"], nb)
    a2 = classLogits2angle(end_points["pred_pc2angle_logits"], nb)
    ar = classLogits2angle(end_points["pred_remaining_angle_logits"], nb)
    return a2 - a1 + ar
