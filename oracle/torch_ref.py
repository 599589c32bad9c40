"""torch-CPU restatement (fp32 or fp64, autograd) of tp8 forward + loss + TF-Adam.

Oracle = test infrastructure (see oracle/__init__.py).  Independent of np_forward.py so the
two restatements check each other; this one also supplies gradients and is the timed CPU
baseline in bench.py.  Citations are into /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .arch import Arch, BN_EPS, bn_names, stage_specs, trainable_specs, weight_names


def to_torch(d: Dict[str, np.ndarray], dtype=torch.float64, requires_grad: bool = False) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in d.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        if requires_grad:
            t.requires_grad_(True)
        out[k] = t
    return out


# --------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------

def _bn(z, spec, branch, params, state, new_state, training, bn_decay, axes):
    """utils/tf_util.py:455-492 (moments biased; shadows via s -= (1-d)(s-stat); eps 1e-3)."""
    n = bn_names(spec, branch)
    gamma, beta = params[n["gamma"]], params[n["beta"]]
    if training:
        mean = z.mean(dim=axes)
        var = ((z - mean.detach()) ** 2).mean(dim=axes)        # squared_difference(x, stop_gradient(mean)) [TF-sem]
        d = 0.9 if bn_decay is None else float(bn_decay)
        with torch.no_grad():
            new_state[n["moving_mean"]] = state[n["moving_mean"]] - (1.0 - d) * (state[n["moving_mean"]] - mean)
            new_state[n["moving_var"]] = state[n["moving_var"]] - (1.0 - d) * (state[n["moving_var"]] - var)
    else:
        mean, var = state[n["moving_mean"]], state[n["moving_var"]]
    inv = gamma * torch.rsqrt(var + BN_EPS)
    return z * inv + (beta - mean * inv)


SIM_BF16 = False   # model the rounding points of the engine's bf16 fast mode (tests only)
# With SIM_BF16: layers with a dimension (rows, inputs, outputs) below this stay exact.  0 = the fused kernels' FC path
# (every FC layer on the tensor cores); 8 = the engine's layer-by-layer tensor-core path, which keeps 3-wide layers on
# the CUDA cores in fp32 (csrc/gemm_tc.cuh: use_tensor_cores).
SIM_BF16_MIN_DIM = 0


def _sim_round(x, w):
    return SIM_BF16 and min(x.reshape(-1, x.shape[-1]).shape[0], w.shape[0], w.shape[1]) >= SIM_BF16_MIN_DIM


def _r16(t):
    return t.to(torch.bfloat16).to(t.dtype)


def _conv_stack(p, specs, branch, params, state, new_state, training, bn_decay):
    """models/tp8.py:49-59.  With SIM_BF16 the inputs of every conv after the first (activations and
    weights) are rounded to bf16, which is where the engine's tensor-core path rounds."""
    x = p
    for i, s in enumerate(specs):
        n = weight_names(s)
        w = params[n["weights"]]
        if i > 0 and _sim_round(x, w):
            x, w = _r16(x), _r16(w)
        z = x @ w + params[n["biases"]]
        x = torch.relu(_bn(z, s, branch, params, state, new_state, training, bn_decay, (0, 1)))
    return x.max(dim=1).values


def _mlp(g, specs, branch, params, state, new_state, training, bn_decay, keep, mask):
    """models/tp8.py:75-82.  With SIM_BF16 every FC layer rounds its inputs and weights to bf16 (the
    engine's bf16 mode runs all FC GEMMs on the tensor cores with fp32 accumulation)."""
    x = g
    for s in specs[:-1]:
        n = weight_names(s)
        w = params[n["weights"]]
        if _sim_round(x, w):
            x, w = _r16(x), _r16(w)
        z = x @ w + params[n["biases"]]
        x = torch.relu(_bn(z, s, branch, params, state, new_state, training, bn_decay, (0,)))
    if training and keep is not None and mask is not None:
        x = x / keep * mask
    n = weight_names(specs[-1])
    w = params[n["weights"]]
    if _sim_round(x, w):
        x, w = _r16(x), _r16(w)
    return x @ w + params[n["biases"]]


def tf_get_angles(logits, nb: int):
    """models/tp8.py:294-301 + :202-212.  argmax carries no gradient; the residual does
    (d angle / d residual_logit = pi/nb); tf.mod has unit gradient [TF-sem]."""
    k = torch.argmax(logits[:, :nb], dim=1)
    res = logits[:, nb:] * (math.pi / nb)
    r = res.gather(1, k[:, None])[:, 0]
    a = k.to(logits.dtype) * (2.0 * math.pi / nb) + r
    return torch.remainder(a + math.pi, 2.0 * math.pi) - math.pi


def _rot_z_rows(p, a):
    """models/tp8.py:26-27,125-127: p @ Rz(-a), row vectors."""
    c, s = torch.cos(-a), torch.sin(-a)
    z, o = torch.zeros_like(a), torch.ones_like(a)
    R = torch.stack([torch.stack([c, -s, z], -1), torch.stack([s, c, z], -1), torch.stack([z, z, o], -1)], -2)
    return torch.matmul(p, R)


def _embedding_net(pcs, branch, arch: Arch, params, state, new_state, training, bn_decay, masks):
    """models/tp8.py:101-132."""
    st = stage_specs(arch)
    mu = pcs.mean(dim=1)
    g1 = _conv_stack(pcs - mu[:, None, :], st["s1_conv"], branch, params, state, new_state, training, bn_decay)
    d1 = _mlp(g1, st["s1_fc"], branch, params, state, new_state, training, bn_decay, arch.s1_keep,
              None if masks is None else masks.get(f"s1_b{branch}"))
    c1 = d1 + mu
    g2 = _conv_stack(pcs - c1[:, None, :], st["s2_conv"], branch, params, state, new_state, training, bn_decay)
    o2 = _mlp(g2, st["s2_fc"], branch, params, state, new_state, training, bn_decay, arch.s2_keep,
              None if masks is None else masks.get(f"s2_b{branch}"))
    c2 = o2[:, :3] + c1
    logits = o2[:, 3:]
    ang = tf_get_angles(logits, arch.num_bins)
    q = _rot_z_rows(pcs - c2[:, None, :], ang)
    e = _conv_stack(q, st["emb_conv"], branch, params, state, new_state, training, bn_decay)
    return e, mu, c1, c2, logits


def get_model(pcs1, pcs2, arch: Arch, params, state, is_training: bool, bn_decay: Optional[float] = None, masks=None):
    """models/tp8.py:135-158.  Returns (end_points, new_state)."""
    new_state = dict(state)
    e1, _, s1c1, s2c1, lg1 = _embedding_net(pcs1, 0, arch, params, state, new_state, is_training, bn_decay, masks)
    e2, _, s1c2, s2c2, lg2 = _embedding_net(pcs2, 1, arch, params, state, new_state, is_training, bn_decay, masks)
    st = stage_specs(arch)
    o = _mlp(torch.cat([e1, e2], dim=1), st["head_fc"], 0, params, state, new_state, is_training, bn_decay,
             arch.head_keep, None if masks is None else masks.get("head"))
    end_points = {
        "pred_s1_pc1centers": s1c1, "pred_s1_pc2centers": s1c2,
        "pred_s2_pc1centers": s2c1, "pred_s2_pc2centers": s2c2,
        "pred_pc1angle_logits": lg1, "pred_pc2angle_logits": lg2,
        "pred_translations": o[:, :3] + (s2c2 - s2c1),
        "pred_remaining_angle_logits": o[:, 3:],
    }
    return end_points, new_state


# --------------------------------------------------------------------------------------
# loss (models/tp8.py:173-199, 266-354)
# --------------------------------------------------------------------------------------

def huber_loss(error, delta: float):
    """models/tp8.py:173-178."""
    abs_error = error.abs()
    quadratic = torch.clamp(abs_error, max=delta)
    linear = abs_error - quadratic
    return (0.5 * quadratic ** 2 + delta * linear).mean()


def tf_angle2class(angle, nb: int):
    """models/tp8.py:181-199.  angle: [B,1] or [B,B].  Returns class_id[:,0] ([B], int) and the
    element-wise residual (same shape as `angle`)."""
    twopi = 2.0 * math.pi
    angle = torch.remainder(angle, twopi)
    apc = twopi / nb
    shifted = torch.remainder(angle + apc / 2.0, twopi)
    class_id = (shifted / apc).to(torch.int32)          # tf.to_int32 truncates; shifted >= 0
    residual = shifted - (class_id.to(angle.dtype) * apc + apc / 2.0)
    return class_id[:, 0].long(), residual


def _tf_get_angle_loss(logits, target_angles, nb: int):
    """models/tp8.py:266-281.  NOTE the [B] - [B,1] (or [B,B]) broadcast to [B,B] (Q3)."""
    cls_logits = logits[:, :nb]
    res_norm = logits[:, nb:]
    tgt_cls, tgt_res = tf_angle2class(target_angles, nb)
    cls_loss = torch.nn.functional.cross_entropy(cls_logits, tgt_cls, reduction="mean")
    onehot = torch.nn.functional.one_hot(tgt_cls, nb).to(logits.dtype)
    label = tgt_res / (math.pi / nb)
    pred = (res_norm * onehot).sum(dim=1)               # [B]
    res_loss = huber_loss(pred - label, 1.0)            # [B] - [B,1|B] -> [B,B]
    return torch.stack([cls_loss + 20.0 * res_loss, cls_loss, res_loss])


def tf_get_angle_losses(logits, target_angles, nb: int, accept_inverted_angle: bool):
    """models/tp8.py:284-291.  The tf.cond keeps the LARGER total (Q5)."""
    losses = _tf_get_angle_loss(logits, target_angles, nb)
    if accept_inverted_angle:
        losses_180 = _tf_get_angle_loss(logits, target_angles + math.pi, nb)
        losses = losses if bool(losses[0] > losses_180[0]) else losses_180
    return losses[0], losses[1], losses[2]


def get_loss(translations, rel_angles, pc1_centers, pc2_centers, pc1_angles, pc2_angles, end_points, arch: Arch,
             return_parts: bool = False):
    """models/tp8.py:304-354 _get_loss_separate (pcs1/pcs2/rel_angles are unused by this loss)."""
    nb = arch.num_bins
    B = translations.shape[0]
    inv = arch.accept_inverted_angle
    l_s1_1 = huber_loss(end_points["pred_s1_pc1centers"] - pc1_centers, 1.0)
    l_s1_2 = huber_loss(end_points["pred_s1_pc2centers"] - pc2_centers, 1.0)
    stage1_t = (l_s1_1 + l_s1_2) / 2.0
    l_s2_1 = huber_loss(end_points["pred_s2_pc1centers"] - pc1_centers, 1.0)
    l_s2_2 = huber_loss(end_points["pred_s2_pc2centers"] - pc2_centers, 1.0)
    a1 = tf_get_angle_losses(end_points["pred_pc1angle_logits"], pc1_angles, nb, inv)
    a2 = tf_get_angle_losses(end_points["pred_pc2angle_logits"], pc2_angles, nb, inv)
    stage2_t = (l_s2_1 + l_s2_2) / 2.0
    stage2_a = (a1[0] + a2[0]) / 2.0
    stage3_t = huber_loss(end_points["pred_translations"] - translations, 2.0)
    pc1_pred = tf_get_angles(end_points["pred_pc1angle_logits"], nb)
    pc2_pred = tf_get_angles(end_points["pred_pc2angle_logits"], nb)
    remaining = (pc2_angles - pc1_angles) - (pc2_pred - pc1_pred)          # [B,1] - [B] -> [B,B] (Q4)
    a3 = tf_get_angle_losses(end_points["pred_remaining_angle_logits"], remaining, nb, inv)
    esf, af = arch.early_stage_factor, arch.angle_factor
    loss_t = esf * (stage1_t + stage2_t) + stage3_t
    loss_a = esf * stage2_a + a3[0]
    loss = loss_t + af * loss_a
    per_transform = loss / B
    if return_parts:
        parts = dict(translation=loss_t, angle=loss_a, s1_pc1=l_s1_1, s1_pc2=l_s1_2, s2_pc1=l_s2_1, s2_pc2=l_s2_2,
                     s3_t=stage3_t, s2_pc1_angle=a1[0], s2_pc2_angle=a2[0], s3_angle=a3[0],
                     s3_angle_cls=a3[1], s3_angle_res=a3[2])
        return per_transform, parts
    return per_transform


def get_loss_p2p(pcs1, pc1_centers, end_points):
    """models/tp8.py:374-398 _get_loss_p2p as the reference actually computes it (quirk Q6, a21).
    `tf_translate_pcs` (tp8.py:357-358) RETURNS the tiled translation instead of adding it, so every step of
    `tf_transform_pcs` (:361-371) overwrites the cloud and the result is tile(rotation_centers); the predicted
    translation and angle drop out.  `tf.norm(..., axis=1)` (:386) reduces over the POINT axis, hence
        loss = mean_{b,d} N * (pred_s2_pc1centers - pc1_centers)^2 ;   per_transform_loss = loss / B,
    and the `accept_inverted_angle` variant (:388-393) is identical to it.  No shipped config selects this loss
    (`default.json` training.loss.loss = 'separate'); the engine rejects it, the oracle restates it for the record."""
    B, N = pcs1.shape[0], pcs1.shape[1]
    d = end_points["pred_s2_pc1centers"] - pc1_centers
    return (N * d * d).mean() / B


# --------------------------------------------------------------------------------------
# optimiser + schedules (train.py:133-174, 211-217) [TF-sem for Adam]
# --------------------------------------------------------------------------------------

def learning_rate(step: int, batch_size: int, base_lr: float, decay_step: int, rate: float) -> float:
    """train.py:133-156: max(lr0 * rate^floor(step*B/decay_step), 1e-5) (staircase)."""
    return max(base_lr * rate ** math.floor(step * batch_size / decay_step), 1e-5)


def bn_decay(step: int, batch_size: int, init: float, decay_step: int, rate: float, clip: float) -> float:
    """train.py:159-174: min(clip, 1 - init * rate^floor(step*B/decay_step))."""
    return min(clip, 1.0 - init * rate ** math.floor(step * batch_size / decay_step))


def adam_step(params, grads, m, v, lr: float, t: int, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer [TF-sem]: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA;
    theta -= lr_t * m / (sqrt(v) + eps).  t counts from 1.  In place on numpy / torch arrays."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    for k in params:
        g = grads[k]
        m[k] = beta1 * m[k] + (1.0 - beta1) * g
        v[k] = beta2 * v[k] + (1.0 - beta2) * g * g
        params[k] = params[k] - lr_t * m[k] / (v[k] ** 0.5 + eps)
    return params, m, v


def momentum_step(params, grads, accum, lr: float, momentum: float):
    """tf.train.MomentumOptimizer(lr, momentum) (train.py:211-212) [TF-sem: ApplyMomentum, use_nesterov = False]:
    accum = momentum * accum + g ; theta -= lr * accum.  In place on numpy / torch arrays."""
    for k in params:
        accum[k] = momentum * accum[k] + grads[k]
        params[k] = params[k] - lr * accum[k]
    return params, accum


# --------------------------------------------------------------------------------------
# convenience: one full train-mode evaluation with gradients
# --------------------------------------------------------------------------------------

def loss_and_grads(batch: Dict[str, np.ndarray], arch: Arch, params_np, state_np, bn_decay_value=None, masks_np=None,
                   dtype=torch.float64, training: bool = True):
    """Returns (loss float, end_points numpy, grads numpy dict keyed like params, new_state numpy)."""
    params = to_torch(params_np, dtype, requires_grad=True)
    state = to_torch(state_np, dtype)
    masks = None if masks_np is None else to_torch(masks_np, dtype)
    b = to_torch(batch, dtype)
    ep, new_state = get_model(b["pcs1"], b["pcs2"], arch, params, state, training, bn_decay_value, masks)
    loss = get_loss(b["translations"], b["rel_angles"], b["pc1_centers"], b["pc2_centers"], b["pc1_angles"],
                    b["pc2_angles"], ep, arch)
    names = [n for n, _ in trainable_specs(arch)]
    gl = torch.autograd.grad(loss, [params[n] for n in names], allow_unused=True)
    grads = {n: (np.zeros_like(params_np[n], dtype=np.float64) if g is None else g.detach().numpy().astype(np.float64))
             for n, g in zip(names, gl)}
    ep_np = {k: v.detach().numpy() for k, v in ep.items()}
    st_np = {k: v.detach().numpy() for k, v in new_state.items()}
    return float(loss.detach()), ep_np, grads, st_np
