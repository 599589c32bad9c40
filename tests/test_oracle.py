"""CPU tests of the oracle itself: the two restatements agree, the committed golden vectors are
reproduced, the reference's only known-answer vectors (euler2mat doctests) hold, and invariances
that follow from the reference code hold."""
import math

import numpy as np
import pytest
import torch

from oracle import arch as A, np_forward as NF, rigid as RG, torch_ref as TR
from alignnet_b200 import synth
from helpers import BATCH_KEYS, OUTPUT_KEYS, golden_case


def test_param_count_matches_survey():
    # 2,165,073 trainable parameters for the shipped architecture
    assert A.num_trainable(A.Arch()) == 2165073
    assert sum(int(np.prod(s)) for _, s in A.state_specs(A.Arch())) == 2 * 8576


def test_rz_doctest_vectors():
    # utils/eulerangles.py:152-154: euler2mat(z=pi/2) . I == [[0,-1,0],[1,0,0],[0,0,1]]
    assert np.allclose(RG.rot_z(math.pi / 2) @ np.eye(3), [[0, -1, 0], [1, 0, 0], [0, 0, 1]])
    # composition / column-vector convention (eulerangles.py:129-149): CCW rotation of e_x gives +e_y
    assert np.allclose(RG.rot_z(math.pi / 2) @ np.array([1.0, 0, 0]), [0, 1, 0])
    # scipy's rotvec convention used by get_mat_angle (pointcloud.py:288) is the same matrix
    from scipy.spatial.transform import Rotation
    for th in (-2.5, -0.3, 0.0, 0.7, 3.0):
        assert np.allclose(Rotation.from_rotvec(np.array([0, 0, 1.0]) * th).as_matrix(), RG.rot_z(th))
    # tp8.py:125-127 row-vector form p @ Rz(-a) equals column-vector Rz(a) p
    p = np.random.default_rng(0).normal(size=(5, 7, 3)).astype(np.float32)
    a = np.array([0.3, -1.2, 2.0, 0.0, 3.1], np.float32)
    q = NF.rot_z_rows(p, a)
    for b in range(5):
        assert np.allclose(q[b], (RG.rot_z(float(a[b])) @ p[b].T).T, atol=1e-5)


def test_rigid_apply_matches_composition():
    rng = np.random.default_rng(1)
    pts = rng.normal(size=(50, 3))
    t, th, c = rng.normal(size=3), 0.8, rng.normal(size=3)
    out = RG.rigid_apply(pts, t, th, c)
    ref = (RG.rot_z(th) @ (pts - c).T).T + c + t
    assert np.allclose(out, ref)
    # list composition (pointcloud.py:293-295) == product of matrices
    hom = np.concatenate([pts, np.ones((50, 1))], 1)
    m1, m2 = RG.get_mat_angle(t, th, c), RG.get_mat_angle(-t, -0.2, c)
    assert np.allclose(RG.transform_points(hom.copy(), [m1, m2]), hom @ m1.T @ m2.T)
    # new centre of rotation keeps the same rigid motion
    new_c = rng.normal(size=3)
    t2 = RG.translate_transform_to_new_center_of_rotation([t], [[th]], [c], [new_c])[0]
    assert np.allclose(RG.rigid_apply(pts, t2, th, new_c), out)


@pytest.mark.parametrize("name", ["tiny_B4_N16", "shipped_B4_N16", "shipped_B32_N200", "default_B32_N64"])
def test_golden_reproduced(name):
    g, arch, params, state, batch, masks = golden_case(name)
    ep, _ = NF.get_model(batch["pcs1"], batch["pcs2"], arch, params, state, False)
    for k in OUTPUT_KEYS:
        np.testing.assert_allclose(ep[k], g["eval/" + k], atol=2e-5, rtol=0)
    ep, _ = NF.get_model(batch["pcs1"], batch["pcs2"], arch, params, state, True, 0.5, masks)
    for k in OUTPUT_KEYS:
        np.testing.assert_allclose(ep[k], g["train/" + k], atol=5e-5, rtol=0)


@pytest.mark.parametrize("name", ["tiny_B4_N16", "shipped_B4_N16"])
def test_numpy_vs_torch_fp64_and_loss_golden(name):
    g, arch, params, state, batch, masks = golden_case(name)
    loss, ep64, grads, st64 = TR.loss_and_grads(batch, arch, params, state, 0.5, masks)
    assert abs(loss - float(g["train/loss"])) < 1e-9
    for k in OUTPUT_KEYS:
        np.testing.assert_allclose(ep64[k], g["train/" + k], atol=1e-4, rtol=0)
    for k, v in grads.items():
        assert abs(np.sqrt((v ** 2).sum()) - float(g["gradnorm/" + k])) <= 1e-9 + 1e-7 * float(g["gradnorm/" + k])


def test_invariances():
    arch = A.tiny_arch()
    params, state = A.randomize_for_test(arch, A.init_params(arch, 5), A.init_state(arch), 6)
    b = synth.make_batch(3, 20, seed=7)
    ep, _ = NF.get_model(b["pcs1"], b["pcs2"], arch, params, state, False)
    # point-permutation invariance (max-pool, tp8.py:58)
    perm = np.random.default_rng(0).permutation(20)
    ep_p, _ = NF.get_model(b["pcs1"][:, perm], b["pcs2"][:, perm], arch, params, state, False)
    for k in OUTPUT_KEYS:
        np.testing.assert_allclose(ep[k], ep_p[k], atol=2e-5)
    # translation equivariance of the centres (mean-centring, tp8.py:104-109), invariance of the logits
    shift = np.array([1.5, -2.0, 0.25], np.float32)
    ep_s, _ = NF.get_model(b["pcs1"] + shift, b["pcs2"] + shift, arch, params, state, False)
    for k in ("pred_s1_pc1centers", "pred_s2_pc1centers", "pred_s1_pc2centers", "pred_s2_pc2centers"):
        np.testing.assert_allclose(ep_s[k], ep[k] + shift, atol=5e-5)
    for k in ("pred_pc1angle_logits", "pred_remaining_angle_logits", "pred_translations"):
        np.testing.assert_allclose(ep_s[k], ep[k], atol=5e-5)


def test_gradients_against_finite_differences():
    arch = A.tiny_arch()
    params, state = A.randomize_for_test(arch, A.init_params(arch, 2), A.init_state(arch), 3)
    b = synth.make_batch(4, 10, seed=4)
    loss0, _, grads, _ = TR.loss_and_grads(b, arch, params, state, 0.5, None)
    rng = np.random.default_rng(0)
    checked = 0
    for name in ("fc3/weights", "siamese/transformer2/mlp/fc3/biases", "siamese/embedding/conv2/weights",
                 "siamese_1/transformer1/embedding/conv1/bn/gamma", "siamese/transformer1/mlp/fc1/weights"):
        idx = tuple(rng.integers(0, s) for s in params[name].shape)
        eps = 1e-6
        vals = []
        for sgn in (+1, -1):
            p2 = {k: v.astype(np.float64).copy() for k, v in params.items()}
            p2[name][idx] += sgn * eps
            l, _, _, _ = TR.loss_and_grads(b, arch, p2, state, 0.5, None)
            vals.append(l)
        fd = (vals[0] - vals[1]) / (2 * eps)
        assert abs(fd - grads[name][idx]) < 1e-5 + 1e-3 * abs(fd), (name, fd, grads[name][idx])
        checked += 1
    assert checked == 5


def test_loss_quirks():
    """Q3/Q4/Q5: [B,B] broadcast and keep-the-larger selection are present in the restatement."""
    nb = 6
    B = 3
    rng = np.random.default_rng(0)
    logits = torch.tensor(rng.normal(size=(B, 2 * nb)))
    target = torch.tensor(rng.uniform(-3, 3, size=(B, 1)))
    tot, cls, res = TR._tf_get_angle_loss(logits, target, nb)
    tgt_cls, tgt_res = TR.tf_angle2class(target, nb)
    pred = logits[:, nb:][torch.arange(B), tgt_cls]
    lab = (tgt_res / (math.pi / nb))
    manual = torch.stack([TR.huber_loss(pred[j] - lab[i, 0], 1.0) for i in range(B) for j in range(B)]).mean()
    assert abs(float(res) - float(manual)) < 1e-12
    l0 = TR._tf_get_angle_loss(logits, target, nb)[0]
    l1 = TR._tf_get_angle_loss(logits, target + math.pi, nb)[0]
    sel = TR.tf_get_angle_losses(logits, target, nb, True)[0]
    assert float(sel) == max(float(l0), float(l1))


def test_adam_and_schedules():
    # train.py:133-174 with per='epoch': decay_step = step * B * batches_per_epoch
    assert TR.learning_rate(0, 128, 0.005, 30 * 128 * 10, 0.5) == 0.005
    assert TR.learning_rate(300, 128, 0.005, 30 * 128 * 10, 0.5) == 0.0025
    assert TR.learning_rate(10 ** 7, 128, 0.005, 30 * 128 * 10, 0.5) == 1e-5
    assert TR.bn_decay(0, 128, 0.5, 30 * 128 * 10, 0.5, 0.99) == 0.5
    assert TR.bn_decay(10 ** 7, 128, 0.5, 30 * 128 * 10, 0.5, 0.99) == 0.99
    p, g = {"w": np.array([1.0, -2.0])}, {"w": np.array([0.1, -0.3])}
    m, v = {"w": np.zeros(2)}, {"w": np.zeros(2)}
    p, m, v = TR.adam_step(p, g, m, v, 0.01, 1)
    # first Adam step moves each weight by ~lr against the gradient sign
    np.testing.assert_allclose(p["w"], [1.0 - 0.01, -2.0 + 0.01], atol=1e-6)


def test_host_decode_quirk_q1():
    nb = 4
    logits = np.zeros((2, 2 * nb), np.float32)
    logits[0, 3] = 5.0          # class 3 -> 3*pi/2 > pi -> wraps
    logits[0, nb + 3] = 0.25    # residual added UNSCALED
    logits[1, 1] = 5.0
    logits[1, nb + 1] = -0.5
    a = NF.classLogits2angle(logits, nb)
    assert abs(a[0] - (3 * math.pi / 2 + 0.25 - 2 * math.pi)) < 1e-6
    assert abs(a[1] - (math.pi / 2 - 0.5)) < 1e-6
    s = NF.get_angles(logits, nb)
    assert abs(s[1] - (math.pi / 2 - 0.5 * math.pi / nb)) < 1e-6
