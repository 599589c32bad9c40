"""Pins the oracle to the REFERENCE'S OWN CODE.

`tests/golden/reference_*.npz` were produced by importing the unmodified /root/reference
`models/tp8.py` + `utils/tf_util.py` (and `tp_utils/pointcloud.py`, `utils/eulerangles.py`) and
executing them on the TF1 shim (`oracle/tf1_shim`; generator: tests/golden/make_reference_golden.py).
Here the two oracle restatements are held to those vectors; where /root/reference is present (the
build container) the fixtures are additionally regenerated live and the shim's primitives are
checked against independent formulations.
"""
import math
import os
import sys

import numpy as np
import pytest
import torch

from oracle import arch as A, np_forward as NF, reference_run as RR, rigid as RG, torch_ref as TR
from helpers import GOLDEN, OUTPUT_KEYS, golden_case

CASES = ["tiny_B4_N16", "shipped_B4_N16", "shipped_B32_N200", "default_B32_N64"]
live = pytest.mark.skipif(not RR.available(), reason="/root/reference not present on this box")


def ref_case(name, suffix=""):
    return np.load(os.path.join(GOLDEN, f"reference_{name}{suffix}.npz"))


@pytest.mark.parametrize("name", CASES)
def test_reference_graph_creates_the_variables_the_layout_assumes(name):
    """Names, shapes and trainability of every variable the reference graph creates."""
    r = ref_case(name)
    arch = A.tiny_arch() if name.startswith("tiny") else (A.default_arch() if name.startswith("default") else A.Arch())
    specs = dict(A.trainable_specs(arch))
    assert set(r["trainable"].tolist()) == set(specs)
    assert sorted(r["var_names"].tolist()) == sorted(specs)          # no non-trainable tf variables besides the shadows
    shapes = dict(zip(r["var_names"].tolist(), r["var_shapes"].tolist()))
    for n, shp in specs.items():
        ref_shape = tuple(int(x) for x in shapes[n].split())
        assert int(np.prod(ref_shape)) == int(np.prod(shp)) and ref_shape[-len(shp):][-1] == shp[-1], (n, ref_shape, shp)
        if len(ref_shape) == 4:                                        # conv kernels [1, kw, Cin, Cout]
            assert ref_shape[0] == 1 and ref_shape[1] * ref_shape[2] == shp[0], (n, ref_shape, shp)
    assert r["shadow_names"].tolist() == sorted(n for n, _ in A.state_specs(arch))


@pytest.mark.parametrize("name", CASES)
def test_numpy_oracle_matches_reference_run(name):
    r = ref_case(name)
    g, arch, params, state, batch, masks = golden_case(name)
    ep, _ = NF.get_model(batch["pcs1"], batch["pcs2"], arch, params, state, False)
    for k in OUTPUT_KEYS:
        np.testing.assert_allclose(ep[k], r["f64/eval/" + k], atol=2e-5, rtol=0, err_msg=k)
        np.testing.assert_allclose(ep[k], r["f32/eval/" + k], atol=2e-5, rtol=0, err_msg=k)
        # the committed oracle fixture (what the GPU tests compare with) agrees with the reference run
        np.testing.assert_allclose(g["eval/" + k], r["f64/eval/" + k], atol=2e-5, rtol=0, err_msg=k)
    # host decode through the reference's classLogits2angle (train.py:453-456); exact where the three
    # arg-maxes agree between the fp32 oracle and the fp64 run
    ours = NF.pred_angles(ep, arch.num_bins)
    same = np.ones(len(ours), bool)
    for k in ("pred_pc1angle_logits", "pred_pc2angle_logits", "pred_remaining_angle_logits"):
        same &= ep[k][:, :arch.num_bins].argmax(1) == r["f64/eval/" + k][:, :arch.num_bins].argmax(1)
    assert same.mean() > 0.9
    np.testing.assert_allclose(ours[same], r["f64/eval/pred_angles"][same], atol=5e-5)
    ep, _ = NF.get_model(batch["pcs1"], batch["pcs2"], arch, params, state, True, 0.5, masks)
    for k in OUTPUT_KEYS:
        np.testing.assert_allclose(ep[k], r["f64/train/" + k], atol=2.5e-4, rtol=0, err_msg=k)


@pytest.mark.parametrize("name", CASES)
def test_torch_oracle_matches_reference_run_fp64(name):
    """Same graph in double precision: outputs, loss, EMA shadows and every gradient agree to rounding."""
    r = ref_case(name)
    g, arch, params, state, batch, masks = golden_case(name)
    loss, ep64, grads, st64 = TR.loss_and_grads(batch, arch, params, state, 0.5, masks)
    assert abs(loss - float(r["f64/train/loss"])) < 1e-10
    for k in OUTPUT_KEYS:
        np.testing.assert_allclose(ep64[k], r["f64/train/" + k], atol=1e-10, rtol=0, err_msg=k)
    for k in [k for k in r.files if k.startswith("gradnorm/")]:
        n = k[9:]
        got = float(np.sqrt((grads[n].astype(np.float64) ** 2).sum()))
        assert abs(got - float(r[k])) <= 1e-9 + 1e-8 * float(r[k]), n
    for k in [k for k in r.files if k.startswith("grad/")]:
        np.testing.assert_allclose(grads[k[5:]].reshape(r[k].shape), r[k], atol=1e-9, rtol=1e-8, err_msg=k)
    for k in [k for k in r.files if k.startswith("state/")]:
        np.testing.assert_allclose(st64[k[6:]], r[k], atol=1e-10, rtol=0, err_msg=k)
    # fp32 run of the reference vs its fp64 run: the noise floor the 1e-4 tolerance has to live above
    worst = max(float(np.abs(r["f32/eval/" + k] - r["f64/eval/" + k]).max()) for k in OUTPUT_KEYS)
    assert worst < 5e-5, worst


def test_loss_without_inverted_angle_matches_reference_run():
    r = ref_case("tiny_B4_N16", "_noinv")
    g, arch, params, state, batch, masks = golden_case("tiny_B4_N16")
    arch.accept_inverted_angle = False
    loss, ep64, grads, _ = TR.loss_and_grads(batch, arch, params, state, 0.5, masks)
    assert abs(loss - float(r["f64/train/loss"])) < 1e-10
    for k in [k for k in r.files if k.startswith("grad/")]:
        np.testing.assert_allclose(grads[k[5:]].reshape(r[k].shape), r[k], atol=1e-9, rtol=1e-8, err_msg=k)
    with_inv = ref_case("tiny_B4_N16")
    assert float(with_inv["f64/train/loss"]) >= float(r["f64/train/loss"])     # Q5: keeps the LARGER loss


def test_rigid_oracle_matches_reference_functions():
    """a17-a20: oracle/rigid.py against get_mat_angle / transform_points /
    translate_transform_to_new_center_of_rotation (pointcloud.py:279-318) and euler2mat (eulerangles.py:98)."""
    r = np.load(os.path.join(GOLDEN, "reference_rigid.npz"))
    n = len(r["theta"])
    for i in range(n):
        m = RG.get_mat_angle(r["t"][i], float(r["theta"][i]), r["c"][i])
        np.testing.assert_allclose(m, r["mats"][i], atol=1e-12)
        np.testing.assert_allclose(RG.rot_z(float(r["theta"][i])), r["rz"][i], atol=1e-12)
        hom = np.concatenate([r["pts"][i], np.ones((r["pts"].shape[1], 1))], axis=1)
        np.testing.assert_allclose(RG.transform_points(hom.copy(), m), r["moved"][i], atol=1e-12)
        np.testing.assert_allclose(RG.rigid_apply(r["pts"][i], r["t"][i], float(r["theta"][i]), r["c"][i]),
                                   r["moved"][i][:, :3], atol=1e-12)
        np.testing.assert_allclose(RG.transform_points(hom.copy(), [m, r["mats"][(i + 1) % n]]), r["composed"][i], atol=1e-11)
    t_new = RG.translate_transform_to_new_center_of_rotation(r["t"], r["theta"][:, None], r["c"], r["new_c"])
    np.testing.assert_allclose(t_new, r["t_new"], atol=1e-12)


# ------------------------------------------------------------------------------------------------
# live: only where /root/reference is mounted
# ------------------------------------------------------------------------------------------------
@live
def test_live_reference_run_reproduces_fixture():
    r = ref_case("tiny_B4_N16")
    g, arch, params, state, batch, masks = golden_case("tiny_B4_N16")
    out = RR.run(batch, arch, params, state, True, 0.5, masks, double=True, with_grads=True)
    assert abs(out["loss"] - float(r["f64/train/loss"])) < 1e-12
    for k in OUTPUT_KEYS:
        np.testing.assert_allclose(out["end_points"][k], r["f64/train/" + k], atol=1e-12, rtol=0)
    ev = RR.run(batch, arch, params, state, False, double=False)
    for k in OUTPUT_KEYS:
        np.testing.assert_allclose(ev["end_points"][k], r["f32/eval/" + k], atol=1e-6, rtol=0)
    # eval mode reads the shadows and leaves them alone; train mode applies the EMA update (tf_util.py:475-480)
    for k, v in ev["new_state"].items():
        np.testing.assert_array_equal(v, state[k])
    k0 = "siamese/transformer1/embedding/conv1/bn/moments/Squeeze/ExponentialMovingAverage"
    assert np.abs(out["new_state"][k0] - state[k0]).max() > 1e-3


@live
def test_live_placeholder_contract():
    """a1: placeholder_inputs (tp8.py:13-23) -> the 8 feeds with the shapes the engine's API takes."""
    tf, tp8, _ = RR.load()
    RR.configure("SynthCars")
    ph = tp8.placeholder_inputs(5, 7)
    assert [tuple(p.shape) for p in ph] == [(5, 7, 3), (5, 7, 3), (5, 3), (5, 1), (5, 3), (5, 3), (5, 1), (5, 1)]
    from alignnet_b200 import tp8 as ours
    assert ours.placeholder_inputs.__code__.co_varnames[:2] == tp8.placeholder_inputs.__code__.co_varnames[:2]
    for fn in ("get_model", "get_loss", "classLogits2angle"):
        ref_args = getattr(tp8, fn).__code__.co_varnames[:getattr(tp8, fn).__code__.co_argcount]
        our_args = getattr(ours, fn).__code__.co_varnames[:getattr(ours, fn).__code__.co_argcount]
        if fn != "get_loss":       # the reference's get_loss is *args over _get_loss_separate's nine arguments
            assert ref_args == our_args, (fn, ref_args, our_args)
    sep = tp8._get_loss_separate.__code__
    assert ours.get_loss.__code__.co_varnames[:ours.get_loss.__code__.co_argcount] == sep.co_varnames[:sep.co_argcount]


@live
def test_p2p_loss_quirk_q6_matches_reference_code():
    """a21: the reference's own _get_loss_p2p (executed) against the closed form the oracle states for it."""
    g, arch, params, state, batch, masks = golden_case("tiny_B4_N16")
    for inv in (True, False):
        arch.accept_inverted_angle = inv
        out = RR.run(batch, arch, params, state, False, double=True, loss="p2p")
        ep = {k: torch.tensor(v) for k, v in out["end_points"].items()}
        ours = TR.get_loss_p2p(torch.tensor(batch["pcs1"], dtype=torch.float64),
                               torch.tensor(batch["pc1_centers"], dtype=torch.float64), ep)
        assert abs(float(ours) - out["loss"]) < 1e-10 * max(1.0, abs(out["loss"])), (float(ours), out["loss"])
    assert out["loss"] > 0


@live
def test_transform_pcs_quirk_q6_matches_reference_code():
    """a21: the reference's own tf_transform_pcs (executed on the shim, every None-combination) against the oracle's
    statement-by-statement restatement; and the p2p loss built on it against the closed form pinned above."""
    from oracle import rigid as RG
    tf, tp8, _ = RR.load()
    RR.configure("SynthCars")
    rng = np.random.default_rng(3)
    B, N = 4, 9
    pcs, t, c = rng.normal(size=(B, N, 3)), rng.normal(size=(B, 3)), rng.normal(size=(B, 3))
    a = rng.uniform(-3, 3, size=(B,))
    for args in [(t, a, c), (None, a, None), (t, None, None), (None, None, c), (t, a, None), (None, None, None)]:
        tf.reset()
        out = tp8.tf_transform_pcs(tf.constant(pcs), *[None if x is None else tf.constant(x) for x in args])
        val = np.asarray(out.numpy() if hasattr(out, "numpy") else getattr(out, "value", out), dtype=np.float64)
        np.testing.assert_allclose(val, RG.tf_transform_pcs(pcs, *args), atol=2e-6)
    # with rotation centres the cloud collapses to tile(centres): translation and angle drop out (quirk Q6)
    np.testing.assert_array_equal(RG.tf_transform_pcs(pcs, t, a, c), np.repeat(c[:, None, :], N, axis=1))
    ep = {"pred_s2_pc1centers": torch.tensor(c + 0.1 * rng.normal(size=c.shape))}
    per, loss = RG.loss_p2p(pcs, t, a, ep["pred_s2_pc1centers"].numpy(), t, a[:, None], c)
    closed = float(TR.get_loss_p2p(torch.tensor(pcs), torch.tensor(c), ep))
    assert abs(per - closed) < 1e-12 and abs(loss - per * B) < 1e-12


@live
def test_shim_scoping_follows_tf1():
    tf, _, _ = RR.load()
    tf.reset()
    with tf.variable_scope("siamese"):
        with tf.variable_scope("conv1"):
            w0 = tf.get_variable("weights", [2, 3], initializer=tf.constant_initializer(1.0))
            b0 = tf.Variable(tf.constant(0.0, shape=[3]), name="beta")
    with tf.variable_scope("siamese", reuse=tf.AUTO_REUSE):
        with tf.variable_scope("conv1"):
            w1 = tf.get_variable("weights", [2, 3], initializer=tf.constant_initializer(1.0))
            b1 = tf.Variable(tf.constant(0.0, shape=[3]), name="beta")
    with tf.variable_scope(""):
        with tf.variable_scope("fc1"):
            tf.get_variable("weights", [2, 3], initializer=tf.constant_initializer(1.0))
    assert w0 is w1 and b0 is not b1
    assert sorted(tf.variables()) == ["fc1/weights", "siamese/conv1/beta", "siamese/conv1/weights", "siamese_1/conv1/beta"]
    with pytest.raises(ValueError):
        with tf.variable_scope("siamese"):
            with tf.variable_scope("conv1"):
                tf.get_variable("weights", [2, 3], initializer=tf.constant_initializer(1.0))
    tf.reset()


@live
def test_shim_primitives_against_independent_formulations():
    tf, _, _ = RR.load()
    tf.reset()
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.normal(size=(2, 5, 3, 1))).as_subclass(tf.Tensor)
    w = torch.tensor(rng.normal(size=(1, 3, 1, 4)))
    y = tf.nn.conv2d(x, w, [1, 1, 1, 1], padding="VALID")               # first layer: [1,3] window over xyz
    np.testing.assert_allclose(y.numpy()[:, :, 0, :], np.einsum("bnk,kc->bnc", x.numpy()[..., 0], w.numpy()[0, :, 0, :]), atol=1e-12)
    m, v = tf.nn.moments(y, [0, 1, 2])
    np.testing.assert_allclose(m.numpy(), y.numpy().reshape(-1, 4).mean(0), atol=1e-12)
    np.testing.assert_allclose(v.numpy(), y.numpy().reshape(-1, 4).var(0), atol=1e-12)      # biased
    bn = tf.nn.batch_normalization(y, m, v, torch.ones(4) * 0.5, torch.ones(4) * 2.0, 1e-3)
    np.testing.assert_allclose(bn.numpy(), (y.numpy() - m.numpy()) / np.sqrt(v.numpy() + 1e-3) * 2.0 + 0.5, atol=1e-12)
    p = tf.nn.max_pool(y, ksize=[1, 5, 1, 1], strides=[1, 2, 2, 1], padding="VALID")
    np.testing.assert_allclose(p.numpy()[:, 0, 0, :], y.numpy()[:, :, 0, :].max(1), atol=0)
    assert float(tf.mod(torch.tensor(-0.5), 2.0 * math.pi)) == pytest.approx(2 * math.pi - 0.5)
    assert int(tf.argmax(torch.tensor([[1.0, 3.0, 3.0]]), axis=1)[0]) == 1            # first maximal index
    assert int(tf.to_int32(torch.tensor(2.9))) == 2 and int(tf.to_int32(torch.tensor(-2.9))) == -2
    ema = tf.train.ExponentialMovingAverage(decay=0.9)
    ema.apply([m])
    np.testing.assert_allclose(ema.average(m).numpy(), 0.1 * m.numpy(), atol=1e-12)       # zero-initialised shadow (Q7)
    tf.reset()
