"""The reference's second optimiser branch (train.py:211-212): `tf.train.MomentumOptimizer(learning_rate,
momentum=cfg.training.optimizer.momentum)` -- TF's ApplyMomentum with use_nesterov = False:
`accum = momentum * accum + g ; var -= lr * accum`, one slot per variable named `<var>/Momentum`.

CPU part: the oracle's restatement against a hand-computed sequence, the config surface, the C-ABI argument checks.
GPU part (-m gpu): `an3d_momentum_step` against the oracle through eager steps and through the CUDA-graph replay, and the
optimiser slot through a checkpoint round trip."""
import ctypes as C
import json

import numpy as np
import pytest

from oracle import arch as A, torch_ref as TR
from helpers import engine_arch


def test_oracle_momentum_known_answer():
    p, g, acc = {"w": np.array([1.0, -2.0])}, {"w": np.array([0.5, -1.0])}, {"w": np.zeros(2)}
    p, acc = TR.momentum_step(p, g, acc, lr=0.1, momentum=0.9)
    np.testing.assert_allclose(acc["w"], [0.5, -1.0])                      # accum = 0.9 * 0 + g
    np.testing.assert_allclose(p["w"], [0.95, -1.9])                       # var -= 0.1 * accum
    p, acc = TR.momentum_step(p, g, acc, lr=0.1, momentum=0.9)
    np.testing.assert_allclose(acc["w"], [0.95, -1.9])                     # 0.9 * 0.5 + 0.5
    np.testing.assert_allclose(p["w"], [0.855, -1.71])
    # momentum = 0 is plain gradient descent
    p0, a0 = TR.momentum_step({"w": np.array([1.0])}, {"w": np.array([2.0])}, {"w": np.array([7.0])}, 0.5, 0.0)
    assert float(p0["w"][0]) == 0.0 and float(a0["w"][0]) == 2.0


def test_config_selects_the_optimizer(tmp_path):
    from alignnet_b200 import config

    def load(overlay):
        path = tmp_path / "Run.json"
        path.write_text(json.dumps(overlay))
        config.reset_config()
        return config.load_config(str(path))

    assert config.optimizer_from_config(load({})) == ("adam", None)                        # configs/default.json:33-35
    assert config.optimizer_from_config(load({"training": {"optimizer": {"optimizer": "momentum", "momentum": 0.9}}})) \
        == ("momentum", 0.9)
    with pytest.raises(ValueError):                                                        # train.py:212 would KeyError
        config.validate(load({"training": {"optimizer": {"optimizer": "momentum"}}}))
    with pytest.raises(ValueError):                                                        # train.py:215-216 assert False
        config.validate(load({"training": {"optimizer": {"optimizer": "rmsprop"}}}))
    config.reset_config()


def test_abi_argument_checks():
    from alignnet_b200 import _lib
    lib = _lib.load()
    buf = (C.c_float * 72)()
    assert lib.an3d_momentum_step(None, buf, buf, 64, 0.1, 0.9, 1.0, None) == -1           # AN3D_ERR_INVALID
    assert lib.an3d_momentum_step(buf, buf, buf, -1, 0.1, 0.9, 1.0, None) == -1
    base = C.addressof(buf)
    base += (-base) % 16
    assert lib.an3d_momentum_step(C.c_void_p(base + 4), C.c_void_p(base), C.c_void_p(base), 8, 0.1, 0.9, 1.0, None) == -6   # _ALIGN
    import torch
    if not torch.cuda.is_available():                                                      # no CPU fallback
        assert lib.an3d_momentum_step(buf, buf, buf, 64, 0.1, 0.9, 1.0, None) in (-3, -5)
        assert lib.an3d_last_error()


def _engine(arch, params, state, precision="fp32"):
    from alignnet_b200 import engine
    e = engine.Engine(engine_arch(arch), "cuda:0", precision)
    e.set_params(params)
    e.set_state(state)
    return e


@pytest.mark.gpu
def test_momentum_step_matches_the_oracle():
    import torch
    arch = A.tiny_arch()
    e = _engine(arch, A.init_params(arch, 0), A.init_state(arch))
    with pytest.raises(RuntimeError):
        e.momentum_step(0.1)                                   # no slot before set_optimizer
    with pytest.raises(ValueError):
        e.set_optimizer("momentum")                            # train.py:212 needs cfg.training.optimizer.momentum
    with pytest.raises(ValueError):
        e.set_optimizer("rmsprop")                             # train.py:215
    e.set_optimizer("momentum", momentum=0.9)
    rng = np.random.default_rng(0)
    n = e.params.numel()
    p = {"w": e.params.cpu().numpy().astype(np.float64)}
    acc = {"w": np.zeros(n)}
    for _ in range(4):
        gnp = rng.normal(size=n).astype(np.float32) * 0.01
        e.grads.copy_(torch.from_numpy(gnp))
        e.optimizer_step(0.05, grad_scale=0.5)                 # data-parallel form: 1/G folded into the update
        p, acc = TR.momentum_step(p, {"w": gnp.astype(np.float64) * 0.5}, acc, 0.05, 0.9)
    torch.cuda.synchronize()
    np.testing.assert_allclose(e.params.cpu().numpy(), p["w"], atol=2e-6)
    np.testing.assert_allclose(e.mom_accum.cpu().numpy(), acc["w"], atol=1e-7)
    assert e.step == 4
    assert not e.adam_m.any() and not e.adam_v.any()           # the Adam slots stay untouched


@pytest.mark.gpu
def test_training_with_momentum_eager_graph_and_checkpoint(tmp_path):
    """Three optimiser steps: oracle gradients + oracle momentum update vs the engine (fp32 mode), the same through the
    CUDA-graph path, and the `<var>/Momentum` slots through a checkpoint."""
    import torch
    from alignnet_b200 import synth, tf_checkpoint
    from helpers import MASK_KEYS

    def to_dev(d):
        return {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in d.items()}

    arch = A.tiny_arch()
    params, state = A.init_params(arch, 3), A.init_state(arch)
    e = _engine(arch, params, state)
    e.set_optimizer("momentum", momentum=0.8)
    lr = 1e-3
    p = {k: v.astype(np.float64) for k, v in params.items()}
    st = {k: v.astype(np.float64) for k, v in state.items()}
    acc = {k: np.zeros_like(v) for k, v in p.items()}
    gmax = {k: 0.0 for k in p}
    for step in range(1, 4):
        batch = synth.make_batch(8, 32, seed=100 + step)
        rng = np.random.default_rng(step)
        masks = {k: (rng.uniform(size=(8, 8)) < 0.7).astype(np.float32) for k in MASK_KEYS}
        loss_ref, _, grads, st = TR.loss_and_grads(batch, arch, p, st, 0.5, masks)
        for k in grads:
            gmax[k] = max(gmax[k], float(np.abs(grads[k]).max()))
        p, acc = TR.momentum_step(p, grads, acc, lr, 0.8)
        loss = e.train_step(to_dev(batch), lr, 0.5, masks=to_dev(masks))
        assert abs(float(loss[0].cpu()) - loss_ref) < 2e-3 * max(1.0, abs(loss_ref)), (step, float(loss[0].cpu()), loss_ref)
    got = e.get_params()
    for name, ref in p.items():
        # The update is linear in the gradient, so the engine's parameters may sit as far from the oracle's as its
        # gradients do (tests/test_gpu_parity.py: ~1e-2 of a tensor's largest gradient element where ReLU masks flip, plus
        # what the slightly different parameters of steps 2 and 3 add): 5 % of the largest movement of the tensor, against
        # which a wrong update rule is off by 100 % ((1 - momentum)-damped accumulators: x0.2; Nesterov: x1.8)
        moved = float(np.abs(ref - params[name].astype(np.float64)).max())
        err = float(np.abs(got[name].reshape(ref.shape) - ref).max())
        assert err <= 5e-2 * moved + 5e-2 * lr * 5.24 * 1e-3 * max(gmax.values()) + 1e-6, (name, err, moved)
    # checkpoint: one slot per variable under TF's name, no Adam slots; restoring brings the accumulators back
    prefix = str(tmp_path / "mom")
    tf_checkpoint.save_from_engine(e, prefix)
    ck = tf_checkpoint.read_checkpoint(prefix)
    assert "fc1/weights/Momentum" in ck and not any(k.endswith("/Adam") for k in ck) and "beta1_power" not in ck
    assert int(ck["Variable"]) == 3
    e2 = _engine(arch, A.init_params(arch, 5), A.init_state(arch))
    e2.set_optimizer("momentum", momentum=0.8)
    tf_checkpoint.load_into_engine(e2, prefix)
    torch.testing.assert_close(e2.mom_accum, e.mom_accum, rtol=0, atol=0)
    torch.testing.assert_close(e2.params, e.params, rtol=0, atol=0)
    assert e2.step == 3
    # graph replay == eager steps (dropout masks come from the device seed there: compare with keep_prob = 1)
    arch1 = A.tiny_arch(s1_keep=1.0, s2_keep=1.0, head_keep=1.0)
    a, b = _engine(arch1, params, state), _engine(arch1, params, state)
    for x in (a, b):
        x.set_optimizer("momentum", momentum=0.8)
    dev = to_dev(synth.make_batch(8, 32, seed=7))
    for _ in range(3):
        a.train_step(dev, lr=lr, bn_decay=0.5)
        b.train_step_graph(dev, lr=lr, bn_decay=0.5)
    torch.cuda.synchronize()
    assert a.step == b.step == 3
    # fp32 atomics in the backward's weight-gradient sums (tests/test_gpu_determinism.py): equal up to their noise
    d = (b.params - a.params).abs().max().item()
    assert d <= 1e-4, d
    assert (b.mom_accum - a.mom_accum).abs().max().item() <= 1e-3 * max(1.0, a.mom_accum.abs().max().item())
