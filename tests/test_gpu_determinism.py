"""Run-to-run reproducibility of the bf16 fast mode (the mode bench.py times).

Round 1 measured a gradient cosine of only 0.979 between two identical c3 steps: fp32 atomics added the per-CTA
partial sums of the BN statistics (Gram matrices) and the split-K slices of the 3- / 103-wide FC layers in arrival
order, a last-bit difference moved a bf16 rounding or an arg-max downstream, and the model is discontinuous there.
Now every cross-CTA fp32 sum of the forward pass goes through per-CTA slots added in a fixed order
(`sum_parts_kernel`), training never splits K, and the remaining atomics of the forward are fp64 sums that are
rounded to fp32 afterwards (order-dependent in the 16th digit only) or exact (`atomicMax`).  So:
  * two training forwards on the same inputs return IDENTICAL bits (outputs, loss, moving averages);
  * the backward still reduces weight gradients with fp32 atomics, but nothing discrete depends on them: gradient
    cosine >= 0.9999 (measured: 1 - 3e-7), every element within 5e-3 of the largest (measured 9e-4);
  * inference is bit-reproducible with Engine(deterministic=True) / AN3D_DETERMINISTIC (no split-K reductions)."""
import numpy as np
import pytest
import torch

from oracle import arch as A
from helpers import OUTPUT_KEYS, engine_arch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _build():
    import __graft_entry__ as ge
    ge.build()


def _dev(batch):
    return {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in batch.items()}


def _engines(n, **kw):
    from alignnet_b200 import engine
    arch = A.Arch()
    params, state = A.randomize_for_test(arch, A.init_params(arch, 70), A.init_state(arch), 71)
    out = []
    for _ in range(n):
        e = engine.Engine(engine_arch(arch), "cuda:0", "bf16", **kw)
        e.set_params(params)
        e.set_state(state)
        out.append(e)
    return out


@pytest.mark.parametrize("B,N", [(4096, 200), (96, 512)])
def test_training_step_is_reproducible(B, N):
    from alignnet_b200 import synth
    dev = _dev(synth.make_batch_fast(B, N, seed=72))
    runs = []
    for e in _engines(2):
        ep = e.forward(dev["pcs1"], dev["pcs2"], True, 0.5, None, seed=5)
        loss = e.backward(dev["pcs1"], dev["pcs2"], dev, ep)
        torch.cuda.synchronize()
        runs.append(({k: ep[k].clone() for k in OUTPUT_KEYS}, loss.clone(), e.bn_state.clone(), e.grads.clone()))
    (ep_a, l_a, s_a, g_a), (ep_b, l_b, s_b, g_b) = runs
    for k in OUTPUT_KEYS:
        assert torch.equal(ep_a[k], ep_b[k]), k                         # bit for bit
    assert torch.equal(s_a, s_b)                                        # moving averages
    assert float(l_a[0]) == float(l_b[0])
    cos = float(torch.dot(g_a.double(), g_b.double()) / (g_a.double().norm() * g_b.double().norm()))
    rel = float((g_a - g_b).abs().max() / g_a.abs().max())
    print(f"B={B} N={N}: gradient cosine between two runs 1 - {1 - cos:.2e}, max |diff| / max |g| = {rel:.2e}")
    assert cos >= 0.9999 and rel <= 5e-3, (cos, rel)


def test_optimiser_trajectories_stay_together():
    """Three optimiser steps on two engines.  The first loss is identical (bit-reproducible forward).  From the second
    step on the two runs differ a little: the backward still adds weight gradients with fp32 atomics (relative 1e-3 of
    the largest element), and Adam's first updates are sign-like (m / sqrt(v) = +-1 at t = 1), so a gradient element at
    the noise level can move its weight by +lr in one run and -lr in the other (measured: the second loss of two runs
    differs by 1.4 %).  Bounds: losses within 5 %, no weight further apart than 2 lr per step taken.  Bit-reproducible
    TRAINING would need ordered sums in the six backward kernels that still reduce with fp32 atomics."""
    from alignnet_b200 import synth
    dev = _dev(synth.make_batch_fast(1024, 200, seed=73))
    ea, eb = _engines(2)
    lr = 1e-3
    la = [float(ea.train_step(dev, lr=lr, bn_decay=0.5, seed=i)[0].cpu()) for i in range(3)]
    lb = [float(eb.train_step(dev, lr=lr, bn_decay=0.5, seed=i)[0].cpu()) for i in range(3)]
    assert la[0] == lb[0], (la, lb)
    for a, b in zip(la, lb):
        assert abs(a - b) <= 5e-2 * abs(b), (la, lb)
    assert float((ea.params - eb.params).abs().max()) <= 2 * 3 * lr * 1.05


def test_deterministic_inference_flag():
    from alignnet_b200 import synth
    dev = _dev(synth.make_batch_fast(1024, 200, seed=74))
    outs = []
    for e in _engines(2, deterministic=True):
        ep = e.forward(dev["pcs1"], dev["pcs2"], False)
        ep = e.forward(dev["pcs1"], dev["pcs2"], False)                 # second call: cached folds
        torch.cuda.synchronize()
        outs.append({k: ep[k].clone() for k in OUTPUT_KEYS})
    for k in OUTPUT_KEYS:
        assert torch.equal(outs[0][k], outs[1][k]), k


@pytest.mark.parametrize("precision", ["bf16", "bf16x3", "bf16x6"])
def test_layer_by_layer_tensor_core_path_reproducibility(precision):
    """The split-operand GEMM path (csrc/gemm_tc.cuh; the reference's default architecture takes it in bf16 mode).
    bf16 (one image per operand): the forward and dgrad forms never slice K across CTAs, so two inference calls return
    identical bits -- with bf16 roundings downstream a last-bit difference would flip arg-max bins.  bf16x3 / bf16x6
    slice K of under-filled GEMMs on purpose (short tensor-core accumulations keep them at fp32 grade) and meet in fp32
    reductions: two calls agree to 5e-5 in inference (measured 1e-5) and to 3e-4 (six products) / 1e-3 (three) in training mode."""
    from alignnet_b200 import engine, synth
    arch = A.Arch(num_bins=36, s1_conv=(128, 128, 256), s2_conv=(64, 64, 64, 128, 1024), emb_conv=(64, 64, 64, 128, 1024),
                  accept_inverted_angle=False, early_stage_factor=0.1)
    params, state = A.randomize_for_test(arch, A.init_params(arch, 80), A.init_state(arch), 81)
    batch = _dev(synth.make_batch_fast(160, 200, seed=82))
    runs = []
    for _ in range(2):
        e = engine.Engine(engine_arch(arch), "cuda:0", precision)
        e.set_params(params)
        e.set_state(state)
        ev = {k: v.clone() for k, v in e.forward(batch["pcs1"], batch["pcs2"], False).items()}
        tr = {k: v.clone() for k, v in e.forward(batch["pcs1"], batch["pcs2"], True, 0.5, None, seed=3).items()}
        torch.cuda.synchronize()
        runs.append((ev, tr))
    same_bins = torch.ones(160, dtype=torch.bool, device="cuda")
    for k in ("pred_pc1angle_logits", "pred_pc2angle_logits"):
        for i in (0, 1):
            same_bins &= runs[0][i][k][:, :36].argmax(1) == runs[1][i][k][:, :36].argmax(1)
    for k in OUTPUT_KEYS:
        if precision == "bf16":
            assert torch.equal(runs[0][0][k], runs[1][0][k]), k
        else:
            # (training: batch-statistics BN over 160 samples amplifies the last bits: measured up to 2.6e-4 with three
            # products, 2e-5 with six)
            for i, bound in ((0, 5e-5), (1, 1e-3 if precision == "bf16x3" else 3e-4)):
                d = (runs[0][i][k] - runs[1][i][k]).abs()
                assert d[same_bins].max().item() < bound and same_bins.float().mean().item() > 0.98, (k, i)
