"""The CUDA path against the CPU oracle at BASELINE.json's full sizes -- not properties, the oracle itself.

  * c2 (B=1024, N=200, eval forward): fp32 engine <= 1e-4 abs of the fp64 oracle on all 8 end_points and on the
    host-decoded pred_angles (north_star); bf16 engine inside the bf16 mode's stated bound.
  * c2-size training step (B=1024, N=200): loss and every parameter gradient of the fp32 engine against torch-CPU
    autograd, gradients of the bf16 engine against the same oracle (cosine per tensor + of the whole vector).
  * c3 (B=4096, N=200) training-mode forward + loss: fp32 and bf16 engines against the fp32 oracle run without
    autograd.  (fwd+bwd through torch autograd at B=4096 holds ~60 GB of [M,1024] activations on the host, so the
    gradient check is the B=1024 one; the loss at B=4096 covers the [B,B] loss couplings at full size.)
  * the c4 / c5 cloud sizes N=512 / N=1024 (B=64) end to end, eval and training, both precisions.
  * the split-operand tensor-core modes in every case above: bf16x6 held to the fp32 mode's tolerances everywhere,
    bf16x3 to the fp32 tolerance in eval mode and to its own bound behind batch-statistics BN.
  * the reference's default architecture (configs/default.json:13-15: [128,128,256] and two five-layer stacks, 36 bins,
    no inverted-angle acceptance), which the fused kernels do not cover: fp32 / bf16x6 to the parity tolerance, bf16
    (layer-by-layer tensor-core path) to the fast mode's bound.
Oracle cost on 8 host cores: ~8 s (fp64 eval, B=1024), ~2 x 16 s (fp32 fwd+bwd with and without the bf16 rounding model, B=1024), ~25 s (fp32 forward, B=4096).
Measured numbers are printed (run with -s) and written to gpurun_out/parity_fullsize.json when that directory exists.
The file name sorts last on purpose: these are the slowest GPU tests."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import arch as A, np_forward as NF, torch_ref as TR
from helpers import MASK_KEYS, OUTPUT_KEYS, engine_arch, top2_margin

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RECORD = {}

TOL_FP32 = 1e-4            # north_star: abs, eval mode
TOL_FP32_TRAIN = 2.5e-4    # batch-statistics BN amplifies fp32 rounding (tests/test_gpu_parity.py header)
PARITY = ("fp32", "bf16x6")  # the modes held to the parity tolerances: CUDA-core fp32 and six-product split bf16 on tensor cores
X3_TRAIN_MAX = 5e-3        # three-product split bf16 (~2^-18 per product) behind batch-statistics BN; eval mode meets TOL_FP32
BF16_MAX, BF16_MEAN = 1.2e-1, 1.5e-2          # tests/test_gpu_bf16.py: eval-mode bound of the fast mode
BF16_MAX_TRAIN, BF16_MEAN_TRAIN = 8e-1, 2.5e-1


@pytest.fixture(scope="module", autouse=True)
def _build():
    import __graft_entry__ as ge
    ge.build()
    yield
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out) and RECORD:
        with open(os.path.join(out, "parity_fullsize.json"), "w") as f:
            json.dump(RECORD, f, indent=1)


def _note(name, **kw):
    RECORD[name] = kw
    print(name, kw)


def _engine(arch, params, state, precision):
    from alignnet_b200 import engine
    e = engine.Engine(engine_arch(arch), "cuda:0", precision)
    e.set_params(params)
    e.set_state(state)
    return e


def _dev(d):
    return {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in d.items()}


def _case(B, N, seed):
    from alignnet_b200 import synth
    arch = A.Arch()
    params, state = A.randomize_for_test(arch, A.init_params(arch, seed), A.init_state(arch), seed + 1)
    batch = synth.make_batch_fast(B, N, seed=seed + 2)
    rng = np.random.default_rng(seed + 3)
    masks = {k: (rng.uniform(size=(B, 256)) < 0.7).astype(np.float32) for k in MASK_KEYS}
    return arch, params, state, batch, masks


def _oracle_forward(arch, params, state, batch, training, masks=None, dtype=torch.float64):
    """end_points (+ loss in training mode) of the oracle without autograd."""
    p, s, b = TR.to_torch(params, dtype), TR.to_torch(state, dtype), TR.to_torch(batch, dtype)
    m = None if masks is None else TR.to_torch(masks, dtype)
    with torch.no_grad():
        ep, _ = TR.get_model(b["pcs1"], b["pcs2"], arch, p, s, training, 0.5 if training else None, m)
        loss = None
        if training:
            loss = float(TR.get_loss(b["translations"], b["rel_angles"], b["pc1_centers"], b["pc2_centers"], b["pc1_angles"],
                                     b["pc2_angles"], ep, arch))
    return {k: v.numpy().astype(np.float64) for k, v in ep.items()}, loss


def _stable_rows(got, ref, nb, margin=None):
    """rows whose stage-2 arg-max bins agree (and, with `margin`, whose oracle top-2 margin exceeds it)"""
    ok = np.ones(ref["pred_translations"].shape[0], bool)
    for k in ("pred_pc1angle_logits", "pred_pc2angle_logits"):
        ok &= got[k][:, :nb].argmax(1) == ref[k][:, :nb].argmax(1)
        if margin is not None:
            ok &= top2_margin(ref[k], nb) > margin
    return ok


def _errors(got, ref, nb, margin=None):
    """(fraction of stable rows, max-abs error, mean-abs error) over the 8 end_points; the two outputs downstream
    of the canonicalisation are compared on the stable rows only (a flipped bin rotates the cloud by 2 pi / nb: a
    discontinuity of the reference function, not an error)."""
    stable = _stable_rows(got, ref, nb, margin)
    worst_max = worst_mean = 0.0
    for k in OUTPUT_KEYS:
        assert np.isfinite(got[k]).all(), k
        d = np.abs(got[k] - ref[k])
        if k in ("pred_translations", "pred_remaining_angle_logits"):
            d = d[stable]
        worst_max, worst_mean = max(worst_max, float(d.max())), max(worst_mean, float(d.mean()))
    return float(stable.mean()), worst_max, worst_mean


def _host(ep):
    torch.cuda.synchronize()
    return {k: ep[k].cpu().numpy().astype(np.float64) for k in OUTPUT_KEYS}


def test_c2_eval_forward_vs_fp64_oracle():
    """BASELINE configs[1] at full size: B=1024, N=200, eval forward, both precisions, against the fp64 oracle."""
    arch, params, state, batch, _ = _case(1024, 200, 300)
    nb = arch.num_bins
    ref, _ = _oracle_forward(arch, params, state, batch, False)
    dev = _dev(batch)
    # fp32 parity mode: 1e-4 abs everywhere; downstream outputs where the oracle's arg-max is not a near-tie
    e32 = _engine(arch, params, state, "fp32")
    got = _host(e32.forward(dev["pcs1"], dev["pcs2"], False))
    frac, emax, emean = _errors(got, ref, nb, margin=1e-3)
    _note("c2_eval_fp32", stable=frac, max_abs=emax, mean_abs=emean)
    assert frac > 0.97 and emax <= TOL_FP32, (frac, emax)
    # host decode of the reported angle (train.py:453-456, quirk Q1) on rows where all three arg-maxes are unambiguous
    ok = np.ones(1024, bool)
    for k in ("pred_pc1angle_logits", "pred_pc2angle_logits", "pred_remaining_angle_logits"):
        ok &= top2_margin(ref[k], nb) > 1e-3
    ok &= _stable_rows(got, ref, nb)
    ref_ang = NF.pred_angles(ref, nb)
    got_ang = e32.pred_angles(e32.forward(dev["pcs1"], dev["pcs2"], False)).cpu().numpy()
    aerr = float(np.abs(got_ang[ok] - ref_ang[ok]).max())
    _note("c2_eval_fp32_pred_angles", rows=int(ok.sum()), max_abs=aerr)
    assert ok.mean() > 0.9 and aerr <= 2 * TOL_FP32, aerr        # three decoded logits add up
    # the same tolerance on the tensor cores: every GEMM as three / six bf16 products of split operands
    for prec in ("bf16x3", "bf16x6"):
        ex = _engine(arch, params, state, prec)
        got = _host(ex.forward(dev["pcs1"], dev["pcs2"], False))
        frac, emax, emean = _errors(got, ref, nb, margin=1e-3)
        _note("c2_eval_" + prec, stable=frac, max_abs=emax, mean_abs=emean)
        assert frac > 0.97 and emax <= TOL_FP32, (prec, frac, emax)
        del ex
    # bf16 fast mode: its own stated bound, at full size
    e16 = _engine(arch, params, state, "bf16")
    got = _host(e16.forward(dev["pcs1"], dev["pcs2"], False))
    frac, emax, emean = _errors(got, ref, nb)
    _note("c2_eval_bf16", stable=frac, max_abs=emax, mean_abs=emean)
    assert frac > 0.6 and emax < BF16_MAX and emean < BF16_MEAN, (frac, emax, emean)
    # and the replayed graph (what bench.py times) returns the same numbers as the eager call
    g = _host(e16.forward_graph(dev["pcs1"], dev["pcs2"]))
    g = _host(e16.forward_graph(dev["pcs1"], dev["pcs2"]))
    frac_g, emax_g, emean_g = _errors(g, ref, nb)
    assert frac_g > 0.6 and emax_g < BF16_MAX and emean_g < BF16_MEAN, (frac_g, emax_g, emean_g)


def _grad_report(grads, grads_ref):
    """per-tensor (relative max error, cosine) for tensors carrying >= 1e-2 of the largest gradient norm, and the
    cosine of the whole gradient vector"""
    gmax = max(float(np.linalg.norm(v)) for v in grads_ref.values())
    rows, dot, n1, n2 = [], 0.0, 0.0, 0.0
    for n, ref in grads_ref.items():
        g = grads[n].reshape(ref.shape).astype(np.float64)
        assert np.isfinite(g).all(), n
        dot += float((g * ref).sum()); n1 += float((g * g).sum()); n2 += float((ref * ref).sum())
        rn = float(np.linalg.norm(ref))
        if rn < 1e-2 * gmax:
            continue
        rows.append((float(np.abs(g - ref).max() / max(np.abs(ref).max(), 1e-30)),
                     float((g * ref).sum() / (np.linalg.norm(g) * rn + 1e-30)), n))
    return rows, dot / np.sqrt(n1 * n2 + 1e-300)


def test_c2_size_training_step_vs_oracle_autograd():
    """B=1024, N=200 training step: loss, end_points and all gradients against torch-CPU autograd (fp32 arithmetic:
    fp64 autograd at this size holds ~30 GB; the fp32 oracle's own distance from fp64 is the tolerance's floor)."""
    arch, params, state, batch, masks = _case(1024, 200, 310)
    nb = arch.num_bins
    loss_ref, ep_ref, grads_ref, _ = TR.loss_and_grads(batch, arch, params, state, 0.5, masks, dtype=torch.float32)
    ep_ref = {k: v.astype(np.float64) for k, v in ep_ref.items()}
    # the yardstick for the bf16 mode: the SAME oracle with the engine's rounding points (bf16 operands of every conv
    # layer after the first) -- how far bf16 arithmetic alone moves the gradient of this (discontinuous) loss
    TR.SIM_BF16 = True
    try:
        _, _, grads_sim, _ = TR.loss_and_grads(batch, arch, params, state, 0.5, masks, dtype=torch.float32)
    finally:
        TR.SIM_BF16 = False
    _, cos_sim = _grad_report(grads_sim, grads_ref)
    dev, dm = _dev(batch), _dev(masks)
    out = {}
    for prec in ("fp32", "bf16x6", "bf16x3", "bf16"):
        e = _engine(arch, params, state, prec)
        ep = e.forward(dev["pcs1"], dev["pcs2"], True, 0.5, dm)
        loss = e.backward(dev["pcs1"], dev["pcs2"], dev, ep)
        got = _host(ep)
        lv = float(loss[0].cpu())
        frac, emax, emean = _errors(got, ep_ref, nb)
        rows, cos_all = _grad_report(e.get_grads(), grads_ref)
        worst_rel = max(r[0] for r in rows)
        worst_cos = min(r[1] for r in rows)
        _note(f"c2size_train_{prec}", stable=frac, out_max_abs=emax, out_mean_abs=emean, loss=lv, loss_ref=loss_ref,
              grad_cos_all=cos_all, grad_worst_tensor_cos=worst_cos, grad_worst_tensor_relmax=worst_rel)
        out[prec] = (frac, emax, emean, lv, cos_all, worst_cos, worst_rel)
        del e
        torch.cuda.empty_cache()
    # bf16x3 (~2^-18 per product): batch-statistics BN amplifies its error past the parity tolerance; its own bound
    frac, emax, emean, lv, cos_all, worst_cos, worst_rel = out["bf16x3"]
    assert frac > 0.95 and emax <= X3_TRAIN_MAX, (frac, emax)
    assert abs(lv - loss_ref) <= 1e-3 * max(1.0, abs(loss_ref)), (lv, loss_ref)
    assert cos_all > 0.99, cos_all
    for prec in PARITY:     # the two modes held to the parity tolerance
        frac, emax, emean, lv, cos_all, worst_cos, worst_rel = out[prec]
        # two fp32 evaluations of this graph differ through flipped ReLU masks / arg rows (tests/test_gpu_parity.py
        # header: up to 3e-2 of a tensor's max |grad| between the oracle's own fp32 and fp64 runs)
        assert frac > 0.97 and emax <= 2 * TOL_FP32_TRAIN, (prec, frac, emax)
        assert abs(lv - loss_ref) <= 1e-4 * max(1.0, abs(loss_ref)), (prec, lv, loss_ref)
        assert cos_all > 0.999 and worst_cos > 0.99 and worst_rel < 5e-2, (prec, cos_all, worst_cos, worst_rel)
    frac, emax, emean, lv, cos_all, worst_cos, worst_rel = out["bf16"]
    assert frac > 0.6 and emax < BF16_MAX_TRAIN and emean < BF16_MEAN_TRAIN, (frac, emax, emean)
    assert abs(lv - loss_ref) <= 3e-2 * max(1.0, abs(loss_ref)), (lv, loss_ref)
    # bf16 gradients: this loss is discontinuous in roundings (arg-max bins, class targets built from sample 0's decoded
    # angle -- quirk Q4, max-pool rows), so bf16 arithmetic alone moves its gradient: the oracle's own rounding model
    # sits at cosine ~0.82 from the fp32 oracle here (7 % of the bins flip).  The engine must be no further away than
    # that model (measured 0.84 vs 0.82); what the distance means for training is tests/test_gpu_convergence.py.
    _note("c2size_train_bf16_rounding_model", grad_cos_all_vs_fp32_oracle=cos_sim)
    assert cos_all > 0.75 and cos_all >= cos_sim - 0.05 and worst_cos > 0.6, (cos_all, cos_sim, worst_cos)


def test_c3_training_forward_and_loss_vs_oracle():
    """BASELINE configs[2] at full size: B=4096, N=200, training-mode forward (batch-statistics BN, dropout masks
    injected) and the full loss with its [B,B] couplings, against the fp32 oracle run without autograd."""
    arch, params, state, batch, masks = _case(4096, 200, 320)
    nb = arch.num_bins
    ep_ref, loss_ref = _oracle_forward(arch, params, state, batch, True, masks, dtype=torch.float32)
    dev, dm = _dev(batch), _dev(masks)
    for prec, (fmin, tmax, tmean, ltol) in (("fp32", (0.97, 2 * TOL_FP32_TRAIN, 1e-4, 1e-4)),
                                            ("bf16x6", (0.97, 2 * TOL_FP32_TRAIN, 1e-4, 1e-4)),
                                            ("bf16x3", (0.95, X3_TRAIN_MAX, 5e-4, 1e-3)),
                                            ("bf16", (0.6, BF16_MAX_TRAIN, BF16_MEAN_TRAIN, 3e-2))):
        e = _engine(arch, params, state, prec)
        ep = e.forward(dev["pcs1"], dev["pcs2"], True, 0.5, dm)
        lv = float(e.loss(dev, ep)[0].cpu())
        frac, emax, emean = _errors(_host(ep), ep_ref, nb)
        _note(f"c3_train_forward_{prec}", stable=frac, out_max_abs=emax, out_mean_abs=emean, loss=lv, loss_ref=loss_ref)
        assert frac > fmin and emax <= tmax and emean <= tmean, (prec, frac, emax, emean)
        assert abs(lv - loss_ref) <= ltol * max(1.0, abs(loss_ref)), (prec, lv, loss_ref)
        del e
        torch.cuda.empty_cache()


@pytest.mark.parametrize("N", [512, 1024])
def test_c4_c5_cloud_sizes_vs_oracle(N):
    """The cloud sizes of BASELINE configs[3] / [4] (N=512: three 176-point work items per cloud in training, two
    256-point items in inference; N=1024: five / four) end to end at B=64 against the fp64 oracle: eval forward and a
    training step with gradients, both precisions."""
    B = 64
    arch, params, state, batch, masks = _case(B, N, 330 + N)
    nb = arch.num_bins
    ref_eval, _ = _oracle_forward(arch, params, state, batch, False)
    loss_ref, ep_ref, grads_ref, _ = TR.loss_and_grads(batch, arch, params, state, 0.5, masks)
    dev, dm = _dev(batch), _dev(masks)
    for prec in ("fp32", "bf16x6", "bf16x3", "bf16"):
        e = _engine(arch, params, state, prec)
        frac, emax, emean = _errors(_host(e.forward(dev["pcs1"], dev["pcs2"], False)), ref_eval, nb,
                                    margin=1e-3 if prec != "bf16" else None)
        ep = e.forward(dev["pcs1"], dev["pcs2"], True, 0.5, dm)
        loss = e.backward(dev["pcs1"], dev["pcs2"], dev, ep)
        lv = float(loss[0].cpu())
        tfrac, tmax, tmean = _errors(_host(ep), ep_ref, nb)
        rows, cos_all = _grad_report(e.get_grads(), grads_ref)
        worst_cos = min(r[1] for r in rows)
        _note(f"N{N}_{prec}", eval_stable=frac, eval_max_abs=emax, eval_mean_abs=emean, train_stable=tfrac,
              train_max_abs=tmax, loss=lv, loss_ref=loss_ref, grad_cos_all=cos_all, grad_worst_tensor_cos=worst_cos)
        if prec in PARITY:
            assert frac > 0.9 and emax <= TOL_FP32, (N, frac, emax)
            assert tfrac > 0.9 and tmax <= 2 * TOL_FP32_TRAIN, (N, tfrac, tmax)
            assert abs(lv - loss_ref) <= 1e-4 * max(1.0, abs(loss_ref)), (lv, loss_ref)
            assert cos_all > 0.999 and worst_cos > 0.99, (cos_all, worst_cos)
        elif prec == "bf16x3":
            assert frac > 0.9 and emax <= TOL_FP32, (N, frac, emax)
            assert tfrac > 0.9 and tmax <= X3_TRAIN_MAX, (N, tfrac, tmax)
            assert abs(lv - loss_ref) <= 1e-3 * max(1.0, abs(loss_ref)), (lv, loss_ref)
            assert cos_all > 0.99, cos_all
        else:
            assert frac > 0.5 and emax < BF16_MAX and emean < BF16_MEAN, (N, frac, emax, emean)
            assert tfrac > 0.5 and tmax < BF16_MAX_TRAIN, (N, tfrac, tmax)
            assert abs(lv - loss_ref) <= 5e-2 * max(1.0, abs(loss_ref)), (lv, loss_ref)
            assert cos_all > 0.7, cos_all


def _default_arch():
    """configs/default.json:8-22 of the reference"""
    return A.Arch(num_bins=36, s1_conv=(128, 128, 256), s1_fc=(512, 256), s1_keep=0.7, s2_conv=(64, 64, 64, 128, 1024),
                  s2_fc=(512, 256), s2_keep=0.7, emb_conv=(64, 64, 64, 128, 1024), head_fc=(512, 256), head_keep=0.7,
                  angle_factor=1.0, early_stage_factor=0.1, accept_inverted_angle=False)


def test_default_architecture_all_modes_vs_oracle():
    """The reference's own default config (five-layer conv stacks, models/tp8.py:49-59 builds any depth): eval forward
    against the fp64 oracle and a training step with gradients against fp64 autograd, B=48, N=256."""
    from alignnet_b200 import synth
    B, N = 48, 256
    arch = _default_arch()
    nb = arch.num_bins
    params, state = A.randomize_for_test(arch, A.init_params(arch, 400), A.init_state(arch), 401)
    batch = synth.make_batch_fast(B, N, seed=402)
    rng = np.random.default_rng(403)
    masks = {k: (rng.uniform(size=(B, 256)) < 0.7).astype(np.float32) for k in MASK_KEYS}
    ref_eval, _ = _oracle_forward(arch, params, state, batch, False)
    loss_ref, ep_ref, grads_ref, _ = TR.loss_and_grads(batch, arch, params, state, 0.5, masks)
    # yardstick for the bf16 mode's gradient, as in the B=1024 test: the oracle with bf16 rounding at the same points
    TR.SIM_BF16 = True
    try:
        _, _, grads_sim, _ = TR.loss_and_grads(batch, arch, params, state, 0.5, masks)
    finally:
        TR.SIM_BF16 = False
    _, cos_sim = _grad_report(grads_sim, grads_ref)
    _note("default_arch_bf16_rounding_model", grad_cos_all_vs_fp64_oracle=cos_sim)
    dev, dm = _dev(batch), _dev(masks)
    for prec in ("fp32", "bf16x3", "bf16x6", "bf16"):
        e = _engine(arch, params, state, prec)
        frac, emax, emean = _errors(_host(e.forward(dev["pcs1"], dev["pcs2"], False)), ref_eval, nb,
                                    margin=1e-3 if prec != "bf16" else None)
        ep = e.forward(dev["pcs1"], dev["pcs2"], True, 0.5, dm)
        loss = e.backward(dev["pcs1"], dev["pcs2"], dev, ep)
        lv = float(loss[0].cpu())
        tfrac, tmax, tmean = _errors(_host(ep), ep_ref, nb)
        rows, cos_all = _grad_report(e.get_grads(), grads_ref)
        worst_cos = min(r[1] for r in rows)
        _note(f"default_arch_{prec}", eval_stable=frac, eval_max_abs=emax, eval_mean_abs=emean, train_stable=tfrac,
              train_max_abs=tmax, loss=lv, loss_ref=loss_ref, grad_cos_all=cos_all, grad_worst_tensor_cos=worst_cos)
        if prec in PARITY:
            assert frac > 0.9 and emax <= TOL_FP32, (prec, frac, emax)
            assert tfrac > 0.9 and tmax <= 2 * TOL_FP32_TRAIN, (prec, tfrac, tmax)
            assert abs(lv - loss_ref) <= 1e-4 * max(1.0, abs(loss_ref)), (prec, lv, loss_ref)
            assert cos_all > 0.999 and worst_cos > 0.99, (prec, cos_all, worst_cos)
        elif prec == "bf16x3":
            assert frac > 0.9 and emax <= TOL_FP32, (prec, frac, emax)
            assert tfrac > 0.9 and tmax <= X3_TRAIN_MAX, (prec, tfrac, tmax)
            assert abs(lv - loss_ref) <= 1e-3 * max(1.0, abs(loss_ref)), (prec, lv, loss_ref)
            assert cos_all > 0.99, cos_all
        else:
            # (B=48: batch-statistics BN over 48 samples in the FC layers, 8 % of the stage-2 bins flip; two runs of this
            # mode differ -- its under-filled FC GEMMs slice K and meet in fp32 reductions: measured max 0.66 / 0.89,
            # gradient cosine 0.67 / 0.61 against the rounding model's 0.70)
            assert frac > 0.5 and emax < BF16_MAX and emean < BF16_MEAN, (frac, emax, emean)
            assert tfrac > 0.5 and tmax < 2 * BF16_MAX_TRAIN and tmean < BF16_MEAN_TRAIN, (tfrac, tmax, tmean)
            assert abs(lv - loss_ref) <= 5e-2 * max(1.0, abs(loss_ref)), (lv, loss_ref)
            assert cos_all > 0.5 and cos_all >= cos_sim - 0.15, (cos_all, cos_sim)
        del e
        torch.cuda.empty_cache()
