"""bf16 tensor-core (tcgen05) path against the fp64 oracle.

The fast mode rounds the 64- and 128-wide activations and the conv weights to bf16 (8-bit
mantissa, fp32 accumulation in TMEM), so it is NOT held to the 1e-4 parity bound of the fp32
mode; the stated bound here is 1.2e-1 max-abs / 1.5e-2 mean-abs on the 8 end_points (units: metres
for centres / translations, logit units for the angle heads), with samples whose stage-2 arg-max
is ambiguous at that precision excluded from everything downstream of the canonicalisation."""
import numpy as np
import pytest
import torch

from oracle import arch as A, np_forward as NF, torch_ref as TR
from helpers import MASK_KEYS, OUTPUT_KEYS, engine_arch, golden_case, top2_margin

pytestmark = pytest.mark.gpu

MAX_ABS = 1.2e-1
MEAN_ABS = 1.5e-2
# training mode normalises with batch statistics (over as few as 4 samples in the FC layers of these
# small test batches), which amplifies the bf16 rounding of the pooled features
MAX_ABS_TRAIN = 8e-1
MEAN_ABS_TRAIN = 2.5e-1


@pytest.fixture(scope="module", autouse=True)
def _build():
    import __graft_entry__ as ge
    ge.build()


def make_engine(arch, params, state, precision="bf16"):
    from alignnet_b200 import engine
    e = engine.Engine(engine_arch(arch), "cuda:0", precision)
    e.set_params(params)
    e.set_state(state)
    return e


def to_dev(d):
    return {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in d.items()}


def compare(ep, ref, arch, max_abs=MAX_ABS, mean_abs=MEAN_ABS):
    """Worst (max-abs, mean-abs) over the outputs.  Outputs downstream of the canonicalisation are
    compared on the samples whose stage-2 arg-max bin agrees with the oracle's (a flipped bin
    rotates the cloud by 2*pi/nb: a discontinuity of the reference function, not an error)."""
    nb = arch.num_bins
    stable = np.ones(ref["pred_translations"].shape[0], bool)
    for k in ("pred_pc1angle_logits", "pred_pc2angle_logits"):
        stable &= ep[k].cpu().numpy()[:, :nb].argmax(1) == ref[k][:, :nb].argmax(1)
    assert stable.mean() >= 0.5, f"only {stable.mean():.2f} of the samples keep their arg-max bin"
    worst_max = worst_mean = 0.0
    for k in OUTPUT_KEYS:
        got = ep[k].cpu().numpy()
        assert np.isfinite(got).all(), k
        downstream = k in ("pred_translations", "pred_remaining_angle_logits")
        d = np.abs(got - ref[k])
        if downstream:
            d = d[stable]
        worst_max, worst_mean = max(worst_max, float(d.max())), max(worst_mean, float(d.mean()))
        assert d.max() < max_abs and d.mean() < mean_abs, (k, float(d.max()), float(d.mean()))
    return worst_max, worst_mean


@pytest.mark.parametrize("name", ["shipped_B4_N16", "shipped_B32_N200"])
def test_bf16_forward_eval(name):
    g, arch, params, state, batch, masks = golden_case(name)
    e = make_engine(arch, params, state)
    dev = to_dev(batch)
    ep = e.forward(dev["pcs1"], dev["pcs2"], False)
    torch.cuda.synchronize()
    compare(ep, {k: g["eval/" + k] for k in OUTPUT_KEYS}, arch)
    st = e.get_state()
    for k, v in state.items():
        np.testing.assert_array_equal(st[k], v)


@pytest.mark.parametrize("name", ["shipped_B32_N200"])
def test_bf16_forward_train(name):
    g, arch, params, state, batch, masks = golden_case(name)
    e = make_engine(arch, params, state)
    dev, dm = to_dev(batch), to_dev(masks)
    ep = e.forward(dev["pcs1"], dev["pcs2"], True, 0.5, dm)
    torch.cuda.synchronize()
    compare(ep, {k: g["train64/" + k] for k in OUTPUT_KEYS}, arch, MAX_ABS_TRAIN, MEAN_ABS_TRAIN)
    st = e.get_state()
    for k in [k for k in g.files if k.startswith("state/") and "/conv" in k]:
        # conv-layer batch statistics (over B*N points) are stable under bf16 rounding; the FC-layer ones
        # (over 32 samples) are checked against the rounding-model oracle below
        np.testing.assert_allclose(st[k[6:]], g[k], atol=2e-2, rtol=2e-2, err_msg=k)


@pytest.mark.parametrize("B,N", [(3, 24), (2, 256), (5, 200), (2, 512), (2, 1000), (150, 40)])
def test_bf16_forward_shapes(B, N):
    """Ragged tiles: N not a multiple of 16, clouds split into several <=256-point items (cross-item
    max via atomics), more / fewer items than SMs."""
    from alignnet_b200 import synth
    arch = A.Arch()
    params, state = A.randomize_for_test(arch, A.init_params(arch, 40), A.init_state(arch), 41)
    batch = synth.make_batch_fast(B, N, seed=B * 1000 + N)
    ref, _ = NF.get_model(batch["pcs1"], batch["pcs2"], arch, params, state, False)
    e = make_engine(arch, params, state)
    dev = to_dev(batch)
    ep = e.forward(dev["pcs1"], dev["pcs2"], False)
    torch.cuda.synchronize()
    compare(ep, ref, arch)
    if B >= 4:
        ref_t, _ = NF.get_model(batch["pcs1"], batch["pcs2"], arch, params, state, True, 0.5, None)
        arch_nodrop = arch
        e2 = make_engine(arch, params, state)
        ones = {k: torch.ones(B, 256, device="cuda") for k in MASK_KEYS}
        ep_t = e2.forward(dev["pcs1"], dev["pcs2"], True, 0.5, ones)
        torch.cuda.synchronize()
        # oracle with masks=None applies no dropout; an all-ones mask still divides by keep_prob
        ref_t2, _ = NF.get_model(batch["pcs1"], batch["pcs2"], arch, params, state, True, 0.5,
                                 {k: np.ones((B, 256), np.float32) for k in MASK_KEYS})
        compare(ep_t, ref_t2, arch, MAX_ABS_TRAIN, MEAN_ABS_TRAIN)


def test_bf16_fc_tensor_core_path_matches_rounding_model(monkeypatch):
    """B >= 64 routes the hidden FC GEMMs (forward here) through the tcgen05 FC kernel (the size
    threshold that normally keeps small GEMMs on the SIMT kernel is lifted for the test)."""
    monkeypatch.setenv("AN3D_FC_TENSOR_MIN_FLOP", "0")
    from alignnet_b200 import synth
    arch = A.Arch()
    params, state = A.randomize_for_test(arch, A.init_params(arch, 80), A.init_state(arch), 81)
    B, N = 200, 40
    batch = synth.make_batch_fast(B, N, seed=82)
    for training in (False, True):
        TR.SIM_BF16 = True
        try:
            t = TR.to_torch(batch)
            ones = {k: torch.ones(B, 256, dtype=torch.float64) for k in MASK_KEYS}
            ep_ref, _ = TR.get_model(t["pcs1"], t["pcs2"], arch, TR.to_torch(params), TR.to_torch(state), training, 0.5, ones)
        finally:
            TR.SIM_BF16 = False
        e = make_engine(arch, params, state)
        dev = to_dev(batch)
        ep = e.forward(dev["pcs1"], dev["pcs2"], training, 0.5, {k: torch.ones(B, 256, device="cuda") for k in MASK_KEYS})
        torch.cuda.synchronize()
        # every FC layer (inputs and weights) is rounded to bf16 in this mode; the three chained MLPs and the
        # centre / angle cascade between the stages amplify that to a few 1e-1 on +-10 m translations
        compare(ep, {k: v.numpy() for k, v in ep_ref.items()}, arch, 5e-1, 8e-2)


@pytest.mark.parametrize("name,training", [("shipped_B32_N200", True), ("shipped_B32_N200", False)])
def test_bf16_matches_rounding_model(name, training):
    """Tight check: against the fp64 oracle with the SAME rounding points (bf16 activations / weights
    into conv layers 2 and 3) the engine agrees to 8e-2 max-abs / 1.5e-2 mean-abs, i.e. the larger
    train-mode deviations above are bf16 rounding amplified by batch-statistics BN, not a defect."""
    g, arch, params, state, batch, masks = golden_case(name)
    TR.SIM_BF16 = True
    try:
        t = TR.to_torch(batch)
        ep_ref, st_ref = TR.get_model(t["pcs1"], t["pcs2"], arch, TR.to_torch(params), TR.to_torch(state), training, 0.5,
                                      TR.to_torch(masks))
    finally:
        TR.SIM_BF16 = False
    ref = {k: v.numpy() for k, v in ep_ref.items()}
    e = make_engine(arch, params, state)
    dev, dm = to_dev(batch), to_dev(masks)
    ep = e.forward(dev["pcs1"], dev["pcs2"], training, 0.5, dm)
    torch.cuda.synchronize()
    # training mode: batch-statistics BN over only 32 rows in the FC layers amplifies the bf16 rounding of
    # their inputs (near-tie roundings differ between the fp64 model and the fp32-accumulating engine)
    compare(ep, ref, arch, *((6e-1, 1.2e-1) if training else (8e-2, 1.5e-2)))
    if training:
        st = e.get_state()
        # EMA shadows.  Conv layers see ~6e3 rows per statistic, FC layers only 32 (and every FC input is
        # bf16-rounded).  Layers downstream of the canonicalisation (final embedding, head) also see the
        # samples whose arg-max bin flipped, i.e. a different rotation: held on the mean deviation only.
        for k, v in st_ref.items():
            downstream = not ("transformer1" in k or "transformer2" in k)
            d = np.abs(st[k] - v.numpy())
            if downstream:
                assert d.mean() < 5e-2, (k, float(d.mean()), float(d.max()))
            else:
                tol = 1e-2 if "/embedding/" in k else 6e-2
                np.testing.assert_allclose(st[k], v.numpy(), atol=tol, rtol=tol, err_msg=k)


def test_bf16_other_architectures_take_the_layer_by_layer_path():
    """A conv stack that is not [64, 128, C] used to be refused in bf16 mode; it now runs layer by layer through the
    split-operand tensor-core GEMM with one bf16 image per operand (csrc/gemm_tc.cuh).  The fused kernels' diagnostic
    entry still refuses it."""
    from alignnet_b200 import _lib
    arch = A.tiny_arch()
    params, state = A.randomize_for_test(arch, A.init_params(arch, 0), A.init_state(arch), 1)
    e16 = make_engine(arch, params, state)
    e32 = make_engine(arch, params, state, "fp32")
    g = torch.Generator().manual_seed(5)
    x1, x2 = torch.randn(8, 40, 3, generator=g).cuda(), torch.randn(8, 40, 3, generator=g).cuda()
    a, b = e16.forward(x1, x2, False), e32.forward(x1, x2, False)
    torch.cuda.synchronize()
    for k in ("pred_s1_pc1centers", "pred_s2_pc2centers", "pred_pc1angle_logits"):      # upstream of the arg-max canonicalisation
        d = (a[k] - b[k]).abs().max().item()
        assert np.isfinite(a[k].cpu().numpy()).all() and d < MAX_ABS, (k, d)
    lib = _lib.load()
    z = torch.zeros(8, 3, device="cuda")
    rc = lib.an3d_selftest_conv_stack(e16.ctx, e16.params.data_ptr(), e16.bn_state.data_ptr(), 0, 0, x1.data_ptr(), z.data_ptr(), None,
                                      8, 40, z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), None, z.data_ptr(), 0, None)
    assert rc == -2, rc                                                                    # AN3D_ERR_UNSUPPORTED


def _rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64).ravel() - b.ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


@pytest.mark.parametrize("B,N", [(32, 200), (48, 450), (192, 48)])
def test_bf16_backward_vs_rounding_model_autograd(B, N):
    """End-to-end loss + parameter gradients of the fast mode against fp64 autograd through the oracle with
    the same forward rounding points.  The tight check of the tensor-core backward kernels is
    tests/test_gpu_conv_stack.py (<= 2e-2 per tensor with the model's discontinuities removed).  End to end
    the gradient of THIS loss is chaotic at bf16 resolution -- batch-statistics BN over 32 samples, the
    arg-max bin of the canonicalisation, class targets built from sample 0's decoded angle (quirk Q4): on
    the same case the fp32 engine, which matches fp64 autograd to <1e-2, sits 0.4-0.8 (relative L2) from
    this rounding-model oracle.  So the bound here is directional: cosine >= 0.8 for every tensor that
    carries at least 1e-2 of the largest gradient norm (>= 0.7 if up to 5 % of the bins flipped), and the
    loss within 3e-2 relative, on the seeded batch (of up to six) with the fewest flipped arg-max bins.  The FC GEMM
    (forward / wgrad / dgrad forms) has its own tight test in tests/test_gpu_fc_gemm.py."""
    from alignnet_b200 import synth
    arch = A.Arch()
    nb = arch.num_bins
    params, state = A.randomize_for_test(arch, A.init_params(arch, 50), A.init_state(arch), 51)
    rng = np.random.default_rng(3)
    masks = {k: (rng.uniform(size=(B, 256)) < 0.7).astype(np.float32) for k in MASK_KEYS}
    trials = []
    for trial in range(6):
        batch = synth.make_batch_fast(B, N, seed=52 + B + 1000 * trial)
        TR.SIM_BF16 = True
        try:
            loss_ref, ep_ref, grads_ref, _ = TR.loss_and_grads(batch, arch, params, state, 0.5, masks)
        finally:
            TR.SIM_BF16 = False
        e = make_engine(arch, params, state)
        dev, dm = to_dev(batch), to_dev(masks)
        ep = e.forward(dev["pcs1"], dev["pcs2"], True, 0.5, dm)
        loss = e.backward(dev["pcs1"], dev["pcs2"], dev, ep)
        torch.cuda.synchronize()
        grads = e.get_grads()
        for n in grads_ref:
            assert np.isfinite(grads[n]).all(), n
        # a flipped arg-max bin (canonicalisation angle, class targets of the stage-3 loss) is a discontinuity
        # of the reference function: compare gradients only on batches where every bin agrees with the model's
        flips = sum(int((ep[k].cpu().numpy()[:, :nb].argmax(1) != ep_ref[k][:, :nb].argmax(1)).sum())
                    for k in ("pred_pc1angle_logits", "pred_pc2angle_logits", "pred_remaining_angle_logits"))
        got_loss = float(loss[0].cpu())
        gnorm_max = max(float(np.linalg.norm(v)) for v in grads_ref.values())
        report = []
        for n, ref in grads_ref.items():
            rn = float(np.linalg.norm(ref))
            if rn < 1e-2 * gnorm_max:
                continue
            g = grads[n].reshape(ref.shape).astype(np.float64)
            report.append((float((g * ref).sum() / (np.linalg.norm(g) * rn + 1e-30)), n))
        report.sort()
        print(f"trial {trial}: {flips} flipped bins, loss {got_loss:.5f} vs {loss_ref:.5f}, lowest cosines:", report[:4])
        trials.append((flips, abs(got_loss - loss_ref) / max(1.0, abs(loss_ref)), report))
        if flips == 0:
            break
    flips, loss_err, report = min(trials, key=lambda t: t[0])
    assert flips <= max(1, (3 * B) // 20), f"every trial flipped more than 5% of the arg-max bins ({flips})"
    assert loss_err < 3e-2, loss_err
    assert report[0][0] > (0.8 if flips == 0 else 0.7), (flips, report[:8])


def test_bf16_train_step_runs_and_learns():
    """A few optimiser steps in the fast mode reduce the loss on a fixed batch."""
    from alignnet_b200 import synth
    arch = A.Arch()
    e = make_engine(arch, A.init_params(arch, 0), A.init_state(arch))
    batch = to_dev(synth.make_batch_fast(64, 200, seed=77))
    losses = []
    for i in range(60):
        losses.append(float(e.train_step(batch, lr=0.001, bn_decay=0.5, seed=i)[0].cpu()))
    assert np.isfinite(losses).all()
    # Adam's first steps overshoot on a 64-pair batch; over 60 steps the fixed batch is being fitted
    # (convergence against the fp32 mode: tests/test_gpu_convergence.py)
    assert max(losses) < 3 * losses[0] and np.mean(losses[-10:]) < 0.9 * losses[0], losses


def test_graph_replay_matches_eager_steps():
    """Engine.train_step_graph / forward_graph (CUDA-graph replay with the global step and the dropout seed in
    device memory) against the eager calls."""
    from alignnet_b200 import synth
    arch = A.Arch()
    params = A.init_params(arch, 3)
    batch = to_dev(synth.make_batch_fast(64, 200, seed=5))
    ea, eb, ec = (make_engine(arch, params, A.init_state(arch)) for _ in range(3))
    # first step: same parameters, same dropout seed (the graph path seeds with the new step count t = 1).  Two
    # eager runs of the bf16 step differ by the reordering of fp32 atomics; the graph run must sit inside that noise.
    l_a = float(ea.train_step(batch, lr=0.002, bn_decay=0.5, seed=1)[0].cpu())
    l_c = float(ec.train_step(batch, lr=0.002, bn_decay=0.5, seed=1)[0].cpu())
    l_b = float(eb.train_step_graph(batch, lr=0.002, bn_decay=0.5)[0].cpu())
    # measured: two eager runs of this step differ by up to ~1 % in the loss (fp32 atomics reorder the Gram / statistics
    # sums, bf16 roundings and arg-max bins flip downstream); the graph run has to sit inside that spread
    noise = abs(l_a - l_c)
    assert abs(l_b - 0.5 * (l_a + l_c)) <= 3 * noise + 3e-2 * abs(l_a), (l_a, l_c, l_b)
    # another seed draws other masks
    ed = make_engine(arch, params, A.init_state(arch))
    l_d = float(ed.train_step(batch, lr=0.002, bn_decay=0.5, seed=12345)[0].cpu())
    assert l_d != l_a
    losses = [l_b]
    for _ in range(4):
        losses.append(float(eb.train_step_graph(batch, lr=0.002, bn_decay=0.5)[0].cpu()))
    assert eb.step == 5 and int(eb.step_dev.cpu()) == 5 and int(eb.seed_dev.cpu()) == 5
    assert np.isfinite(losses).all() and len(set(losses)) == 5          # every replay is a new step
    # Adam with the step count read on the device == Adam with the host step count
    g = torch.randn_like(ea.grads)
    for e in (ea, ec):
        e.set_params(params); e.adam_m.zero_(); e.adam_v.zero_(); e.grads.copy_(g)
    ea.step = 6
    ea.adam_step(0.01)                           # host t = 7
    ec.step_dev.fill_(7)
    ec._adam_step_dev(0.01, 1.0)
    torch.cuda.synchronize()
    np.testing.assert_allclose(ec.params.cpu().numpy(), ea.params.cpu().numpy(), rtol=0, atol=2e-7)
    # eval forward: replay == eager call, up to the run-to-run variation of the bf16 path (split-K fp32 reductions
    # reorder, a bf16 rounding flips here and there downstream): nearly all elements agree tightly, none is far off
    ef = make_engine(arch, params, A.init_state(arch))
    for i in range(3):
        ef.forward(batch["pcs1"], batch["pcs2"], True, 0.5, None, seed=i)       # populate the BN shadows
    x = ef.forward(batch["pcs1"], batch["pcs2"], False)
    xa = {k: v.clone() for k, v in x.items()}
    for _ in range(2):
        xg = ef.forward_graph(batch["pcs1"], batch["pcs2"])
    for k in xa:
        a, b = xg[k].cpu().numpy(), xa[k].cpu().numpy()
        close = np.abs(a - b) <= 1e-3 + 1e-3 * np.abs(b)
        assert close.mean() > 0.9 and np.abs(a - b).max() <= 5e-2 * max(1.0, np.abs(b).max()), (k, close.mean())


@pytest.mark.parametrize("B,N", [(1, 1), (1, 200), (2, 15), (300, 8)])
def test_bf16_edge_shapes_run_and_stay_finite(B, N):
    """Degenerate shapes on the tensor-core path: a single pair, a single point, fewer items than SMs, more items
    than SMs with tiny clouds.  Eval against the oracle on the bf16 bound; a training step must stay finite."""
    from alignnet_b200 import synth
    arch = A.Arch()
    params, state = A.randomize_for_test(arch, A.init_params(arch, 40), A.init_state(arch), 41)
    batch = synth.make_batch_fast(B, N, seed=B * 1000 + N)
    ref, _ = NF.get_model(batch["pcs1"], batch["pcs2"], arch, params, state, False)
    e = make_engine(arch, params, state)
    dev = to_dev(batch)
    ep = e.forward(dev["pcs1"], dev["pcs2"], False)
    torch.cuda.synchronize()
    compare(ep, ref, arch)
    loss = e.train_step(dev, lr=0.001, bn_decay=0.5, seed=1)
    torch.cuda.synchronize()
    assert np.isfinite(loss.cpu().numpy()).all() and torch.isfinite(e.params).all() and torch.isfinite(e.grads).all()


@pytest.mark.parametrize("archname", ["default", "tiny"])
def test_layer_by_layer_bf16_path_equals_the_rounding_model(archname):
    """The layer-by-layer tensor-core path (csrc/gemm_tc.cuh with one bf16 image per operand: the bf16 mode of every
    architecture the fused kernels do not cover, e.g. configs/default.json's five-layer stacks) rounds exactly where the
    oracle's rounding model does (oracle.torch_ref.SIM_BF16 with SIM_BF16_MIN_DIM = 8: inputs and weights of every conv
    layer after the first and of every FC layer wider than 3 to bf16, everything else exact) and accumulates in fp32.
    Unlike the fused kernels -- which fold BN into bf16 weights and are held to the loose fast-mode bound -- it must agree
    with that model, not merely with the fp64 function:
      * inference: to fp32 accumulation noise on the narrow architecture (measured 1.3e-6), and up to the few activations
        that fp32 vs fp64 accumulation moves across a bf16 rounding boundary on the 1024-wide five-layer stacks
        (measured max 4.8e-3 / mean 3.3e-4 -- ten times closer than to the unrounded oracle);
      * training: outputs upstream of every arg-max; loss and the whole gradient on a batch where no arg-max bin differs
        (the loss builds class targets from decoded bins -- quirk Q4 -- so one flipped bin changes it at order one).
    Measured values are printed."""
    from alignnet_b200 import synth
    arch = A.default_arch() if archname == "default" else A.tiny_arch()
    B, N = (40, 96) if archname == "default" else (24, 50)
    nb = arch.num_bins
    params, state = A.randomize_for_test(arch, A.init_params(arch, 60), A.init_state(arch), 61)
    rng = np.random.default_rng(63)
    masks = {k: (rng.uniform(size=(B, arch.s1_fc[-1])) < 0.7).astype(np.float32) for k in MASK_KEYS}
    e = make_engine(arch, params, state)
    upstream = ("pred_s1_pc1centers", "pred_s1_pc2centers", "pred_s2_pc1centers", "pred_s2_pc2centers", "pred_pc1angle_logits",
                "pred_pc2angle_logits")
    checked_grad = False
    for trial in range(4):
        batch = synth.make_batch_fast(B, N, seed=62 + 100 * trial)
        TR.SIM_BF16, TR.SIM_BF16_MIN_DIM = True, 8        # (3-wide layers stay on the CUDA cores in fp32)
        try:
            p64, s64, b64 = TR.to_torch(params, torch.float64), TR.to_torch(state, torch.float64), TR.to_torch(batch, torch.float64)
            with torch.no_grad():
                ev_ref, _ = TR.get_model(b64["pcs1"], b64["pcs2"], arch, p64, s64, False, None, None)
            ev_ref = {k: v.numpy() for k, v in ev_ref.items()}
            loss_ref, tr_ref, grads_ref, _ = TR.loss_and_grads(batch, arch, params, state, 0.5, masks)
        finally:
            TR.SIM_BF16, TR.SIM_BF16_MIN_DIM = False, 0
        e.set_state(state)
        dev, dm = to_dev(batch), to_dev(masks)
        ev = {k: v.cpu().numpy().astype(np.float64) for k, v in e.forward(dev["pcs1"], dev["pcs2"], False).items()}
        ep = e.forward(dev["pcs1"], dev["pcs2"], True, 0.5, dm)
        loss = float(e.backward(dev["pcs1"], dev["pcs2"], dev, ep)[0].cpu())
        tr = {k: v.cpu().numpy().astype(np.float64) for k, v in ep.items()}
        grads = e.get_grads()

        def flips(got, ref):
            return sum(int((got[k][:, :nb].argmax(1) != ref[k][:, :nb].argmax(1)).sum())
                       for k in ("pred_pc1angle_logits", "pred_pc2angle_logits", "pred_remaining_angle_logits"))

        def worst(got, ref, keys):
            d = [np.abs(got[k] - ref[k]) for k in keys]
            return max(float(x.max()) for x in d), max(float(x.mean()) for x in d)

        ef, tf = flips(ev, ev_ref), flips(tr, tr_ref)
        emax, emean = worst(ev, ev_ref, OUTPUT_KEYS if ef == 0 else upstream)
        tmax, tmean = worst(tr, tr_ref, upstream)
        dot = n1 = n2 = 0.0
        for n, ref in grads_ref.items():
            g = grads[n].reshape(ref.shape).astype(np.float64)
            dot += float((g * ref).sum()); n1 += float((g * g).sum()); n2 += float((ref * ref).sum())
        cos = dot / np.sqrt(n1 * n2)
        print(f"{archname} trial {trial}: eval flips {ef} max {emax:.2e} mean {emean:.2e}; train flips {tf} upstream max {tmax:.2e} mean "
              f"{tmean:.2e}; loss {loss:.6f} vs {loss_ref:.6f}; gradient cosine {cos:.5f}")
        if archname == "tiny":
            assert emax < 1e-4, emax
        else:
            assert emax < 2e-2 and emean < 1e-3, (emax, emean)
        assert tmax < 1.5e-1 and tmean < 2e-2, (tmax, tmean)
        if tf == 0:
            assert abs(loss - loss_ref) <= 1e-2 * max(1.0, abs(loss_ref)), (loss, loss_ref)
            assert cos > 0.97, cos
            checked_grad = True
            break
    # (the five-layer architecture at this batch size flips 4-12 of its 120 bins in every trial: 36 narrow bins, batch
    # statistics over 40 samples; its loss / gradient comparison is the yardstick test of tests/test_zz_fullsize_oracle.py)
    assert checked_grad or archname == "default", "no batch without a flipped arg-max bin in four trials"
