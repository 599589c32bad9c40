"""The bf16 tensor-core conv stack (forward + backward kernels) in isolation against fp64 autograd
of the same stack with the same rounding points and the same upstream gradient.

Isolating one stack removes the model's discontinuities (arg-max bin of the canonicalisation, loss
class targets) that make end-to-end gradient comparisons in bf16 inconclusive; what remains
(max-pool arg rows, ReLU masks) perturbs the gradient only at the rounding level.
Stated bound: relative L2 error <= 2e-2 per gradient tensor, <= 1e-2 on the pooled feature."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import arch as A, torch_ref as TR
from helpers import engine_arch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64).ravel() - np.asarray(b, np.float64).ravel()) /
                 max(np.linalg.norm(np.asarray(b, np.float64).ravel()), 1e-30))


@pytest.mark.parametrize("stage,branch,B,N,rotate", [(0, 0, 16, 200, False), (1, 1, 8, 64, False), (2, 0, 12, 200, True),
                                                     (2, 1, 3, 450, True), (1, 0, 150, 24, False),
                                                     (2, 0, 4, 512, True), (0, 1, 3, 1024, False)])   # c4 / c5 cloud sizes
def test_conv_stack_fwd_bwd(stage, branch, B, N, rotate):
    import __graft_entry__ as ge
    ge.build()
    from alignnet_b200 import _lib, engine, synth
    arch = A.Arch()
    params, state = A.randomize_for_test(arch, A.init_params(arch, 60 + stage), A.init_state(arch), 61)
    e = engine.Engine(engine_arch(arch), "cuda:0", "bf16")
    e.set_params(params)
    rng = np.random.default_rng(stage * 10 + branch)
    pcs = synth.make_batch_fast(B, N, seed=70 + stage)["pcs1"]
    center = pcs.mean(axis=1) + rng.normal(0, 0.2, (B, 3)).astype(np.float32)
    angle = rng.uniform(-3, 3, B).astype(np.float32) if rotate else None
    key = ("s1_conv", "s2_conv", "emb_conv")[stage]
    specs = A.stage_specs(arch)[key]
    C3 = specs[-1].cout
    dG = (rng.normal(0, 1, (B, C3)) * (rng.uniform(size=(B, C3)) < 0.8)).astype(np.float32)

    # ---- reference: fp64 autograd through the rounding-model conv stack ----
    TR.SIM_BF16 = True
    try:
        tp = TR.to_torch(params, requires_grad=True)
        ts = TR.to_torch(state)
        tc = torch.tensor(center, dtype=torch.float64, requires_grad=True)
        x = torch.tensor(pcs, dtype=torch.float64) - tc[:, None, :]
        if rotate:
            ta = torch.tensor(angle, dtype=torch.float64, requires_grad=True)
            x = TR._rot_z_rows(x, ta)
        g_ref = TR._conv_stack(x, specs, branch, tp, ts, {}, True, 0.5)
        (g_ref * torch.tensor(dG, dtype=torch.float64)).sum().backward()
    finally:
        TR.SIM_BF16 = False

    # ---- engine ----
    lib = _lib.load()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    d_pcs, d_center, d_dG = dev(pcs), dev(center), dev(dG)
    d_angle = dev(angle) if rotate else None
    g_out = torch.empty(B, C3, device="cuda")
    grads = torch.empty(e.params.numel(), device="cuda")
    dcenter = torch.empty(B, 3, device="cuda")
    dangle = torch.empty(B, device="cuda")
    flags = _lib.TRAINING | _lib.PRECISION_BF16
    ws = torch.empty(e.workspace_bytes(B, N, flags) + 256, dtype=torch.uint8, device="cuda")
    _lib.check(lib.an3d_selftest_conv_stack(e.ctx, e.params.data_ptr(), e.bn_state.data_ptr(), stage, branch,
                                            d_pcs.data_ptr(), d_center.data_ptr(),
                                            d_angle.data_ptr() if rotate else None, B, N, d_dG.data_ptr(),
                                            g_out.data_ptr(), grads.data_ptr(), dcenter.data_ptr(), dangle.data_ptr(),
                                            ws.data_ptr(), ws.numel(), None), "an3d_selftest_conv_stack")
    torch.cuda.synchronize()
    assert _rel(g_out.cpu().numpy(), g_ref.detach().numpy()) < 1e-2
    got = e._unflatten(e.params_layout, grads.cpu().numpy())
    report = []
    for i, s in enumerate(specs):
        wn, bn = A.weight_names(s), A.bn_names(s, branch)
        for name in (wn["weights"], bn["gamma"], bn["beta"]):
            report.append((_rel(got[name], tp[name].grad.numpy()), name))
    report.append((_rel(dcenter.cpu().numpy(), tc.grad.numpy()), "dcenter"))
    if rotate:
        report.append((_rel(dangle.cpu().numpy(), ta.grad.numpy()), "dangle"))
    print(sorted(report, reverse=True))
    for err, name in report:
        assert err < 2e-2, sorted(report, reverse=True)
