"""The tcgen05 FC GEMM of the bf16 mode in isolation (through the C ABI), against a float64 product of
the same bf16-rounded operands: forward form with the BN+ReLU+dropout prologue and the fused column
statistics, wgrad form (both operands MN-major, split-K reductions), dgrad form, ragged / unaligned
shapes (the 103- and 3-wide output layers, batches that are not multiples of the 128-row tile)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from alignnet_b200 import _lib
    return _lib.load()


def r16(t):
    return t.to(torch.bfloat16).to(torch.float64)


def run(lib, A, a_mn, Bm, b_mn, M, N, K, bias=None, scale=None, shift=None, mask=None, mask_scale=1.0, ksplit=1,
        accumulate=0, stats=False, c_init=None):
    from alignnet_b200 import _lib
    dev = lambda t: None if t is None else t.float().contiguous().cuda()   # noqa: E731
    a, b, bi, sc, sh, mk = dev(A), dev(Bm), dev(bias), dev(scale), dev(shift), dev(mask)
    c = torch.zeros(M, N, device="cuda") if c_init is None else c_init.float().cuda()
    if not (ksplit > 1 or accumulate):
        c.fill_(float("nan"))
    ssum = torch.zeros(N, dtype=torch.float64, device="cuda") if stats else None
    ssq = torch.zeros(N, dtype=torch.float64, device="cuda") if stats else None
    ptr = lambda t: None if t is None else t.data_ptr()   # noqa: E731
    _lib.check(lib.an3d_selftest_fc_gemm(ptr(a), a.shape[1], a_mn, ptr(b), b.shape[1], b_mn, ptr(c), N, M, N, K, ptr(bi),
                                         ptr(sc), ptr(sh), ptr(mk), mask_scale, ksplit, accumulate, ptr(ssum), ptr(ssq),
                                         None), "an3d_selftest_fc_gemm")
    torch.cuda.synchronize()
    return c.cpu().double(), (None if not stats else (ssum.cpu(), ssq.cpu()))


@pytest.mark.parametrize("M,K,N", [(4096, 256, 512), (200, 512, 256), (48, 2048, 512), (1024, 256, 103), (130, 256, 3),
                                   (32, 512, 512)])
@pytest.mark.parametrize("training", [True, False])
def test_fc_forward_form(lib, M, K, N, training):
    g = torch.Generator().manual_seed(M + K + N)
    X = torch.randn(M, K, generator=g)
    W = torch.randn(K, N, generator=g) / np.sqrt(K)
    bias = torch.randn(N, generator=g)
    scale, shift = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.3
    mask = (torch.rand(M, K, generator=g) < 0.7).float()
    act = torch.relu(X.double() * scale.double() + shift.double()) * mask.double() * (1 / 0.7)
    ref = r16(act) @ r16(W) + bias.double()
    ksplit = 1 if training else max(1, min(K // 128, 4))
    c, st = run(lib, X, 0, W, 1, M, N, K, bias, scale, shift, mask, 1 / 0.7, ksplit=ksplit, stats=training)
    tol = 2e-3 * max(1.0, ref.abs().max().item())
    assert (c - ref).abs().max().item() < tol
    if training:
        np.testing.assert_allclose(st[0].numpy(), ref.sum(0).numpy(), atol=2e-3 * M ** 0.5 * max(1.0, ref.abs().max().item()))
        np.testing.assert_allclose(st[1].numpy(), (ref ** 2).sum(0).numpy(), rtol=2e-3, atol=1e-2)


@pytest.mark.parametrize("R,cin,cout", [(4096, 512, 256), (200, 256, 103), (48, 2048, 512), (130, 256, 3)])
def test_fc_wgrad_and_dgrad_forms(lib, R, cin, cout):
    g = torch.Generator().manual_seed(R + cin + cout)
    X = torch.randn(R, cin, generator=g)
    dZ = torch.randn(R, cout, generator=g)
    W = torch.randn(cin, cout, generator=g) / np.sqrt(cin)
    scale, shift = torch.rand(cin, generator=g) + 0.5, torch.randn(cin, generator=g) * 0.3
    act = torch.relu(X.double() * scale.double() + shift.double())
    # wgrad: [cin, cout] += act^T dZ  (A = X MN-major with the prologue, B = dZ MN-major, K = rows), accumulated
    prev = torch.randn(cin, cout, generator=g)
    ref_w = prev.double() + r16(act).T @ r16(dZ)
    ks = max(1, min((R + 511) // 512, 8))
    c, _ = run(lib, X, 1, dZ, 1, cin, cout, R, None, scale, shift, None, 1.0, ksplit=ks, accumulate=1, c_init=prev)
    assert (c - ref_w).abs().max().item() < 2e-3 * max(1.0, ref_w.abs().max().item())
    # dgrad: [R, cin] = dZ W^T  (A = dZ K-major, B = W K-major over cout)
    ref_d = r16(dZ) @ r16(W).T
    c, _ = run(lib, dZ, 0, W, 0, R, cin, cout)
    assert (c - ref_d).abs().max().item() < 2e-3 * max(1.0, ref_d.abs().max().item())
