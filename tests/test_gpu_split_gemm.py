"""The split-operand tcgen05 GEMM of the materialised path in isolation (through the C ABI) against a float64 product of
the SAME fp32 operands: forward form with the BN + ReLU + dropout prologue and bias, wgrad form (both operands MN-major,
K = rows, accumulated into C), dgrad form; ragged shapes (3-, 15-, 103-wide outputs, row counts off the 128-row tile);
the long reductions of the conv layers' wgrad (hundreds of thousands of rows).

Error model (csrc/gemm_tc.cuh): with operands split into n bf16 images the dropped products are below
2^-9 (n = 1), 2^-18 (n = 2), 2^-27 (n = 3) of |a||b|; the fp32 accumulation of the tensor core adds ~2^-24 per step.
Bounds are relative to max(|A| |B|), the natural scale of the rounding error of a dot product."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# relative to (|A| @ |B|).max(): measured values are printed with -s
BOUND = {1: 6e-3, 2: 1.2e-5, 3: 6e-7}


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from alignnet_b200 import _lib
    return _lib.load()


def run(lib, A, a_mn, Bm, b_mn, M, N, K, nsplit, bias=None, scale=None, shift=None, mask=None, mask_scale=1.0, accumulate=0,
        c_init=None, ldc=None):
    from alignnet_b200 import _lib
    dev = lambda t: None if t is None else t.float().contiguous().cuda()   # noqa: E731
    a, b, bi, sc, sh, mk = dev(A), dev(Bm), dev(bias), dev(scale), dev(shift), dev(mask)
    ldc = ldc or N
    c = torch.full((M, ldc), float("nan"), device="cuda") if c_init is None else c_init.float().cuda()
    ptr = lambda t: None if t is None else t.data_ptr()   # noqa: E731
    _lib.check(lib.an3d_selftest_split_gemm(ptr(a), a.shape[1], a_mn, ptr(b), b.shape[1], b_mn, ptr(c), ldc, M, N, K, ptr(bi),
                                            ptr(sc), ptr(sh), ptr(mk), mask_scale, nsplit, accumulate, None),
               "an3d_selftest_split_gemm")
    torch.cuda.synchronize()
    return c.cpu().double()


@pytest.mark.parametrize("nsplit", [1, 2, 3])
@pytest.mark.parametrize("M,K,N", [(4096, 256, 512), (200, 512, 256), (48, 2048, 512), (1024, 256, 103), (130, 64, 15),
                                   (64, 8, 16), (25600, 128, 1024), (1000, 24, 40), (19001, 64, 200), (40000, 128, 256)])
def test_forward_form(lib, M, K, N, nsplit):
    g = torch.Generator().manual_seed(M + K + N)
    X = torch.randn(M, K, generator=g)
    W = torch.randn(K, N, generator=g) / np.sqrt(K)
    bias = torch.randn(N, generator=g)
    scale, shift = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.3
    mask = (torch.rand(M, K, generator=g) < 0.7).float()
    # the prologue is fp32 arithmetic in the pack kernel: evaluate it the same way, then go to float64
    act = (torch.relu(torch.addcmul(shift, X, scale)) * (mask * np.float32(1 / 0.7))).double()
    ref = act @ W.double() + bias.double()
    # ldc > N: the output is a column block of a wider matrix (the head input [B, 2C]); the rest must stay untouched
    ldc = N + 4 if N % 4 == 0 else N
    c = run(lib, X, 0, W, 1, M, N, K, nsplit, bias, scale, shift, mask, 1 / 0.7, ldc=ldc)
    unit = (act.abs() @ W.double().abs()).max().item()
    err = (c[:, :N] - ref).abs().max().item() / unit
    print(f"forward {M}x{K}x{N} nsplit={nsplit}: err/unit = {err:.3e}")
    assert err < BOUND[nsplit]
    if ldc > N:
        assert torch.isnan(c[:, N:]).all()


@pytest.mark.parametrize("nsplit", [1, 2, 3])
@pytest.mark.parametrize("R,cin,cout", [(4096, 512, 256), (200, 256, 103), (48, 2048, 512), (130, 64, 15), (204800, 128, 256),
                                        (819200, 64, 128)])
def test_wgrad_and_dgrad_forms(lib, R, cin, cout, nsplit):
    g = torch.Generator().manual_seed(R + cin + cout)
    X = torch.randn(R, cin, generator=g)
    dZ = torch.randn(R, cout, generator=g)
    W = torch.randn(cin, cout, generator=g) / np.sqrt(cin)
    scale, shift = torch.rand(cin, generator=g) + 0.5, torch.randn(cin, generator=g) * 0.3
    act = torch.relu(torch.addcmul(shift, X, scale)).double()
    # wgrad: [cin, cout] += act^T dZ, accumulated onto what C holds
    prev = torch.randn(cin, cout, generator=g)
    ref_w = prev.double() + act.T @ dZ.double()
    c = run(lib, X, 1, dZ, 1, cin, cout, R, nsplit, None, scale, shift, None, 1.0, accumulate=1, c_init=prev)
    # a sum of R products with random signs: the error scale is sqrt(R) |a||b|, not R |a||b|
    unit = float(np.sqrt(R)) * 3.0
    err = (c - ref_w).abs().max().item() / unit
    print(f"wgrad R={R} {cin}x{cout} nsplit={nsplit}: err/unit = {err:.3e}  (max |ref| {ref_w.abs().max().item():.1f})")
    # (long reductions: the fp32 accumulation of ~10^5 partial products -- TMEM tile, then the K slices' reductions -- is the
    # floor: 1.4e-5 measured at R = 204800 with three images, where the split itself contributes 1e-7)
    assert err < max(BOUND[nsplit] * 4, 4e-5)
    if R > 100000:
        return
    # dgrad: [R, cin] = dZ W^T
    ref_d = dZ.double() @ W.double().T
    c = run(lib, dZ, 0, W, 0, R, cin, cout, nsplit)
    unit = (dZ.double().abs() @ W.double().abs().T).max().item()
    err = (c - ref_d).abs().max().item() / unit
    print(f"dgrad R={R} {cin}x{cout} nsplit={nsplit}: err/unit = {err:.3e}")
    assert err < BOUND[nsplit]
