"""Pins the tcgen05 descriptor conventions of csrc/umma.cuh on hardware (K-major and MN-major
operands in the un-swizzled plane layout) against a float64 matmul of the same bf16 values."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize("n,k", [(16, 16), (128, 64), (208, 128), (256, 32), (112, 128)])
def test_umma_selftest(a_mn, b_mn, n, k):
    import __graft_entry__ as ge
    ge.build()
    from alignnet_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(n * 1000 + k + a_mn * 7 + b_mn * 3)
    A = torch.randn(128, k, generator=g).to(torch.bfloat16)
    B = torch.randn(n, k, generator=g).to(torch.bfloat16)
    ref = A.double() @ B.double().T
    a_dev = (A.T.contiguous() if a_mn else A.contiguous()).cuda()
    b_dev = (B.T.contiguous() if b_mn else B.contiguous()).cuda()
    d = torch.full((128, n), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(lib.an3d_selftest_umma(a_dev.data_ptr(), b_dev.data_ptr(), d.data_ptr(), n, k, a_mn, b_mn, None),
               "an3d_selftest_umma")
    torch.cuda.synchronize()
    err = (d.cpu().double() - ref).abs().max().item()
    assert err < 1e-3 * max(1.0, ref.abs().max().item()), err
