"""Target for compute-sanitizer: the split-operand GEMM in its three kernel variants and one small training step of the
layer-by-layer tensor-core path (default architecture) and of the fused bf16 path.
    compute-sanitizer --tool memcheck  python tools/sanitize_split.py
    compute-sanitizer --tool racecheck python tools/sanitize_split.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge

ge.build()
from alignnet_b200 import _lib, engine, synth

lib = _lib.load()


def gemm(A, a_mn, B, b_mn, M, N, K, nsplit, accumulate=0, bias=None):
    a, b = A.float().contiguous().cuda(), B.float().contiguous().cuda()
    c = torch.zeros(M, N, device="cuda")
    bi = None if bias is None else bias.float().cuda()
    _lib.check(lib.an3d_selftest_split_gemm(a.data_ptr(), a.shape[1], a_mn, b.data_ptr(), b.shape[1], b_mn, c.data_ptr(), N, M, N, K,
                                            None if bi is None else bi.data_ptr(), None, None, None, 1.0, nsplit, accumulate, None), "gemm")
    torch.cuda.synchronize()
    return c


g = torch.Generator().manual_seed(0)
for nsplit in (1, 2, 3):
    X, W = torch.randn(19001, 64, generator=g), torch.randn(64, 200, generator=g)
    c = gemm(X, 0, W, 1, 19001, 200, 64, nsplit, bias=torch.randn(200, generator=g))          # A-stationary kernel, ragged
    X, W = torch.randn(300, 512, generator=g), torch.randn(512, 103, generator=g)
    c = gemm(X, 0, W, 1, 300, 103, 512, nsplit)                                                # generic, 3-stage ring, scalar stores
    X, dZ = torch.randn(5000, 64, generator=g), torch.randn(5000, 136, generator=g)
    c = gemm(X, 1, dZ, 1, 64, 136, 5000, nsplit, accumulate=1)                                 # wgrad form, K slices
    c = gemm(dZ, 0, torch.randn(64, 136, generator=g), 0, 5000, 64, 136, nsplit)               # dgrad form
print("gemm ok")
for arch, prec, B, N in ((engine.default_arch(), "bf16x6", 6, 70), (engine.default_arch(), "bf16", 6, 70),
                         (engine.shipped_arch(), "bf16x3", 5, 33), (engine.shipped_arch(), "bf16", 8, 40)):
    e = engine.Engine(arch, "cuda:0", prec, seed=0)
    batch = {k: torch.from_numpy(v).cuda() for k, v in synth.make_batch_fast(B, N, seed=3).items()}
    loss = e.train_step(batch, lr=1e-3, bn_decay=0.5)
    ep = e.forward(batch["pcs1"], batch["pcs2"], False)
    torch.cuda.synchronize()
    assert np.isfinite(float(loss[0].cpu())) and all(torch.isfinite(v).all() for v in ep.values()), prec
    print("step ok", prec, float(loss[0].cpu()))
