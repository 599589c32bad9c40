"""Target for compute-sanitizer: the entry points added in the third session of round 2 -- an3d_momentum_step (ragged
count, vector and tail paths), one training step with the momentum optimiser on the fused bf16 path, and the device ICP as
the `icp` special mode drives it (ragged clouds, an empty cloud, chunked launches).
    compute-sanitizer --tool memcheck python tools/sanitize_optim_icp.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge

ge.build()
from alignnet_b200 import _lib, engine, icp, synth

lib = _lib.load()
for n in (1, 3, 4, 1021, 4096):                                   # tails of 1-3 elements next to the float4 path
    p, g, a = (torch.randn(n + 4, device="cuda")[:n] for _ in range(3))
    buf = [torch.zeros(n + 8, device="cuda") for _ in range(3)]
    for b, src in zip(buf, (p, g, a)):
        b[:n] = src
    _lib.check(lib.an3d_momentum_step(buf[0].data_ptr(), buf[1].data_ptr(), buf[2].data_ptr(), n, 0.1, 0.9, 0.5, None), "momentum")
    torch.cuda.synchronize()
    ref_a = 0.9 * a + 0.5 * g
    assert torch.allclose(buf[2][:n], ref_a, atol=1e-6) and torch.allclose(buf[0][:n], p - 0.1 * ref_a, atol=1e-6)
    assert not buf[0][n:].any() and not buf[2][n:].any()          # nothing written past `count`
print("momentum kernel ok")
e = engine.Engine(engine.shipped_arch(), "cuda:0", "bf16", seed=1)
e.set_optimizer("momentum", momentum=0.9)
batch = {k: torch.from_numpy(v).cuda() for k, v in synth.make_batch_fast(8, 40, seed=3).items()}
for _ in range(2):
    loss = e.train_step(batch, lr=1e-4, bn_decay=0.5)
torch.cuda.synchronize()
print("momentum training step ok", float(loss[0].cpu()))
rng = np.random.default_rng(0)
src = [rng.normal(size=(n, 4)).astype(np.float32) for n in (1, 37, 300, 0, 1100)]
tgt = [s[:, :3] + 0.01 for s in src[:3]] + [rng.normal(size=(5, 3)).astype(np.float32), src[4][:900, :3] + 0.02]
T, stats = icp.refine(src, tgt, np.stack([np.eye(4)] * 5), radius=0.1, its=5)
assert np.isfinite(T).all()
print("icp ok", stats[:, 0])
