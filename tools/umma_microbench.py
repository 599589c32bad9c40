"""tcgen05 MMA issue-rate micro-benchmark (run on a B200): cycles per 128 x N x 16 bf16 MMA for K-major and
MN-major operands in the un-swizzled plane layout.  Output is committed under profiles/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from alignnet_b200 import _lib
lib = _lib.load()
out = torch.zeros(148, device="cuda")
print("N    K  a_mn b_mn  ctas  cycles/MMA   (nominal = 128*N*16 / 3868 MAC/clk)")
for ctas in (1, 148):
    for n, k in ((208, 128), (112, 128), (64, 128), (256, 128), (128, 208), (80, 208), (64, 208)):
        for a_mn, b_mn in ((0, 0), (1, 1), (0, 1), (1, 0)):
            _lib.check(lib.an3d_bench_umma(n, k, a_mn, b_mn, 4096, ctas, out.data_ptr(), None), "bench")
            torch.cuda.synchronize()
            v = out[:ctas].cpu()
            print(f"{n:4d} {k:4d}  {a_mn}    {b_mn}    {ctas:4d}  {v.mean():8.1f}   nominal {128*n*16/3868:6.1f}")
