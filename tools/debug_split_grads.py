"""Per-tensor gradient differences between the fp32 engine and a split-operand mode on a golden case."""
import sys, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
ge.build()
from helpers import golden_case, engine_arch
from alignnet_b200 import engine

name = sys.argv[1] if len(sys.argv) > 1 else "shipped_B32_N200"
modes = sys.argv[2:] or ["bf16x6"]
g, arch, params, state, batch, masks = golden_case(name)
dev = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in batch.items()}
dm = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in masks.items()}
res = {}
for prec in ["fp32"] + modes:
    e = engine.Engine(engine_arch(arch), "cuda:0", prec)
    e.set_params(params); e.set_state(state)
    ep = e.forward(dev["pcs1"], dev["pcs2"], True, 0.5, dm)
    loss = e.backward(dev["pcs1"], dev["pcs2"], dev, ep)
    torch.cuda.synchronize()
    res[prec] = ({k: v.cpu().numpy().copy() for k, v in ep.items()}, float(loss[0].cpu()), e.get_grads())
ref = res["fp32"]
for prec in modes:
    ep, lv, gr = res[prec]
    print(prec, "loss", lv, "fp32", ref[1])
    for k in ep:
        print("  out", k, float(np.abs(ep[k] - ref[0][k]).max()))
    rows = []
    for n, v in ref[2].items():
        d = gr[n] - v
        rows.append((float(np.linalg.norm(d) / (np.linalg.norm(v) + 1e-30)), float(np.linalg.norm(v)), float(np.linalg.norm(gr[n])), n))
    rows.sort(reverse=True)
    for r in rows[:40]:
        print("  %.3e  |ref| %.4e  |got| %.4e  %s" % r)
    key = "grad/siamese/transformer1/embedding/conv1/weights"
    if key in g.files:
        print("golden norm", float(np.linalg.norm(g[key])))
