"""Debug helper: per-tensor gradient error of the bf16 backward vs the rounding-model oracle."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import arch as A, torch_ref as TR
from helpers import MASK_KEYS, engine_arch
from alignnet_b200 import engine, synth
import __graft_entry__ as ge
ge.build()
B, N = int(sys.argv[1]), int(sys.argv[2])
arch = A.Arch()
params, state = A.randomize_for_test(arch, A.init_params(arch, 50), A.init_state(arch), 51)
# pin the arg-max bins (a flipped bin re-canonicalises the cloud: a discontinuity that swamps everything else)
params["siamese/transformer2/mlp/fc3/biases"][3 + 7] += 8.0
params["fc3/biases"][3 + 11] += 8.0
batch = synth.make_batch_fast(B, N, seed=52 + B)
rng = np.random.default_rng(3)
masks = {k: (rng.uniform(size=(B, 256)) < 0.7).astype(np.float32) for k in MASK_KEYS}
TR.SIM_BF16 = True
loss_ref, ep_ref, grads_ref, _ = TR.loss_and_grads(batch, arch, params, state, 0.5, masks)
TR.SIM_BF16 = False
dev = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in batch.items()}
dm = {k: torch.from_numpy(v).cuda() for k, v in masks.items()}
for prec in ("fp32", "bf16"):
    e = engine.Engine(engine_arch(arch), "cuda:0", prec); e.set_params(params); e.set_state(state)
    ep = e.forward(dev["pcs1"], dev["pcs2"], True, 0.5, dm)
    loss = e.backward(dev["pcs1"], dev["pcs2"], dev, ep)
    torch.cuda.synchronize()
    g = e.get_grads()
    print(f"== {prec}: loss {float(loss[0].cpu()):.6f} ref {loss_ref:.6f}")
    for n, ref in grads_ref.items():
        rn = np.linalg.norm(ref)
        if rn < 1e-9: continue
        err = np.linalg.norm(g[n].reshape(ref.shape).astype(np.float64) - ref) / rn
        cos = float((g[n].reshape(ref.shape).astype(np.float64) * ref).sum() / (np.linalg.norm(g[n]) * rn + 1e-30))
        flag = "  <<<" if err > 0.06 else ""
        print(f"{err:9.4f} cos {cos:7.4f} |ref| {rn:10.3e} |got| {np.linalg.norm(g[n]):10.3e}  {n}{flag}")
