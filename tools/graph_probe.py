import sys, os, time
sys.path.insert(0, "/root/repo")
import torch
import __graft_entry__ as ge
ge.build()
from alignnet_b200 import engine, synth
dev = torch.device("cuda:0")
for name, B, train in (("c3", 4096, True), ("c2", 1024, False)):
    eng = engine.Engine(engine.shipped_arch(), "cuda:0", "bf16", seed=0)
    host = synth.make_batch_fast(B, 200, seed=1)
    batch = {k: torch.from_numpy(v).to(dev) for k, v in host.items()}
    for i in range(10):
        eng.forward(batch["pcs1"], batch["pcs2"], True, 0.5, None, seed=i)
    def step():
        if train:
            return eng.train_step(batch, lr=0.005, bn_decay=0.5)
        return eng.forward(batch["pcs1"], batch["pcs2"], False)
    for _ in range(3): step()
    torch.cuda.synchronize()
    def timeit(fn, n=10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_stream = timeit(step)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            step()
        t_graph = timeit(g.replay)
        print(name, "stream ms", round(t_stream, 3), "graph ms", round(t_graph, 3))
    except Exception as e:
        print(name, "capture failed:", repr(e)[:300], "stream ms", round(t_stream, 3))
