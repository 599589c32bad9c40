import sys; sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests")
import numpy as np, torch
import __graft_entry__ as ge; ge.build()
from oracle import arch as A
from alignnet_b200 import synth
from test_gpu_bf16 import make_engine, to_dev
arch=A.Arch(); params=A.init_params(arch,3)
batch=to_dev(synth.make_batch_fast(64,200,seed=5))
for trial in range(3):
    es=[make_engine(arch,params,A.init_state(arch)) for _ in range(4)]
    la=float(es[0].train_step(batch,lr=0.002,bn_decay=0.5,seed=1)[0].cpu())
    lc=float(es[1].train_step(batch,lr=0.002,bn_decay=0.5,seed=1)[0].cpu())
    lb=float(es[2].train_step_graph(batch,lr=0.002,bn_decay=0.5)[0].cpu())
    ld=float(es[3].train_step(batch,lr=0.002,bn_decay=0.5,seed=12345)[0].cpu())
    print(trial, "eager",la,lc,"graph",lb,"other seed",ld)
