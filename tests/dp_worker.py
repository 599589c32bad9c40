"""One rank of the data-parallel hardware test (launched by tests/test_gpu_dp.py through torchrun, world size 2+).

Checks, on real GPUs over NCCL:
  1. the all-reduced gradient equals the mean of the shard gradients: every rank's own shard gradient is gathered and
     averaged on the host side of the check, and rank 0 additionally recomputes EVERY shard on its own GPU (the forward
     is bit-reproducible, so a shard's gradient does not depend on which GPU computed it beyond the backward's fp32
     reductions);
  2. after K optimiser steps every rank holds bit-identical parameters and Adam state -- through the CUDA-graph path
     with the collective captured inside the graph, and through the eager path;
  3. a G-rank step at global batch B equals G single-GPU replicas at batch B / G with averaged gradients (per-shard
     BN statistics and loss coupling): rank 0 replays the K steps alone with the averaged shard gradients.
Prints one JSON line on rank 0; exit code != 0 on any failed check."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist


def main():
    import __graft_entry__ as ge
    from alignnet_b200 import dist as D
    rank, world, local = D.init("nccl")
    if rank == 0:
        ge.build()
    dist.barrier()
    from alignnet_b200 import engine, synth
    dev = torch.device(f"cuda:{local}")
    B, N, K = 128 * world, 200, 3
    host = synth.make_batch_fast(B, N, seed=4321)                       # the same global batch on every rank
    full = {k: torch.from_numpy(v).to(dev) for k, v in host.items()}
    shard = {k: v.contiguous() for k, v in D.shard_batch(full, rank, world).items()}
    shards = [{k: v.contiguous() for k, v in D.shard_batch(full, r, world).items()} for r in range(world)]
    report = {"world": world}

    def fresh():
        return engine.Engine(engine.shipped_arch(), str(dev), "bf16", seed=11)

    def local_grad(e, batch, seed):
        ep = e.forward(batch["pcs1"], batch["pcs2"], True, 0.5, None, seed=seed)
        e.backward(batch["pcs1"], batch["pcs2"], batch, ep)
        return e.grads.clone()

    # ---- 1. all-reduced gradient == mean of shard gradients
    e = fresh()
    g_local = local_grad(e, shard, seed=1)
    gathered = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(gathered, g_local)
    g_red = g_local.clone()
    scale = D.allreduce_grads(g_red)
    mean_gathered = torch.stack(gathered).double().mean(0)
    err = float((g_red.double() * scale - mean_gathered).abs().max() / mean_gathered.abs().max())
    report["allreduce_vs_gathered_mean_rel"] = err
    assert scale == 1.0 / world and err <= 1e-6, err
    if rank == 0:
        own = [local_grad(fresh(), s, seed=1) for s in shards]          # every shard recomputed on THIS GPU
        mean_own = torch.stack(own).double().mean(0)
        err_own = float((mean_own - mean_gathered).abs().max() / mean_gathered.abs().max())
        cos = float(torch.dot(mean_own, mean_gathered) / (mean_own.norm() * mean_gathered.norm()))
        report["recomputed_on_rank0_rel"], report["recomputed_on_rank0_cos"] = err_own, cos
        assert err_own <= 5e-3 and cos >= 0.9999, (err_own, cos)        # backward fp32 reductions reorder, nothing else

    # ---- 2. K steps: identical parameters on every rank (graph path with the collective inside, then eager path)
    for mode in ("graph", "eager"):
        e = fresh()
        losses = []
        for t in range(K):
            if mode == "graph":
                l = e.train_step_graph(shard, lr=1e-3, bn_decay=0.5, allreduce=D.allreduce_grads)
            else:
                l = e.train_step(shard, lr=1e-3, bn_decay=0.5, allreduce=D.allreduce_grads)
            losses.append(float(l[0].cpu()))
        torch.cuda.synchronize()
        for name, buf in (("params", e.params), ("adam_m", e.adam_m), ("adam_v", e.adam_v)):
            ref = buf.clone()
            dist.broadcast(ref, 0)
            same = bool(torch.equal(ref, buf))
            flag = torch.tensor([1 if same else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            report[f"{mode}_{name}_identical"] = bool(flag.item())
            assert flag.item() == 1, (mode, name)
        report[f"{mode}_losses_rank{rank}"] = losses
        if mode == "graph":
            report["allreduce_in_graph"] = e._ar_in_graph
            p_graph = e.params.clone()
        else:
            # same maths, but the two paths number their dropout seeds differently (t + 1 vs t): after K Adam steps no
            # weight can be further apart than 2 lr per step
            report["graph_vs_eager_params_max_abs"] = float((e.params - p_graph).abs().max())
            assert float((e.params - p_graph).abs().max()) <= 2 * K * 1e-3 * 1.05

    # ---- 3. G ranks at global batch B == G single-GPU replicas at B / G with averaged gradients (rank 0 alone)
    e = fresh()
    for t in range(K):
        e.train_step(shard, lr=1e-3, bn_decay=0.5, seed=100 + t, allreduce=D.allreduce_grads)
    p_dp = e.params.clone()
    if rank == 0:
        solo = fresh()
        replicas = [fresh() for _ in range(world)]
        for t in range(K):
            gs = []
            for r, rep in enumerate(replicas):
                rep.params.copy_(solo.params)
                rep.params_changed()
                gs.append(local_grad(rep, shards[r], seed=100 + t))
            solo.grads.copy_(torch.stack(gs).sum(0))
            solo.adam_step(1e-3, grad_scale=1.0 / world)
        # The two computations differ only through the backward's fp32 reductions (1e-3 of the LARGEST gradient element,
        # tests/test_gpu_determinism.py -- i.e. tens of percent of a typical small one).  Adam normalises every element,
        # so small, noisy elements move their weights by comparable amounts in slightly different directions: the
        # bounds are 2 lr per step for any weight and the direction of the whole K-step update (measured 0.91 .. 0.96 over
        # boxes and kernel versions; the averaged GRADIENT itself agrees to cos 0.999999, checked above).
        p0 = fresh().params
        u_solo, u_dp = (solo.params - p0).double(), (p_dp - p0).double()
        d = float((u_solo - u_dp).abs().max())
        cos = float(torch.dot(u_solo, u_dp) / (u_solo.norm() * u_dp.norm()))
        report["dp_vs_replicas_params_max_abs"], report["dp_vs_replicas_update_cos"] = d, cos
        report["params_moved_max_abs"] = float(u_solo.abs().max())
        assert d <= 2 * K * 1e-3 * 1.05 and cos >= 0.8 and float(u_solo.abs().max()) > 1e-3, (d, cos)

    # ---- 4. the momentum optimiser (train.py:211-212) under data parallelism: identical parameters / accumulators on every
    #         rank through the graph path, and -- the update being LINEAR in the gradient, unlike Adam's -- the data-parallel
    #         run equals single-GPU replicas with averaged gradients tightly
    e = fresh()
    e.set_optimizer("momentum", momentum=0.9)
    for t in range(K):
        e.train_step_graph(shard, lr=1e-4, bn_decay=0.5, allreduce=D.allreduce_grads)
    torch.cuda.synchronize()
    for name, buf in (("params", e.params), ("mom_accum", e.mom_accum)):
        ref = buf.clone()
        dist.broadcast(ref, 0)
        flag = torch.tensor([1 if torch.equal(ref, buf) else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        report[f"momentum_graph_{name}_identical"] = bool(flag.item())
        assert flag.item() == 1, name
    assert not bool(e.adam_m.any()) and e.step == K
    e = fresh()
    e.set_optimizer("momentum", momentum=0.9)
    p_dp = []
    for t in range(K):
        e.train_step(shard, lr=1e-4, bn_decay=0.5, seed=100 + t, allreduce=D.allreduce_grads)
        p_dp.append(e.params.clone())
    if rank == 0:
        solo = fresh()
        solo.set_optimizer("momentum", momentum=0.9)
        replicas = [fresh() for _ in range(world)]
        p0 = fresh().params
        for t in range(K):
            gs = []
            for r, rep in enumerate(replicas):
                rep.params.copy_(solo.params)
                rep.params_changed()
                gs.append(local_grad(rep, shards[r], seed=100 + t))
            solo.grads.copy_(torch.stack(gs).sum(0))
            solo.momentum_step(1e-4, grad_scale=1.0 / world)
            u_solo, u_dp = (solo.params - p0).double(), (p_dp[t] - p0).double()
            cos = float(torch.dot(u_solo, u_dp) / (u_solo.norm() * u_dp.norm()))
            rel = float((u_solo - u_dp).norm() / u_solo.norm())
            report[f"momentum_dp_vs_replicas_step{t + 1}_update_cos"], report[f"momentum_dp_vs_replicas_step{t + 1}_rel_l2"] = cos, rel
            if t == 0:
                # one step: the update is -lr * (mean shard gradient), so the two computations agree as tightly as the
                # gradients do (check 1).  Later steps are reported, not bounded: the parameters of the two runs then differ
                # in the last bits, and this loss is discontinuous in them at a fresh initialisation (the batch-level choice
                # between the angle loss and its 180-degree twin, quirk Q5, is a near-tie; arg-max bins) -- measured cos
                # 0.94 after three steps on 2 GPUs, the same spread the Adam comparison above shows
                assert cos >= 0.9999 and rel <= 2e-2 and float(u_solo.abs().max()) > 0, (cos, rel)
    if rank == 0:
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        msg = traceback.format_exc()
        print(f"[rank {os.environ.get('RANK')}] FAILED\n{msg}", flush=True)
        out = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(out):
            with open(os.path.join(out, f"dp_worker_rank{os.environ.get('RANK')}.log"), "w") as fh:
                fh.write(msg)
        raise
