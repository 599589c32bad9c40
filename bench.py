#!/usr/bin/env python
"""Benchmark of the AlignNet-3D tp8 hot path on B200 (contract: see the task statement / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3]

Prints ONE JSON line on rank 0.  `value` = point-cloud pairs/s with inputs resident in HBM; `e2e` =
the same metric through the public engine API with pinned-host inputs (H2D of the batch and D2H of
the result inside the timed region); `roofline` = dominant-kernel tensor-pipe roofline from CUDA
events recorded by the library around that kernel; `cpu_baseline` = the CPU oracle (a restatement
of the reference's TF1 graph -- TensorFlow 1.8 is not installable here) timed on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: SynthCars synthetic B=1024, 200 pts, bf16, 1xB200 fwd-only pose inference
    "c2": dict(name="c2: SynthCars-shaped synthetic B=1024 N=200 bf16 forward-only pose inference", B=1024, N=200,
               train=False, persons=0.0),
    # BASELINE.json configs[2]: SynthCarsPersons synthetic B=4096, 200 pts, bf16, fwd+bwd training step
    "c3": dict(name="c3: SynthCarsPersons-shaped synthetic B=4096 N=200 bf16 fwd+bwd+Adam training step", B=4096,
               N=200, train=True, persons=0.2),
    # BASELINE.json configs[3] / [4], per-GPU shard (global batch 8192 / 16384 over 4 / 8 GPUs = 2048 per GPU)
    "c4": dict(name="c4: KITTITrackletsCars-shaped synthetic B=2048/GPU N=512 bf16 fwd+bwd+Adam (+ grad all-reduce)",
               B=2048, N=512, train=True, persons=0.0),
    "c5": dict(name="c5: KITTITrackletsCarsHard-shaped synthetic B=2048/GPU N=1024 bf16 fwd+bwd+Adam (+ grad all-reduce)",
               B=2048, N=1024, train=True, persons=0.0),
    # not a BASELINE config: the reference's own configs/default.json (five-layer conv stacks, N=1024, batch 64) and the
    # same at a batch that fills the GPU -- the architecture the fused kernels do not cover (layer-by-layer tensor-core path)
    "cdef": dict(name="configs/default.json of the reference: B=64 N=1024, [128,128,256] + two [64,64,64,128,1024] stacks, training step",
                 B=64, N=1024, train=True, persons=0.0, arch="default"),
    "cdef512": dict(name="configs/default.json architecture at B=512 N=1024, training step", B=512, N=1024, train=True,
                    persons=0.0, arch="default"),
}
DEFAULT_WORKLOAD = "c3"


def flops_per_pair(N: int, train: bool) -> float:
    """MACs(N) = 509,056 N + 2,571,008; fwd = 2 MACs, fwd+bwd = 6 MACs."""
    macs = 509056.0 * N + 2571008.0
    return (6.0 if train else 2.0) * macs


def emb_full_flops_per_pair(N: int) -> float:
    """Algorithmic FLOPs of one launch-unit of the dominant kernel (conv-stack full pass): both
    clouds of a pair through 3->64->128->C3 for the three stages = 2 * N * 254,528 MACs * 2."""
    return 2.0 * 2.0 * N * 254528.0


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, average over the six launches of one c3 step,
    from the committed `ncu --set full` capture (profiles/r2_roofline_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "r2_roofline_traffic.json")
    if os.path.exists(path):
        return json.load(open(path)).get("conv_stack_fwd_kernel<1>")
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, wl):
    """The reference's own CPU implementation of the path.  TensorFlow 1.8 cannot be installed in this
    image (no wheel, Python 3.12, no network), so this arm times the oracle -- the torch-CPU fp32
    restatement of models/tp8.py + utils/tf_util.py -- on all host cores, on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import arch as A, torch_ref as TR
    from alignnet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    arch = A.Arch()
    Bs = 32                                    # BASELINE.json configs[0] batch: the reference's own timing batch
    params = A.init_params(arch, 0)
    state = A.init_state(arch)
    batch = synth.make_batch_fast(Bs, wl["N"], seed=1234)
    p32 = TR.to_torch(params, torch.float32, requires_grad=wl["train"])
    s32 = TR.to_torch(state, torch.float32)
    b32 = TR.to_torch(batch, torch.float32)
    names = [n for n, _ in A.trainable_specs(arch)]
    m = {k: torch.zeros_like(v) for k, v in p32.items()}
    v = {k: torch.zeros_like(v) for k, v in p32.items()}

    def step(t):
        nonlocal s32
        if wl["train"]:
            ep, s32 = TR.get_model(b32["pcs1"], b32["pcs2"], arch, p32, s32, True, 0.5, None)
            loss = TR.get_loss(b32["translations"], b32["rel_angles"], b32["pc1_centers"], b32["pc2_centers"],
                               b32["pc1_angles"], b32["pc2_angles"], ep, arch)
            grads = torch.autograd.grad(loss, [p32[n] for n in names], allow_unused=True)
            with torch.no_grad():
                lr_t = 0.005 * (1 - 0.999 ** t) ** 0.5 / (1 - 0.9 ** t)
                for n, g in zip(names, grads):
                    if g is None:
                        continue
                    m[n].mul_(0.9).add_(g, alpha=0.1)
                    v[n].mul_(0.999).addcmul_(g, g, value=0.001)
                    p32[n].sub_(lr_t * m[n] / (v[n].sqrt() + 1e-8))
            return float(loss.detach())
        with torch.no_grad():
            ep, _ = TR.get_model(b32["pcs1"], b32["pcs2"], arch, p32, s32, False)
        return float(ep["pred_translations"].sum())

    for i in range(args.warmup):
        step(i + 1)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i + 1)
    dt = time.perf_counter() - t0
    value = Bs * args.steps / dt
    line = {
        "impl": "reference", "metric": "point-cloud pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "sample": f"B={Bs} N={wl['N']} per step"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of B={Bs}, N={wl['N']}, "
                                   f"{'fwd+bwd+Adam' if wl['train'] else 'eval forward'}, torch-CPU fp32 restatement "
                                   "of the TF1 graph (TensorFlow 1.8 not installable)"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline_sample(wl, budget_s=12.0):
    import torch
    from oracle import arch as A, torch_ref as TR
    from alignnet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    arch = A.Arch()
    Bs = 32
    params, state = A.init_params(arch, 0), A.init_state(arch)
    batch = synth.make_batch_fast(Bs, wl["N"], seed=1234)
    p32 = TR.to_torch(params, torch.float32, requires_grad=wl["train"])
    s32 = TR.to_torch(state, torch.float32)
    b32 = TR.to_torch(batch, torch.float32)
    names = [n for n, _ in A.trainable_specs(arch)]

    def step():
        if wl["train"]:
            ep, _ = TR.get_model(b32["pcs1"], b32["pcs2"], arch, p32, s32, True, 0.5, None)
            loss = TR.get_loss(b32["translations"], b32["rel_angles"], b32["pc1_centers"], b32["pc2_centers"],
                               b32["pc1_angles"], b32["pc2_angles"], ep, arch)
            torch.autograd.grad(loss, [p32[n] for n in names], allow_unused=True)
        else:
            with torch.no_grad():
                TR.get_model(b32["pcs1"], b32["pcs2"], arch, p32, s32, False)

    def run(budget):
        step()
        n, t0 = 0, time.perf_counter()
        while True:
            step(); n += 1
            dt = time.perf_counter() - t0
            if dt > budget or n >= 200:
                return n, dt

    step()
    n, dt = run(budget_s)
    torch.set_num_threads(1)                # the reference's "single process" wording
    n1, dt1 = run(budget_s / 3)
    torch.set_num_threads(cores)
    return {"value": Bs * n / dt, "unit": "pairs/s", "cores": cores, "kind": "port", "value_1thread": Bs * n1 / dt1,
            "sample": f"{n} steps of B={Bs}, N={wl['N']}, {'fwd+bwd' if wl['train'] else 'eval forward'}; torch-CPU fp32 "
                      "restatement of the reference TF1 graph (TensorFlow 1.8 not installable here)"}


class Ctx:
    """Process-wide state of one bench run (rank layout, device, library handle)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        import __graft_entry__ as ge
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.rank == 0:
            ge.build()
        torch.cuda.set_device(self.local)
        self.numa = None
        if self.world > 1:
            from alignnet_b200 import dist as an3d_dist
            self.numa = an3d_dist.bind_to_gpu_numa_node(self.local)    # before any pinned allocation / NCCL thread exists
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local}"))
            dist.barrier()
        if self.rank != 0:
            ge.build()
        from alignnet_b200 import _lib
        self.lib = _lib.load()
        self.dev = torch.device(f"cuda:{self.local}")
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)   # > 126 MB L2


def measure(ctx, wl, precision, steps, warmup, no_graph=False, tags=True):
    """Times one workload: `steps` device-timed steps with HBM-resident inputs (L2 flushed in between), the tagged
    kernels on the same number of eager steps, and the end-to-end loop with pinned-host inputs.  Returns a dict."""
    import torch
    import torch.distributed as dist
    from alignnet_b200 import engine, synth
    rank, world, dev, lib, flush = ctx.rank, ctx.world, ctx.dev, ctx.lib, ctx.flush
    B, N, train = wl["B"], wl["N"], wl["train"]
    eng = engine.Engine(engine.default_arch() if wl.get("arch") == "default" else engine.shipped_arch(), str(dev), precision, seed=0)
    host = synth.make_batch_fast(B, N, seed=1234 + (2 if train else 1) + rank)
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in host.items()}
    resident = {k: t.to(dev) for k, t in pinned.items()}
    staging = {k: torch.empty_like(t, device=dev) for k, t in pinned.items()}
    staging2 = {k: torch.empty_like(t, device=dev) for k, t in pinned.items()}     # second buffer set: double-buffered H2D
    in_keys = list(host.keys()) if train else ["pcs1", "pcs2"]
    h2d_bytes = sum(pinned[k].numel() * 4 for k in in_keys)
    out_host = torch.empty((B, 3), dtype=torch.float32).pin_memory()
    ang_host = torch.empty((B,), dtype=torch.float32).pin_memory()
    loss_host = torch.empty(20, dtype=torch.float32).pin_memory()

    def allreduce(g):
        dist.all_reduce(g)
        return 1.0 / world

    # eval-mode forward needs populated BN shadows (zero-initialised shadows are degenerate, quirk Q7):
    # ten training-mode forwards with decay 0.5 fill them.
    if not train:
        for i in range(10):
            eng.forward(resident["pcs1"], resident["pcs2"], True, 0.5, None, seed=i)
    torch.cuda.synchronize()

    ar = allreduce if world > 1 else None

    def step(batch):
        """One step through the engine's public API; CUDA-graph replay unless --no-graph."""
        if train:
            if no_graph:
                return eng.train_step(batch, lr=0.005, bn_decay=0.5, allreduce=ar)
            return eng.train_step_graph(batch, lr=0.005, bn_decay=0.5, allreduce=ar)
        if no_graph:
            return eng.forward(batch["pcs1"], batch["pcs2"], False)
        return eng.forward_graph(batch["pcs1"], batch["pcs2"])

    def step_eager(batch):
        if train:
            return eng.train_step(batch, lr=0.005, bn_decay=0.5, allreduce=ar)
        return eng.forward(batch["pcs1"], batch["pcs2"], False)

    def step_e2e():
        for k in in_keys:
            staging[k].copy_(pinned[k], non_blocking=True)
        out = step(staging)
        if train:
            loss_host.copy_(out, non_blocking=True)
        else:
            out_host.copy_(out["pred_translations"], non_blocking=True)
            ang_host.copy_(eng.pred_angles(out), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    copy_stream = torch.cuda.Stream(device=dev)
    sets = [staging, staging2]

    def e2e_pipelined(n):
        """The serving / training loop a user of the API writes: step i computes on buffer set i % 2 while the pinned
        host batch of step i + 1 is copied into the other set on a copy stream.  EVERY step's H2D copy and D2H read of
        the result is issued inside the timed region; the region is one CUDA-event bracket over all `n` steps
        (per-step working set >> L2, so no flush is needed between them)."""
        main = torch.cuda.current_stream()
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]

        def issue_copy(i):
            s_ = i % 2
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(done[s_])           # the step that last read this set has finished
                for k in in_keys:
                    sets[s_][k].copy_(pinned[k], non_blocking=True)
                ready[s_].record(copy_stream)

        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        copy_stream.wait_event(e0)
        issue_copy(0)
        for i in range(n):
            if i + 1 < n:
                issue_copy(i + 1)
            main.wait_event(ready[i % 2])
            out = step(sets[i % 2])
            done[i % 2].record(main)
            if train:
                loss_host.copy_(out, non_blocking=True)
            else:
                out_host.copy_(out["pred_translations"], non_blocking=True)
                ang_host.copy_(eng.pred_angles(out), non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, n):
        total = 0.0
        for _ in range(n):
            flush.fill_(1)                                   # evict L2 between timed iterations (untimed)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        t = torch.tensor([total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(warmup, 3)):
        step(resident)
        step_e2e()
    torch.cuda.synchronize()

    ms_total = timed(lambda: step(resident), steps)
    res = dict(B=B, N=N, train=train, precision=precision, arch=wl.get("arch", "shipped"), steps=steps, ms_per_step=ms_total / steps,
               value=B * world * steps / (ms_total * 1e-3), h2d_bytes=h2d_bytes, d2h_bytes=80 if train else B * 16,
               allreduce_in_graph=getattr(eng, "_ar_in_graph", None))
    if tags:
        # per-kernel device times: CUDA events cannot bracket kernels inside a replayed graph, so the tagged kernels
        # are timed on the same number of EAGER steps of the same workload right after the timed region (the library
        # serialises the two branch streams while it profiles)
        launches0 = lib.an3d_launch_count()
        lib.an3d_profile_begin()
        timed(lambda: step_eager(resident), steps)
        ms_tags, n_tags = (C.c_float * 8)(), (C.c_int32 * 8)()
        lib.an3d_profile_end(C.byref(ms_tags), C.byref(n_tags))
        res["launches"] = int((lib.an3d_launch_count() - launches0) // steps)
        res["ms_tags"] = {str(i): ms_tags[i] / steps for i in range(8) if n_tags[i]}
    for k_ in in_keys:                      # capture the graph of the second buffer set outside the timed region
        staging2[k_].copy_(pinned[k_], non_blocking=True)
    step(staging2)
    torch.cuda.synchronize()
    e2e_pipelined(2)
    ms_e2e = e2e_pipelined(steps)
    res["e2e_value"] = B * world * steps / (ms_e2e * 1e-3)
    del eng, resident, staging, staging2
    torch.cuda.empty_cache()
    return res


def config_entry(r, pk, world):
    """One entry of the `configs` sub-object: a BASELINE.json config other than the headline one, measured in the
    same run with the same method."""
    e = {"batch_per_gpu": r["B"], "num_points": r["N"], "mode": "train" if r["train"] else "eval", "dtype": r["precision"],
         "n_gpus": world, "steps": r["steps"], "ms_per_step": r["ms_per_step"], "value": r["value"], "unit": "pairs/s",
         "e2e": {"value": r["e2e_value"], "unit": "pairs/s", "h2d_bytes_per_step": r["h2d_bytes"],
                 "d2h_bytes_per_step": r["d2h_bytes"]},
         "whole_step_tensor_frac": r["value"] / world * flops_per_pair(r["N"], r["train"]) / (pk["bf16_sustained"] * 1e12)}
    if r["precision"] == "fp32" or r.get("arch") != "shipped":
        e["whole_step_tensor_frac"] = None        # fp32 mode: CUDA cores; other architectures: flops_per_pair() does not apply
    if r.get("arch") != "shipped":
        e["architecture"] = "reference configs/default.json:8-22 (conv stacks [128,128,256] / [64,64,64,128,1024] x 2, 36 bins)"
    # (bf16x3 / bf16x6: algorithmic FLOPs over the bf16 peak -- the split modes issue 3x / 6x those FLOPs on the tensor pipe)
    return e


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    ctx = Ctx()
    rank, world = ctx.rank, ctx.world
    B, N, train = wl["B"], wl["N"], wl["train"]
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    r = measure(ctx, wl, args.precision, args.steps, max(args.warmup, 3), no_graph=args.no_graph, tags=True)
    clocks = sampler.stop() if rank == 0 else None

    # The other BASELINE.json configs, measured in the same run (every rank takes part: the training ones all-reduce):
    # c2 (eval forward), the fp32 parity mode on the headline workload, and -- at the GPU counts BASELINE names for
    # them -- c4 (4 GPUs) / c5 (8 GPUs) with their real global batch.
    extra = {}
    if args.workload == DEFAULT_WORKLOAD and args.precision == "bf16" and not args.no_configs:
        short = max(5, args.steps // 2)
        extra["c2"] = measure(ctx, WORKLOADS["c2"], "bf16", short, 3, tags=False)
        extra["c3_fp32"] = measure(ctx, WORKLOADS["c3"], "fp32", 3, 3, tags=False)
        # the parity tolerance on the tensor cores: every GEMM as six (three) bf16 tcgen05 products of split operands
        extra["c3_bf16x6"] = measure(ctx, WORKLOADS["c3"], "bf16x6", 3, 3, tags=False)
        extra["c3_bf16x3"] = measure(ctx, WORKLOADS["c3"], "bf16x3", 3, 3, tags=False)
        # the reference's default architecture (not [64,128,C]): bf16 layer by layer on the tensor cores vs the fp32 mode
        extra["default_arch_bf16"] = measure(ctx, WORKLOADS["cdef"], "bf16", short, 3, tags=False)
        extra["default_arch_b512_bf16"] = measure(ctx, WORKLOADS["cdef512"], "bf16", 3, 3, tags=False)
        extra["default_arch_b512_fp32"] = measure(ctx, WORKLOADS["cdef512"], "fp32", 3, 3, tags=False)
        if world == 4:
            extra["c4"] = measure(ctx, WORKLOADS["c4"], "bf16", short, 3, tags=False)
        if world == 8:
            extra["c5"] = measure(ctx, WORKLOADS["c5"], "bf16", short, 3, tags=False)

    if rank == 0:
        pk = peaks()
        value, e2e_value = r["value"], r["e2e_value"]
        # dominant kernel: conv-stack full pass (tag 1): 6 launches per step (3 stages x 2 branches)
        k_ms = r["ms_tags"].get("1", 0.0)
        k_flops = B * emb_full_flops_per_pair(N)
        achieved = k_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        line = {
            "metric": "point-cloud pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": B, "num_points": N, "mode": "train" if train else "eval",
                       "parallelism": f"dp{world}", "l2": "flushed between timed iterations (256 MB write)",
                       "launch": "eager stream launches" if args.no_graph else "CUDA-graph replay of the step (Engine.train_step_graph / forward_graph)",
                       "collective": None if world == 1 else ("one NCCL all-reduce of the flat fp32 gradient per step, " +
                                                              ("captured inside the step graph" if r["allreduce_in_graph"] else "issued between two graph segments")),
                       "host_binding": ctx.numa,
                       "e2e": "pinned host batch -> H2D every step (double-buffered on a copy stream, overlapped with the previous step) -> step -> D2H of the result; one event bracket over all steps",
                       "whole_step_tensor_frac": value / world * flops_per_pair(N, train) / (pk["bf16_sustained"] * 1e12)},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": r["h2d_bytes"], "d2h_bytes_per_step": r["d2h_bytes"]},
            "gpu_launches": r["launches"],
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / pk["bf16_sustained"], "traffic": measured_traffic() if train else None,
                         "traffic_source": "constant from the committed ncu --set full capture (profiles/), not measured in this run",
                         "kernel": "conv_stack_fwd_kernel (full pass), 6 launches/step", "kernel_ms_per_step": k_ms,
                         "peak_source": pk["source"] + ", sustained bf16"},
            "clocks": clocks,
            "kernel_ms_by_tag": r["ms_tags"],
        }
        if extra:
            line["configs"] = {k: config_entry(v, pk, world) for k, v in extra.items()}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample(wl)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16x3", "bf16x6"])
    ap.add_argument("--no-configs", action="store_true", help="headline workload only (skip the `configs` sub-object)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (A/B runs of a kernel switch)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the stream instead of replaying a CUDA graph")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
