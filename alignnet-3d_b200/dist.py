"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch on the
GPU box, gloo in CPU tests).  The hot path shards by cloud pair: each rank runs forward + backward
on its contiguous slice of the batch and the ONLY collective is one all-reduce (sum) of the flat
fp32 gradient buffer per step; the 1/world scaling is folded into the Adam kernel
(an3d_adam_step's grad_scale).  BN batch statistics and the loss's [B,B] coupling are per shard, so G
ranks at global batch B equal G reference replicas at batch B/G with averaged gradients."""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def init(backend: str | None = None) -> Tuple[int, int, int]:
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* from the environment (torchrun)."""
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device(f"cuda:{local}")
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: Dict[str, "torch.Tensor"], rank: int, world: int) -> Dict[str, "torch.Tensor"]:
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


def allreduce_grads(flat_grads: torch.Tensor) -> float:
    """In-place sum over ranks of the flat gradient buffer; returns the scale (1/world) the optimiser
    applies.  One collective per step."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()
    return 1.0
