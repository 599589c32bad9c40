"""world_size-2 gloo test (CPU) of the data-parallel host logic: batch sharding and the single
flat-gradient all-reduce with the 1/world scale the Adam kernel folds in."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from alignnet_b200 import dist as D
    r, w, _ = D.init("gloo")
    assert (r, w) == (rank, world)
    batch = {"pcs1": torch.arange(7 * 4 * 3, dtype=torch.float32).reshape(7, 4, 3), "translations": torch.arange(21.).reshape(7, 3)}
    shard = D.shard_batch(batch, rank, world)
    lo, hi = D.shard_bounds(7, rank, world)
    assert shard["pcs1"].shape[0] == hi - lo and torch.equal(shard["translations"], batch["translations"][lo:hi])
    g = torch.full((1000,), float(rank + 1))
    scale = D.allreduce_grads(g)
    out[rank] = (float(g[0]), scale, lo, hi)
    dist.destroy_process_group()


def test_gloo_world2_allreduce_and_sharding():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][0] == res[1][0] == 3.0           # 1 + 2 summed on both ranks
    assert res[0][1] == res[1][1] == 0.5           # scale folded into Adam
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 4, 4, 7)   # contiguous, covers the batch once


def test_shard_bounds_cover_batch_exactly_once():
    from alignnet_b200 import dist as D
    for batch in (1, 7, 8, 4096, 16384):
        for world in (1, 2, 4, 8):
            spans = [D.shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_schedules_match_reference_formulas():
    from alignnet_b200 import config, schedules
    cfg = config.load_shipped("SynthCars")          # lr 0.005, step 30 epochs, rate 0.5, batch 128
    nb = 10
    assert schedules.learning_rate(cfg, 0, nb) == 0.005
    assert schedules.learning_rate(cfg, 30 * nb, nb) == 0.0025
    assert schedules.learning_rate(cfg, 10 ** 7, nb) == 1e-5
    assert schedules.bn_decay(cfg, 0, nb) == 0.5
    assert schedules.bn_decay(cfg, 30 * nb, nb) == 0.75
    assert schedules.bn_decay(cfg, 10 ** 7, nb) == 0.99
    config.reset_config()
