"""Learning-rate and BN-decay schedules of the reference driver (train.py:133-174)."""
from __future__ import annotations

import math


def _decay_step(ext, batch_size: int, num_batches_per_epoch: int) -> int:
    step = ext.step
    if ext.per == "epoch":
        step *= batch_size * num_batches_per_epoch
    elif ext.per != "step":
        raise ValueError(f"per={ext.per!r}")
    return step


def learning_rate(cfg, global_step: int, num_batches_per_epoch: int) -> float:
    """train.py:133-156: staircase exponential decay of cfg.training.learning_rate, clipped at 1e-5."""
    ext = cfg.training.lr_extension
    if ext.mode != "decay":
        raise ValueError("only lr_extension.mode == 'decay' is implemented (the reference asserts on 'clr')")
    bs = cfg.training.batch_size
    lr = cfg.training.learning_rate * ext.rate ** math.floor(global_step * bs / _decay_step(ext, bs, num_batches_per_epoch))
    return max(lr, 0.00001)


def bn_decay(cfg, global_step: int, num_batches_per_epoch: int) -> float:
    """train.py:159-174: min(clip, 1 - init * rate^floor(step*B/decay_step))."""
    ext = cfg.training.bn_extension
    assert ext.mode == "decay"
    bs = cfg.training.batch_size
    momentum = ext.init * ext.rate ** math.floor(global_step * bs / _decay_step(ext, bs, num_batches_per_epoch))
    return min(ext.clip, 1 - momentum)
