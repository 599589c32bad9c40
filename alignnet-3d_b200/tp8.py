"""Drop-in mirror of the reference's `models/tp8.py` call surface on the B200 engine.

`train.py` of the reference uses exactly four attributes of its MODEL module (train.py:190,200,
201,453-455): `placeholder_inputs`, `get_model`, `get_loss`, `classLogits2angle`.  The same four
names, argument orders and return structures are provided here; tensors are torch CUDA tensors
instead of TF graph nodes, and the work is done by libalignnet_b200.so (no TensorFlow, no CPU
fallback).  Like the reference module, this one reads the global config (`config.configGlobal`).

    from alignnet_b200 import config, tp8 as MODEL
    config.load_config("configs/SynthCars.json")          # the reference's own JSON files load unchanged
    pcs1, pcs2, translations, rel_angles, pc1c, pc2c, pc1a, pc2a = MODEL.placeholder_inputs(B, N)
    end_points = MODEL.get_model(pcs1, pcs2, is_training=False)
    loss = MODEL.get_loss(pcs1, pcs2, translations, rel_angles, pc1c, pc2c, pc1a, pc2a, end_points)
    angles = MODEL.classLogits2angle(end_points['pred_pc1angle_logits'].cpu().numpy())
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import config as _config
from . import engine as _engine

cfg = _config.configGlobal
_ENGINE: Optional[_engine.Engine] = None
_ENGINE_KEY = None


def get_engine(precision: Optional[str] = None) -> _engine.Engine:
    """The process-wide engine for the current config (the analogue of the reference's tf.Session +
    variables).  Re-created when the architecture in `cfg` changes."""
    global _ENGINE, _ENGINE_KEY
    arch = _config.arch_from_config(cfg)
    key = (bytes(arch), precision or (_ENGINE.precision if _ENGINE else "fp32"))
    if _ENGINE is None or _ENGINE_KEY != key:
        _ENGINE = _engine.Engine(arch, f"cuda:{getattr(cfg, 'gpu_index', 0)}", key[1])
        _ENGINE_KEY = key
    return _ENGINE


def placeholder_inputs(batch_size: int, num_point: int):
    """models/tp8.py:13-23: the eight input buffers (here: uninitialised CUDA tensors to fill)."""
    dev = torch.device(f"cuda:{getattr(cfg, 'gpu_index', 0)}")
    c = cfg.data.num_channels
    mk = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    return (mk(batch_size, num_point, c), mk(batch_size, num_point, c), mk(batch_size, 3), mk(batch_size, 1),
            mk(batch_size, 3), mk(batch_size, 3), mk(batch_size, 1), mk(batch_size, 1))


def get_model(pcs1: torch.Tensor, pcs2: torch.Tensor, is_training, bn_decay=None) -> Dict[str, torch.Tensor]:
    """models/tp8.py:135-158: returns the end_points dict with the reference's 8 keys."""
    return get_engine().forward(pcs1, pcs2, bool(is_training), bn_decay)


def get_loss(pcs1, pcs2, translations, rel_angles, pc1_centers, pc2_centers, pc1_angles, pc2_angles, end_points):
    """models/tp8.py:401-407: 0-d tensor = per_transform_loss of the configured loss.  `p2p` (:374-398, selected by no
    shipped config) is evaluated as the reference computes it (quirk Q6) -- value only, the training step is `separate`."""
    labels = dict(translations=translations, rel_angles=rel_angles, pc1_centers=pc1_centers, pc2_centers=pc2_centers,
                  pc1_angles=pc1_angles, pc2_angles=pc2_angles)
    if cfg.training.loss.loss == "p2p":
        return get_engine().loss_p2p(pcs1, labels, end_points)[0]
    if cfg.training.loss.loss != "separate":
        raise ValueError(f"training.loss.loss={cfg.training.loss.loss!r}: the reference asserts False here (tp8.py:407)")
    return get_engine().loss(labels, end_points)[0]


def tf_transform_pcs(pcs, translations=None, angles=None, rotation_centers=None):
    """models/tp8.py:361-371 as coded (quirk Q6), on the device."""
    from . import engine as _engine
    return _engine.transform_pcs(pcs, translations, angles, rotation_centers)


def classLogits2angle(logits: np.ndarray, to_label_format: bool = True) -> np.ndarray:
    """models/tp8.py:241-244 (host decode with the unscaled residual, quirk Q1) -- evaluated by the
    library's decode kernel, like everything else on this path."""
    eng = get_engine()
    t = torch.from_numpy(np.ascontiguousarray(logits, dtype=np.float32)).to(eng.device)
    return eng.decode_angles(t, scaled=False).cpu().numpy().astype(np.float64)


def train_op(batch: Dict[str, torch.Tensor], learning_rate: float, bn_decay: float, allreduce=None) -> torch.Tensor:
    """`sess.run([train_op, loss])` of train.py:368: one optimiser step; returns the loss vector."""
    return get_engine().train_step(batch, learning_rate, bn_decay, allreduce=allreduce)
