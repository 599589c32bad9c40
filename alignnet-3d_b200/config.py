"""Config system with the reference's JSON schema (config.py + configs/*.json of the reference).

Same keys, same overlay semantics (run JSON overlaid on the defaults into one global mutable
namespace, reference config.py:32-41,66-72), same derived fields (`name`, `data.basename`,
`logging.logdir`).  Differences: no import-time side effects, the circular provider import is
gone, and the split files are only read when they exist (synthetic runs have no dataset).
The defaults and the shipped run configs are expressed in code, not copied JSON files.
"""
from __future__ import annotations

import copy
import json
import os
from typing import Any, Dict

from . import _lib


class NameSpace:
    """Attribute-style nested config (reference config.py:9-30)."""

    def __repr__(self) -> str:
        return "config:\n" + self.repr(4)[:-1]

    def reset(self) -> None:
        self.__dict__ = dict()

    def repr(self, indent: int) -> str:
        s = ""
        for k, v in self.__dict__.items():
            if isinstance(v, NameSpace):
                s += "%s%s:\n%s" % (" " * indent, k, v.repr(indent + 4))
            else:
                s += "%s%s: %s\n" % (" " * indent, k, v)
        return s

    def has(self, key: str) -> bool:
        return key in self.__dict__


def dump_to_namespace(ns: NameSpace, d: Dict[str, Any]) -> None:
    for k, v in d.items():
        if isinstance(v, dict):
            if k not in ns.__dict__:
                ns.__dict__[k] = NameSpace()
            dump_to_namespace(ns.__dict__[k], v)
        else:
            ns.__dict__[k] = v


def namespace_to_dict(ns: NameSpace) -> Dict[str, Any]:
    return {k: (namespace_to_dict(v) if isinstance(v, NameSpace) else v) for k, v in ns.__dict__.items()}


def _schedule(step=30, rate=0.5, **extra):
    d = {"mode": "decay", "per": "epoch", "step": step, "rate": rate}
    d.update(extra)
    return d


def default_config() -> Dict[str, Any]:
    """Values of the reference's configs/default.json."""
    mlp = [[512, 256], 0.7]
    return {
        "data": {"basepath": "/home/gross/data/SynthCars", "num_channels": 3},
        "gpu_index": 0,
        "model": {
            "model": "tp8", "backbone": "pointnet", "num_points": 1024,
            "options": {
                "angle_factor": 1.0, "early_stage_factor": 0.1,
                "s1transformer": [[128, 128, 256], copy.deepcopy(mlp)],
                "s2transformer": [[64, 64, 64, 128, 1024], copy.deepcopy(mlp)],
                "embedding": [64, 64, 64, 128, 1024],
                "remaining_transform_prediction": copy.deepcopy(mlp),
            },
            "angles": {"num_bins": 36, "accept_inverted_angle": False},
        },
        "logging": {"basedir": "/home/gross/models/alignnet"},
        "evaluation": {"save_every_epoch": True},
        "training": {
            "batch_size": 64, "num_epochs": 100, "optimizer": {"optimizer": "adam"}, "learning_rate": 0.01,
            "lr_extension": _schedule(), "bn_extension": _schedule(init=0.5, clip=0.99),
            "loss": {"loss": "separate",
                     "options": {"soft_angle_classes": False, "soft_angle_classes_sigma_in_degree": 5.0}},
            "pretraining": {"model": ""},
        },
    }


def shipped_config(name: str) -> Dict[str, Any]:
    """Overlay equal to the reference's configs/<name>.json (all eight use one architecture;
    they differ in data path, angle_factor, accept_inverted_angle and pre-training checkpoint)."""
    kitti = name.startswith("KITTI")
    synth20 = name.startswith("Synth20")
    mlp = [[512, 256], 0.7]
    cfg: Dict[str, Any] = {
        "data": {"basepath": f"/home/gross/data/{name}"},
        "model": {
            "model": "tp8", "backbone": "pointnet", "num_points": 512,
            "options": {
                "angle_factor": 0.5 if kitti else 1.0, "early_stage_factor": 0.5,
                "s1transformer": [[64, 128, 256], copy.deepcopy(mlp)],
                "s2transformer": [[64, 128, 512], copy.deepcopy(mlp)],
                "embedding": [64, 128, 1024],
                "remaining_transform_prediction": copy.deepcopy(mlp),
            },
            "angles": {"num_bins": 50, "accept_inverted_angle": not synth20},
        },
        "training": {
            "num_epochs": 200, "batch_size": 128, "learning_rate": 0.005,
            "lr_extension": _schedule(), "bn_extension": _schedule(init=0.5, clip=0.99),
            "loss": {"loss": "separate"},
        },
    }
    pre = {"KITTITrackletsCars": "SynthCars/model-180", "KITTITrackletsCarsHard": "SynthCars/model-190",
           "KITTITrackletsCarsPersons": "SynthCarsPersons/model-65",
           "KITTITrackletsCarsPersonsHard": "SynthCarsPersons/model-65", "Synth20others": "Synth20/model-61"}
    if name in pre:
        cfg["training"]["pretraining"] = {"model": "/home/gross/models/alignnet/" + pre[name]}
    if name in ("KITTITrackletsCarsHard", "KITTITrackletsCarsPersons"):
        cfg["evaluation"] = {"save_every_epoch": True}
    elif not kitti:
        cfg["evaluation"] = {"accept_inverted_angle": True}
    return cfg


SHIPPED = ("SynthCars", "SynthCarsPersons", "Synth20", "Synth20others", "KITTITrackletsCars",
           "KITTITrackletsCarsPersons", "KITTITrackletsCarsHard", "KITTITrackletsCarsPersonsHard")

configGlobal = NameSpace()


def reset_config() -> None:
    configGlobal.reset()
    dump_to_namespace(configGlobal, default_config())


reset_config()


def _finish(name: str) -> NameSpace:
    configGlobal.__dict__["name"] = name
    configGlobal.data.__dict__["basename"] = os.path.basename(configGlobal.data.basepath)
    configGlobal.logging.__dict__["logdir"] = configGlobal.logging.basedir + f"/{name}"
    if configGlobal.evaluation.has("special") and configGlobal.evaluation.special.mode == "icp":
        configGlobal.logging.__dict__["logdir"] = configGlobal.logging.basedir + f"/icp_{configGlobal.data.basename}/{name}"
    for split in ("train", "val"):
        path = f"{configGlobal.data.basepath}/split/{split}.txt"
        n = 0
        if os.path.isfile(path):
            with open(path) as fh:
                n = sum(1 for line in fh if line.strip())
        configGlobal.data.__dict__["n" + split] = n
    return configGlobal


def load_config(filename: str) -> NameSpace:
    """Reference config.py:66-82: overlay a run JSON (the reference's own files work unchanged)."""
    assert filename.endswith(".json")
    with open(filename) as handle:
        dump_to_namespace(configGlobal, json.load(handle))
    return _finish(os.path.basename(filename)[:-5])


def load_shipped(name: str) -> NameSpace:
    reset_config()
    dump_to_namespace(configGlobal, shipped_config(name))
    return _finish(name)


def save_config(filename: str) -> None:
    assert filename.endswith(".json")
    with open(filename, "w") as handle:
        json.dump(namespace_to_dict(configGlobal), handle)


def validate(cfg: NameSpace) -> None:
    """Reject (error, never fall back) what this engine does not implement."""
    if cfg.model.model != "tp8":
        raise ValueError(f"model.model={cfg.model.model!r}: only 'tp8' exists (reference train.py:52-55)")
    if cfg.model.backbone != "pointnet":
        raise ValueError(f"model.backbone={cfg.model.backbone!r} is not implemented (only 'pointnet')")
    if cfg.training.loss.loss != "separate":
        raise ValueError(f"training.loss.loss={cfg.training.loss.loss!r} is not implemented (only 'separate')")
    if cfg.training.loss.has("options") and cfg.training.loss.options.has("soft_angle_classes") \
            and cfg.training.loss.options.soft_angle_classes:
        raise ValueError("training.loss.options.soft_angle_classes=true is not implemented")
    if cfg.data.num_channels != 3:
        raise ValueError("data.num_channels must be 3")
    if cfg.training.has("optimizer") and cfg.training.optimizer.has("optimizer"):
        name = cfg.training.optimizer.optimizer
        if name not in ("adam", "momentum"):
            raise ValueError(f"training.optimizer.optimizer={name!r}: Invalid optimizer (train.py:215)")
        if name == "momentum" and not cfg.training.optimizer.has("momentum"):
            raise ValueError("training.optimizer.optimizer='momentum' needs training.optimizer.momentum (train.py:212; "
                             "configs/default.json does not define it)")
    if cfg.evaluation.has("special"):
        mode = cfg.evaluation.special.mode
        if mode == "icp":                                        # train.py:548-551 -> icp.py:150
            sp = cfg.evaluation.special.icp if cfg.evaluation.special.has("icp") else None
            if sp is None or not sp.has("variant") or sp.variant != "p2point" or not sp.has("with_constraint") \
                    or not sp.with_constraint or sp.has("refine"):
                raise ValueError("evaluation.special.mode='icp': only icp.variant='p2point' with icp.with_constraint=true and "
                                 "no icp.refine is implemented (the yaw-constrained point-to-point ICP of row N4, the "
                                 "reference's icp_<dataset>_o3_p2p.json configs); the global-registration variants of the "
                                 "Open3D fork (icp.py:84-143) are outside the hot path")
        elif mode != "timings":
            raise ValueError(f"evaluation.special.mode={mode!r} is not implemented: 'held' (evaluate_held, evaluation.py:49) "
                             "is outside the hot path; 'timings' (train.py:553-559) and 'icp' (variant p2point) are")


def optimizer_from_config(cfg: NameSpace):
    """train.py:211-216 -> (name, momentum) for Engine.set_optimizer."""
    validate(cfg)
    if cfg.training.has("optimizer") and cfg.training.optimizer.has("optimizer") \
            and cfg.training.optimizer.optimizer == "momentum":
        return "momentum", float(cfg.training.optimizer.momentum)
    return "adam", None


def arch_from_config(cfg: NameSpace) -> "_lib.Arch":
    """models/tp8.py:98,108,115,130,154,307-308,318 -> the C-ABI an3d_arch struct."""
    validate(cfg)
    o = cfg.model.options
    a = _lib.Arch()
    a.num_bins = int(cfg.model.angles.num_bins)
    a.accept_inverted_angle = int(bool(cfg.model.angles.accept_inverted_angle))
    a.angle_factor = float(o.angle_factor)
    a.early_stage_factor = float(o.early_stage_factor)
    convs = [o.s1transformer[0], o.s2transformer[0], o.embedding]
    fcs = [o.s1transformer[1], o.s2transformer[1], o.remaining_transform_prediction]
    for s in range(3):
        if len(convs[s]) > _lib.MAX_LAYERS or len(fcs[s][0]) > _lib.MAX_LAYERS:
            raise ValueError(f"at most {_lib.MAX_LAYERS} layers per stack")
        a.n_conv[s] = len(convs[s])
        for i, c in enumerate(convs[s]):
            a.conv[s][i] = int(c)
        a.n_fc[s] = len(fcs[s][0])
        for i, c in enumerate(fcs[s][0]):
            a.fc[s][i] = int(c)
        keep = fcs[s][1]
        a.keep_prob[s] = 1.0 if keep is None else float(keep)
    return a
