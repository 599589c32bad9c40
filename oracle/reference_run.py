"""Executes the reference's own, unmodified model code (`/root/reference/models/tp8.py`,
`utils/tf_util.py`, `config.py`, `utils/eulerangles.py`, `tp_utils/pointcloud.py`) on the TF1 shim of
`oracle/tf1_shim` -- TEST INFRASTRUCTURE ONLY.

Only usable where /root/reference exists (the build container); it generates the committed fixtures
`tests/golden/reference_*.npz` (`tests/golden/make_reference_golden.py`) and backs the live
cross-checks of `tests/test_reference_run.py`.  Nothing on the product path, in the `-m gpu` tests,
in `smoke()` or in `bench.py` imports this module.

What comes from the reference here: the whole graph of a1-a16/a21 (scopes and variable sharing,
layer order, bias-before-BN, the pooling window, centre/angle chaining, slicing of the 103 outputs,
the in-graph decode, every loss term and its `[B,B]` broadcasts, the inverted-angle selection) and the
host decode `classLogits2angle`.  What is restated: the TensorFlow primitives (see the shim header).
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types
from typing import Dict, Optional

import numpy as np

REFERENCE_ROOT = os.environ.get("AN3D_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf1_shim")

MASK_ORDER = ("s1_b0", "s2_b0", "s1_b1", "s2_b1", "head")   # graph-construction order of tf.nn.dropout calls


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "tp8.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Returns (tf shim, reference tp8 module, reference config module)."""
    if "tp8" in _loaded:
        return _loaded["tf"], _loaded["tp8"], _loaded["config"]
    assert available(), f"{REFERENCE_ROOT} not present"
    for p in (_SHIM, REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "models"), os.path.join(REFERENCE_ROOT, "utils")):
        if p not in sys.path:
            sys.path.insert(0, p)
    # config.py imports provider -> pointcloud (open3d, pyntcloud, ... absent).  Neither is on the tp8 path:
    # config only calls provider.getDataFiles inside load_config(), which is not used here.
    if "provider" not in sys.modules:
        _stub("provider", getDataFiles=lambda *_: [])
    import tensorflow as tf   # the shim
    assert tf.__version__.endswith("shim"), "a real tensorflow shadowed the shim"
    config = importlib.import_module("config")
    tp8 = importlib.import_module("tp8")
    _loaded.update(tf=tf, tp8=tp8, config=config)
    return tf, tp8, config


def configure(config_name: str = "SynthCars", overrides: Optional[dict] = None):
    """cfg = default.json + configs/<name>.json (as config.load_config does, minus the data-split lookup)."""
    _, _, config = load()
    config.reset_config()
    with open(os.path.join(REFERENCE_ROOT, "configs", config_name + ".json")) as fh:
        config.dump_to_namespace(config.configGlobal, json.load(fh))
    if overrides:
        config.dump_to_namespace(config.configGlobal, overrides)
    return config.configGlobal


def arch_overrides(arch, loss: str = "separate") -> dict:
    """oracle.arch.Arch -> the config keys models/tp8.py reads."""
    return {"model": {"backbone": "pointnet",
                      "options": {"angle_factor": arch.angle_factor, "early_stage_factor": arch.early_stage_factor,
                                  "s1transformer": [list(arch.s1_conv), [list(arch.s1_fc), arch.s1_keep]],
                                  "s2transformer": [list(arch.s2_conv), [list(arch.s2_fc), arch.s2_keep]],
                                  "embedding": list(arch.emb_conv),
                                  "remaining_transform_prediction": [list(arch.head_fc), arch.head_keep]},
                      "angles": {"num_bins": arch.num_bins, "accept_inverted_angle": arch.accept_inverted_angle}},
            "training": {"loss": {"loss": loss, "options": {"soft_angle_classes": False}}}}


def run(batch: Dict[str, np.ndarray], arch, params: Dict[str, np.ndarray], state: Dict[str, np.ndarray],
        is_training: bool, bn_decay: Optional[float] = None, masks: Optional[Dict[str, np.ndarray]] = None,
        double: bool = False, with_loss: bool = True, with_grads: bool = False, loss: str = "separate"):
    """One evaluation of the reference graph.  Returns dict(end_points, loss, grads, new_state, var_names, ...)."""
    import torch
    tf, tp8, _ = load()
    configure("SynthCars", arch_overrides(arch, loss))
    tf.reset()
    dt = torch.float64 if double else torch.float32
    tf.set_float_dtype(dt)
    try:
        tf.preload(params)
        tf.preload(state)
        tf.queue_dropout_masks([masks[k] for k in MASK_ORDER] if masks else [])
        as_t = lambda a: torch.as_tensor(np.asarray(a), dtype=dt).as_subclass(tf.Tensor)  # noqa: E731
        feeds = {k: as_t(v) for k, v in batch.items()}
        decay = None if bn_decay is None else as_t(bn_decay)
        ep = tp8.get_model(feeds["pcs1"], feeds["pcs2"], torch.tensor(bool(is_training)), bn_decay=decay)
        out = {"end_points": {k: v.detach().numpy().copy() for k, v in ep.items()},
               "var_names": list(tf.variables().keys()),
               "var_shapes": {k: tuple(v.shape) for k, v in tf.variables().items()},
               "trainable": tf.trainable_variables(),
               "new_state": {k: v.detach().numpy().copy() for k, v in tf.shadow_variables().items()}}
        if with_loss:
            loss = tp8.get_loss(feeds["pcs1"], feeds["pcs2"], feeds["translations"], feeds["rel_angles"],
                                feeds["pc1_centers"], feeds["pc2_centers"], feeds["pc1_angles"], feeds["pc2_angles"], ep)
            out["loss"] = float(loss.detach())
            if with_grads:
                names = tf.trainable_variables()
                vs = tf.variables()
                gs = torch.autograd.grad(loss, [vs[n] for n in names], allow_unused=True)
                out["grads"] = {n: (np.zeros(tuple(vs[n].shape), np.float64 if double else np.float32) if g is None
                                    else g.detach().numpy().copy()) for n, g in zip(names, gs)}
        dec = lambda k: tp8.classLogits2angle(out["end_points"][k])  # noqa: E731  (host decode, train.py:453-455)
        out["pred_angles"] = dec("pred_pc2angle_logits") - dec("pred_pc1angle_logits") + dec("pred_remaining_angle_logits")
        return out
    finally:
        tf.set_float_dtype(torch.float32)


# ------------------------------------------------------------------------------------------------
# rigid-transform helpers of the reference (a17-a20)
# ------------------------------------------------------------------------------------------------
def load_pointcloud_module():
    """tp_utils/pointcloud.py with its visualisation / mesh dependencies stubbed (none is used by
    get_mat_angle, transform_points, translate_transform_to_new_center_of_rotation)."""
    if "pointcloud_ref" in _loaded:
        return _loaded["pointcloud_ref"]
    assert available()
    from unittest import mock
    from scipy.spatial.transform import Rotation
    if not hasattr(Rotation, "as_dcm"):        # pointcloud.py:288 uses the pre-1.4 scipy name of as_matrix
        Rotation.as_dcm = Rotation.as_matrix
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    import tensorflow as tf   # noqa: F401  (the shim; pointcloud.py:22 imports tensorflow.python.util.nest)
    if "tensorflow.python.util" not in sys.modules:
        nest = types.SimpleNamespace(is_sequence=lambda x: isinstance(x, (list, tuple)))
        _stub("tensorflow.python", util=None)
        _stub("tensorflow.python.util", nest=nest)
    for name in ("PIL", "PIL.Image", "open3d", "quaternion", "pyntcloud", "trimesh", "pythreejs", "IPython", "IPython.display", "ipywidgets",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "pandas", "tqdm", "cv2", "transforms3d",
                 "transforms3d.euler", "seaborn", "numba"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = mock.MagicMock(name=name)
    p = os.path.join(REFERENCE_ROOT, "tp_utils")
    if p not in sys.path:
        sys.path.insert(0, p)
    spec = importlib.util.spec_from_file_location("pointcloud_ref", os.path.join(p, "pointcloud.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _loaded["pointcloud_ref"] = mod
    return mod


def load_eulerangles_module():
    spec = importlib.util.spec_from_file_location("eulerangles_ref", os.path.join(REFERENCE_ROOT, "utils", "eulerangles.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
