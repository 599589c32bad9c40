"""Architecture description, TF variable naming and initialisation for the oracle.

Follows (file:line into /root/reference):
  * layer lists          configs/*.json `model.options.*`; models/tp8.py:97-98,108,115,130,154
  * variable scopes      models/tp8.py:52,62-66,77,92,140-143,154
  * Xavier-uniform init  utils/tf_util.py:41-45 (tf.contrib.layers.xavier_initializer, fans
                         include the kernel window), biases 0 (tf_util.py:159,338), gamma=1 /
                         beta=0 (tf_util.py:470-473), EMA shadows zero-initialised (tf_util.py:476).

Oracle = test infrastructure (see oracle/__init__.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

BN_EPS = 1e-3  # utils/tf_util.py:491


@dataclass
class Arch:
    """Shipped architecture (configs/SynthCars.json:8-20) by default."""
    num_bins: int = 50
    s1_conv: Tuple[int, ...] = (64, 128, 256)
    s1_fc: Tuple[int, ...] = (512, 256)
    s1_keep: float = 0.7
    s2_conv: Tuple[int, ...] = (64, 128, 512)
    s2_fc: Tuple[int, ...] = (512, 256)
    s2_keep: float = 0.7
    emb_conv: Tuple[int, ...] = (64, 128, 1024)
    head_fc: Tuple[int, ...] = (512, 256)
    head_keep: float = 0.7
    angle_factor: float = 1.0
    early_stage_factor: float = 0.5
    accept_inverted_angle: bool = True

    @property
    def out_s1(self) -> int:
        return 3

    @property
    def out_s2(self) -> int:
        return 3 + 2 * self.num_bins

    @property
    def out_head(self) -> int:
        return 3 + 2 * self.num_bins


def tiny_arch(**kw) -> Arch:
    """Small architecture for fast CPU tests (same topology, narrow layers)."""
    base = dict(num_bins=6, s1_conv=(8, 16, 24), s1_fc=(16, 8), s2_conv=(8, 16, 32), s2_fc=(16, 8),
                emb_conv=(8, 16, 40), head_fc=(16, 8))
    base.update(kw)
    return Arch(**base)


def default_arch(**kw) -> Arch:
    """The architecture of the reference's configs/default.json:8-22: [128,128,256] for stage 1, five-layer conv stacks
    for stage 2 and the embedding (models/tp8.py:49-59 builds any depth), 36 bins, no inverted-angle acceptance."""
    base = dict(num_bins=36, s1_conv=(128, 128, 256), s1_fc=(512, 256), s1_keep=0.7, s2_conv=(64, 64, 64, 128, 1024),
                s2_fc=(512, 256), s2_keep=0.7, emb_conv=(64, 64, 64, 128, 1024), head_fc=(512, 256), head_keep=0.7,
                angle_factor=1.0, early_stage_factor=0.1, accept_inverted_angle=False)
    base.update(kw)
    return Arch(**base)


@dataclass
class LinearSpec:
    scope: str          # TF scope below the branch prefix, e.g. "transformer1/embedding/conv1"
    cin: int
    cout: int
    bn: bool
    tf_shape: Tuple[int, ...]
    shared_prefix: str  # "siamese/" for siamese layers, "" for the head


def stage_specs(arch: Arch) -> Dict[str, List[LinearSpec]]:
    """Linear layers per stage in execution order.  Scope strings per SURVEY App. C."""
    out: Dict[str, List[LinearSpec]] = {}

    def convs(prefix: str, sizes) -> List[LinearSpec]:
        specs, cin = [], 3
        for i, c in enumerate(sizes):
            # first conv kernel is [1, num_channel] over the xyz axis with C_in = 1 (tp8.py:55)
            tf_shape = (1, 3, 1, c) if i == 0 else (1, 1, cin, c)
            specs.append(LinearSpec(f"{prefix}/conv{i + 1}", cin, c, True, tf_shape, "siamese/"))
            cin = c
        return specs

    def fcs(prefix: str, cin: int, hidden, cout: int, shared_prefix: str) -> List[LinearSpec]:
        specs = []
        for i, c in enumerate(hidden):
            specs.append(LinearSpec(f"{prefix}fc{i + 1}", cin, c, True, (cin, c), shared_prefix))
            cin = c
        specs.append(LinearSpec(f"{prefix}fc{len(hidden) + 1}", cin, cout, False, (cin, cout), shared_prefix))
        return specs

    out["s1_conv"] = convs("transformer1/embedding", arch.s1_conv)
    out["s1_fc"] = fcs("transformer1/mlp/", arch.s1_conv[-1], arch.s1_fc, arch.out_s1, "siamese/")
    out["s2_conv"] = convs("transformer2/embedding", arch.s2_conv)
    out["s2_fc"] = fcs("transformer2/mlp/", arch.s2_conv[-1], arch.s2_fc, arch.out_s2, "siamese/")
    # get_backbone ignores its scope_name (tp8.py:62-66, Q2): final embedding lives in "siamese/embedding"
    out["emb_conv"] = convs("embedding", arch.emb_conv)
    out["head_fc"] = fcs("", 2 * arch.emb_conv[-1], arch.head_fc, arch.out_head, "")
    return out


STAGE_ORDER = ("s1_conv", "s1_fc", "s2_conv", "s2_fc", "emb_conv", "head_fc")


def branch_prefix(branch: int) -> str:
    """BN variables are tf.Variable, so the second siamese pass gets scope 'siamese_1' (Q0)."""
    return "siamese/" if branch == 0 else "siamese_1/"


def bn_names(spec: LinearSpec, branch: int) -> Dict[str, str]:
    p = (branch_prefix(branch) if spec.shared_prefix else "") + spec.scope + "/bn/"
    return {
        "gamma": p + "gamma",
        "beta": p + "beta",
        "moving_mean": p + "moments/Squeeze/ExponentialMovingAverage",
        "moving_var": p + "moments/Squeeze_1/ExponentialMovingAverage",
    }


def weight_names(spec: LinearSpec) -> Dict[str, str]:
    p = spec.shared_prefix + spec.scope + "/"
    return {"weights": p + "weights", "biases": p + "biases"}


def trainable_specs(arch: Arch) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, matrix-shape) of every trainable tensor, in the flat-buffer order the engine uses:
    shared weights/biases stage by stage, then BN gamma/beta for branch 0, branch 1, head."""
    st = stage_specs(arch)
    out: List[Tuple[str, Tuple[int, ...]]] = []
    for key in STAGE_ORDER:
        for s in st[key]:
            n = weight_names(s)
            out.append((n["weights"], (s.cin, s.cout)))
            out.append((n["biases"], (s.cout,)))
    for branch in (0, 1):
        for key in STAGE_ORDER[:-1]:
            for s in st[key]:
                if s.bn:
                    n = bn_names(s, branch)
                    out.append((n["gamma"], (s.cout,)))
                    out.append((n["beta"], (s.cout,)))
    for s in st["head_fc"]:
        if s.bn:
            n = bn_names(s, 0)
            out.append((n["gamma"], (s.cout,)))
            out.append((n["beta"], (s.cout,)))
    return out


def state_specs(arch: Arch) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of the non-trainable BN shadow variables, same ordering as gamma/beta."""
    st = stage_specs(arch)
    out: List[Tuple[str, Tuple[int, ...]]] = []
    for branch in (0, 1):
        for key in STAGE_ORDER[:-1]:
            for s in st[key]:
                if s.bn:
                    n = bn_names(s, branch)
                    out.append((n["moving_mean"], (s.cout,)))
                    out.append((n["moving_var"], (s.cout,)))
    for s in st["head_fc"]:
        if s.bn:
            n = bn_names(s, 0)
            out.append((n["moving_mean"], (s.cout,)))
            out.append((n["moving_var"], (s.cout,)))
    return out


def num_trainable(arch: Arch) -> int:
    return int(sum(int(np.prod(s)) for _, s in trainable_specs(arch)))


def init_params(arch: Arch, seed: int = 0) -> Dict[str, np.ndarray]:
    """Xavier-uniform weights (limit = sqrt(6/(fan_in+fan_out)), fans include the kernel window:
    first conv [1,3,1,C] has fan_in 3, fan_out 3*C), zero biases, gamma 1, beta 0."""
    rng = np.random.Generator(np.random.PCG64(seed))
    st = stage_specs(arch)
    params: Dict[str, np.ndarray] = {}
    for key in STAGE_ORDER:
        for s in st[key]:
            receptive = int(np.prod(s.tf_shape[:-2])) if len(s.tf_shape) == 4 else 1
            fan_in = s.tf_shape[-2] * receptive
            fan_out = s.tf_shape[-1] * receptive
            limit = np.sqrt(6.0 / (fan_in + fan_out))
            n = weight_names(s)
            params[n["weights"]] = rng.uniform(-limit, limit, size=(s.cin, s.cout)).astype(np.float32)
            params[n["biases"]] = np.zeros((s.cout,), np.float32)
    for name, shape in trainable_specs(arch):
        if name.endswith("/gamma"):
            params[name] = np.ones(shape, np.float32)
        elif name.endswith("/beta"):
            params[name] = np.zeros(shape, np.float32)
    return params


def init_state(arch: Arch) -> Dict[str, np.ndarray]:
    """EMA shadows start at zero (Q7)."""
    return {name: np.zeros(shape, np.float32) for name, shape in state_specs(arch)}


def randomize_for_test(arch: Arch, params: Dict[str, np.ndarray], state: Dict[str, np.ndarray], seed: int = 1):
    """Perturb biases / gamma / beta / shadows so that tests exercise every term (at init the
    biases and betas are zero and every gamma is one, which hides indexing mistakes)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    for k, v in params.items():
        if k.endswith("/biases"):
            params[k] = rng.normal(0, 0.1, v.shape).astype(np.float32)
        elif k.endswith("/gamma"):
            g = rng.uniform(0.5, 1.5, v.shape).astype(np.float32)
            g[rng.uniform(size=v.shape) < 0.15] *= -1.0   # negative scales exercise the min-pool path
            params[k] = g
        elif k.endswith("/beta"):
            params[k] = rng.normal(0, 0.2, v.shape).astype(np.float32)
    for k, v in state.items():
        if "Squeeze_1" in k:
            state[k] = rng.uniform(0.5, 1.5, v.shape).astype(np.float32)
        else:
            state[k] = rng.normal(0, 0.3, v.shape).astype(np.float32)
    return params, state
