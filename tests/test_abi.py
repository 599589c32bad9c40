"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/alignnet_b200.h declares, reports the flat layout the oracle expects (TF variable names),
validates its inputs, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import arch as A
from helpers import engine_arch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from alignnet_b200 import _lib
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    from alignnet_b200 import _lib
    header = open(os.path.join(ROOT, "include", "alignnet_b200.h")).read()
    declared = set(re.findall(r"\b(an3d_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.an3d_version() == 100


@pytest.mark.parametrize("arch", [A.Arch(), A.tiny_arch(), A.Arch(s2_conv=(64, 64, 64, 128, 1024), num_bins=36)])
def test_layout_matches_oracle_names(lib, arch):
    from alignnet_b200 import engine
    e = engine.Engine(engine_arch(arch), allocate=False)
    spec = A.trainable_specs(arch)
    assert e.params_layout.order == [n for n, _ in spec]
    off = 0
    for name, shape in spec:
        o, shp = e.params_layout.entries[name]
        # same order and sizes as the reference graph's variables; starts are 16-byte aligned, never overlapping
        assert o >= off and o % 4 == 0 and o - off < 4 and int(np.prod(shp)) == int(np.prod(shape)), name
        off = o + int(np.prod(shape))
    assert off <= e.params_layout.total < off + 4
    assert sum(int(np.prod(s)) for _, s in e.params_layout.entries.values()) == A.num_trainable(arch)
    sspec = A.state_specs(arch)
    assert e.state_layout.order == [n for n, _ in sspec]
    assert e.state_layout.total == sum(int(np.prod(s)) for _, s in sspec)


def test_shipped_layout_size(lib):
    from alignnet_b200 import engine
    e = engine.Engine(engine.shipped_arch(), allocate=False)
    sizes = sum(int(np.prod(shp)) for _, shp in e.params_layout.entries.values())
    assert sizes == 2165073
    assert sizes <= e.params_layout.total < sizes + 4 * len(e.params_layout.entries)
    assert e.state_layout.total == 2 * 8576
    # first conv kernel keeps the reference's [1,3,1,64] shape (models/tp8.py:55)
    assert e.params_layout.entries["siamese/transformer1/embedding/conv1/weights"][1] == (1, 3, 1, 64)
    assert "siamese_1/embedding/conv3/bn/gamma" in e.params_layout.entries   # per-branch BN (Q0)
    assert "fc3/weights" in e.params_layout.entries                           # head scope '' (tp8.py:154)


def test_workspace_bytes_and_validation(lib):
    from alignnet_b200 import engine, _lib
    e = engine.Engine(engine.shipped_arch(), allocate=False)
    small = e.workspace_bytes(4, 16, 0)
    train = e.workspace_bytes(4, 16, _lib.TRAINING)
    big = e.workspace_bytes(32, 200, _lib.TRAINING)
    assert 0 < small < train < big
    with pytest.raises(_lib.An3dError):
        e.workspace_bytes(0, 16, 0)
    bad = engine.shipped_arch()
    bad.num_bins = 1
    with pytest.raises(_lib.An3dError):
        engine.Engine(bad, allocate=False)
    bad = engine.shipped_arch()
    bad.keep_prob[0] = 0.0
    with pytest.raises(_lib.An3dError):
        engine.Engine(bad, allocate=False)


def test_no_cpu_fallback(lib):
    """Without a CUDA device every compute entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from alignnet_b200 import _lib, engine
    buf = (C.c_float * 64)()
    rc = lib.an3d_adam_step(buf, buf, buf, buf, 64, 0.1, 1, 1.0, 0.9, 0.999, 1e-8, None)
    assert rc in (-3, -5) and lib.an3d_last_error()
    rc = lib.an3d_rigid_apply(buf, None, None, None, buf, 1, 4, None)
    assert rc in (-3, -5)
    with pytest.raises(RuntimeError):
        engine.Engine(engine.shipped_arch())


def test_config_schema_roundtrip(tmp_path):
    from alignnet_b200 import config
    cfg = config.load_shipped("SynthCars")
    assert cfg.model.num_points == 512 and cfg.model.angles.num_bins == 50
    assert cfg.training.loss.options.soft_angle_classes is False      # inherited from the defaults
    assert cfg.logging.logdir.endswith("/SynthCars") and cfg.name == "SynthCars"
    a = config.arch_from_config(cfg)
    assert list(a.conv[2])[:3] == [64, 128, 1024] and a.n_fc[2] == 2 and abs(a.keep_prob[0] - 0.7) < 1e-7
    # a reference-style JSON overlay loads unchanged
    p = tmp_path / "MyRun.json"
    p.write_text('{"model": {"num_points": 200, "angles": {"num_bins": 36}}, "training": {"batch_size": 32}}')
    config.reset_config()
    cfg = config.load_config(str(p))
    assert cfg.name == "MyRun" and cfg.model.num_points == 200 and cfg.training.batch_size == 32
    assert cfg.model.options.embedding == [64, 64, 64, 128, 1024]     # default.json value kept
    cfg.model.__dict__["backbone"] = "dgcnn"
    with pytest.raises(ValueError):
        config.validate(cfg)
    config.reset_config()
