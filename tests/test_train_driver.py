"""The drop-in driver `alignnet_b200.train` (reference train.py call surface) end to end on the tiny
dataset in the reference's on-disk format: CLI, config overlay, provider, schedules, optimiser steps, evaluation,
file names of the run directory, checkpoint / eval_only round trip."""
import json
import os
import shutil

import numpy as np
import pytest

from helpers import GOLDEN


def _make_run(tmp_path):
    base = tmp_path / "SynthTiny"
    shutil.copytree(os.path.join(GOLDEN, "dataset_tiny"), base)
    shutil.copy(base / "split" / "val.txt", base / "split" / "train.txt")
    cfg = {"data": {"basepath": str(base)}, "model": {"num_points": 32},
           "training": {"batch_size": 2, "num_epochs": 2, "learning_rate": 0.002},
           "logging": {"basedir": str(tmp_path / "logs")}}
    path = tmp_path / "TinyRun.json"
    path.write_text(json.dumps(cfg))
    return str(path), tmp_path / "logs" / "TinyRun"


def test_cli_matches_reference_flags():
    from alignnet_b200 import train
    f = train.parse_args(["eval_only", "--config", "x.json", "--eval_epoch", "7", "--its", "5"])
    assert f.operation == "eval_only" and f.config == "x.json" and f.eval_epoch == "7" and not f.refineICP
    with pytest.raises(SystemExit):
        train.parse_args(["fit", "--config", "x.json"])
    with pytest.raises(SystemExit):
        train.parse_args(["train"])


@pytest.mark.gpu
def test_train_then_eval_only_round_trip(tmp_path):
    import __graft_entry__ as ge
    ge.build()
    from alignnet_b200 import config as C, train
    cfg_path, logdir = _make_run(tmp_path)
    C.reset_config()
    np.random.seed(3)
    last = train.main(["train", "--config", cfg_path, "--precision", "fp32"])
    for rel in ("config.json", "out.log", "model.ckpt.index", "model.ckpt.data-00000-of-00001", "model-0.index", "model-1.index", "val/eval000001/eval.json",
                "val/eval000001/eval_180.json", "val/eval000001/pred_translations.npy", "val/eval000001/pred_angles.npy",
                "val/eval000001/pred_s2_pc1centers.npy"):
        assert (logdir / rel).exists(), rel
    d = json.load(open(logdir / "val/eval000001/eval.json"))
    assert d["num"] == 6 and set(d) >= {"corr_levels", "eval_5m", "val", "test", "reg_eval", "mean_time"}
    assert last["eval"]["num"] == 6
    from alignnet_b200 import tf_checkpoint
    ck = tf_checkpoint.read_checkpoint(str(logdir / "model-1"))
    w3 = ck["siamese/embedding/conv3/weights"]
    assert int(ck["Variable"]) == 6 and w3.ndim == 4 and w3.shape[:2] == (1, 1)          # TF kernel shape [1, 1, Cin, Cout]
    assert ck["siamese/embedding/conv1/weights"].shape[:3] == (1, 3, 1)
    # eval_only restores the checkpoint and reproduces the predictions (same resampling draws with the same seed)
    pred_a = np.load(logdir / "val/eval000001/pred_translations.npy")
    C.reset_config()
    np.random.seed(11)
    train.main(["eval_only", "--config", cfg_path, "--eval_epoch", "1", "--precision", "fp32"])
    first = np.load(logdir / "val/eval000001/pred_translations.npy")
    C.reset_config()
    np.random.seed(11)
    train.main(["eval_only", "--config", cfg_path, "--eval_epoch", "1", "--precision", "fp32"])
    second = np.load(logdir / "val/eval000001/pred_translations.npy")
    np.testing.assert_allclose(first, second, atol=1e-5)
    assert np.isfinite(pred_a).all() and pred_a.shape == (6, 3)
    with pytest.raises(ValueError):
        train.main(["train", "--config", cfg_path, "--refineICP"])
    # eval_only --refineICP (row N4): results go to the refined_p2p directory, rotation centres are the origin
    C.reset_config()
    np.random.seed(11)
    ref = train.main(["eval_only", "--config", cfg_path, "--eval_epoch", "1", "--precision", "fp32", "--refineICP"])
    rdir = logdir / "val/eval000001/refined_p2p"
    assert (rdir / "eval.json").exists() and ref["eval"]["num"] == 6
    assert not np.load(rdir / "pred_s2_pc1centers.npy").any()
    assert np.isfinite(np.load(rdir / "pred_translations.npy")).all() and np.isfinite(np.load(rdir / "pred_angles.npy")).all()
