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
    for k in ("pred_translations", "pred_angles", "pred_s2_pc1angles", "pred_s2_pc2angles", "pred_s1_pc2centers"):
        a = np.load(logdir / f"val/eval000001/{k}.npy")          # train.py:408-419,534-543: float32, [n,3] / [n,1]
        assert a.dtype == np.float32 and a.shape[0] == 6, (k, a.dtype, a.shape)
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

    # eval_only asserts that the requested epoch's checkpoint exists (train.py:251) instead of falling back
    C.reset_config()
    with pytest.raises(FileNotFoundError):
        train.main(["eval_only", "--config", cfg_path, "--eval_epoch", "7", "--precision", "fp32"])
    # --use_old_results (train.py:421-424,464-465): ICP seeded from the stored predictions, no checkpoint needed
    C.reset_config()
    np.random.seed(11)
    old = train.main(["eval_only", "--config", cfg_path, "--eval_epoch", "1", "--precision", "fp32", "--refineICP", "--use_old_results"])
    assert old["eval"]["num"] == 6

    # pre-training restore (train.py:276-293): all variables but the global step, then an initial evaluation
    cfg = json.load(open(cfg_path))
    cfg["training"]["pretraining"] = {"model": str(logdir / "model-1")}
    cfg["training"]["num_epochs"] = 1
    pre_path = tmp_path / "TinyPre.json"
    pre_path.write_text(json.dumps(cfg))
    C.reset_config()
    np.random.seed(5)
    train.main(["eval_only", "--config", cfg_path, "--eval_epoch", "1", "--precision", "fp32"])
    want = np.load(logdir / "val/eval000001/pred_translations.npy")      # model-1's predictions under this seed's resampling draws
    C.reset_config()
    np.random.seed(5)
    train.main(["train", "--config", str(pre_path), "--precision", "fp32"])
    pre_dir = tmp_path / "logs" / "TinyPre"
    assert (pre_dir / "val/eval0pretr/eval.json").exists()                                 # 'pretr'.zfill(6)
    first_eval = np.load(pre_dir / "val/eval0pretr/pred_translations.npy")
    np.testing.assert_allclose(first_eval, want, atol=1e-4)                                # the restored weights, not a fresh init
    assert int(tf_checkpoint.read_checkpoint(str(pre_dir / "model-0"))["Variable"]) == 3   # the step restarted at 0
    cfg["training"]["pretraining"] = {"model": str(tmp_path / "nope")}
    pre_path.write_text(json.dumps(cfg))
    C.reset_config()
    shutil.rmtree(pre_dir)
    with pytest.raises(FileNotFoundError):
        train.main(["train", "--config", str(pre_path), "--precision", "fp32"])

    # the reference's own benchmark harness (train.py:553-559): batch 32, no restore, ten passes, seconds per pair
    cfg = json.load(open(cfg_path))
    cfg["evaluation"] = {"special": {"mode": "timings"}}
    t_path = tmp_path / "TinyTimings.json"
    t_path.write_text(json.dumps(cfg))
    C.reset_config()
    res = train.main(["eval_only", "--config", str(t_path), "--eval_epoch", "1", "--precision", "fp32"])   # (epoch loop: train.py:297)
    assert res["mean_time"] > 0
    # the same harness on the tensor cores: this run keeps default.json's five-layer conv stacks, which the bf16 mode runs
    # layer by layer (csrc/gemm_tc.cuh)
    C.reset_config()
    res16 = train.main(["eval_only", "--config", str(t_path), "--eval_epoch", "1", "--precision", "bf16"])
    assert res16["mean_time"] > 0
    # what the engine does not implement is rejected, not ignored
    for bad in ({"evaluation": {"special": {"mode": "icp"}}}, {"evaluation": {"special": {"mode": "held"}}},
                {"evaluation": {"special": {"mode": "icp", "icp": {"variant": "o3_gicp", "with_constraint": True}}}},
                {"training": {"optimizer": {"optimizer": "momentum"}}}, {"training": {"optimizer": {"optimizer": "sgd"}}}):
        cfg = json.load(open(cfg_path))
        for k, v in bad.items():
            cfg.setdefault(k, {}).update(v)
        b_path = tmp_path / "TinyBad.json"
        b_path.write_text(json.dumps(cfg))
        C.reset_config()
        with pytest.raises(ValueError):
            train.main(["train", "--config", str(b_path), "--precision", "fp32"])
    C.reset_config()


@pytest.mark.gpu
def test_momentum_optimizer_run(tmp_path):
    """`training.optimizer.optimizer == 'momentum'` (train.py:211-212): the driver trains with the momentum update, the
    checkpoint carries `<var>/Momentum` slots (no Adam slots) and resuming restores them."""
    import __graft_entry__ as ge
    ge.build()
    from alignnet_b200 import config as C, tf_checkpoint, train
    cfg_path, logdir = _make_run(tmp_path)
    cfg = json.load(open(cfg_path))
    cfg["training"]["optimizer"] = {"optimizer": "momentum", "momentum": 0.9}
    cfg["training"]["learning_rate"] = 1e-4
    open(cfg_path, "w").write(json.dumps(cfg))
    C.reset_config()
    np.random.seed(3)
    last = train.main(["train", "--config", cfg_path, "--precision", "fp32"])
    assert last["eval"]["num"] == 6
    ck = tf_checkpoint.read_checkpoint(str(logdir / "model-1"))
    assert int(ck["Variable"]) == 6
    slots = [k for k in ck if k.endswith("/Momentum")]
    assert "siamese/embedding/conv3/weights/Momentum" in slots and len(slots) == len([k for k in ck if k + "/Momentum" in ck])
    assert not any(k.endswith(("/Adam", "/Adam_1")) for k in ck) and "beta1_power" not in ck
    acc = ck["siamese/embedding/conv3/weights/Momentum"]
    assert acc.shape == ck["siamese/embedding/conv3/weights"].shape and np.isfinite(acc).all() and np.abs(acc).max() > 0
    # resume (train.py:267-275): one more epoch from model.ckpt continues at step 6 with the stored accumulators
    cfg["training"]["num_epochs"] = 3
    open(cfg_path, "w").write(json.dumps(cfg))
    C.reset_config()
    np.random.seed(4)
    train.main(["train", "--config", cfg_path, "--precision", "fp32"])
    assert int(tf_checkpoint.read_checkpoint(str(logdir / "model-2"))["Variable"]) == 9
    C.reset_config()


@pytest.mark.gpu
def test_icp_special_mode(tmp_path):
    """`evaluation.special.mode == 'icp'`, variant p2point with the yaw constraint (the reference's icp_<dataset>_o3_p2p.json,
    train.py:548-551 -> icp.py:150-225): ICP from the centroid initialisation over the validation split, results and
    eval files where the reference writes them, equal to the restated algorithm run pair by pair."""
    import __graft_entry__ as ge
    ge.build()
    from alignnet_b200 import config as C, train
    from alignnet_b200 import synth
    from oracle import icp_ref as I
    base = tmp_path / "SynthTiny"
    # 30 box-shaped pairs with small motion (ICP from the centroid initialisation has a small basin), the last six = val
    synth.write_dataset(str(base), 30, seed=5, points_range=(300, 500), max_rel_angle=0.05, max_speed=0.15)
    cfg = {"data": {"basepath": str(base)}, "logging": {"basedir": str(tmp_path / "logs")},
           "evaluation": {"special": {"mode": "icp", "icp": {"variant": "p2point", "with_constraint": True}}}}
    path = tmp_path / "icp_SynthTiny_o3_p2p.json"
    path.write_text(json.dumps(cfg))
    C.reset_config()
    res = train.main(["eval_only", "--config", str(path)])
    edir = tmp_path / "logs" / "icp_SynthTiny" / "icp_SynthTiny_o3_p2p" / "val" / "eval000000"        # config.py:104-105
    for rel in ("pred_translations.npy", "pred_angles.npy", "pred_s1_pc1centers.npy", "eval.json", "eval_180.json"):
        assert (edir / rel).exists(), rel
    t, a, c = (np.load(edir / f"{k}.npy") for k in ("pred_translations", "pred_angles", "pred_s1_pc1centers"))
    assert t.shape == (6, 3) and a.shape == (6, 1) and c.shape == (6, 3) and not c.any()                # icp.py:207
    assert t.dtype == np.float32 and a.dtype == np.float32
    assert res["eval"]["num"] == 6 and res["eval_180"]["num"] == 6 and res["eval"]["mean_time"] > 0
    val = [int(x) for x in open(base / "split" / "val.txt")]
    gt = np.array([json.load(open(base / "meta" / f"{i:08d}.json"))["rel_angle"] for i in val])
    compared = 0
    for k, idx in enumerate(val):
        p1 = np.load(base / "pointcloud1" / f"{idx:08d}.npy")[:, :3]
        p2 = np.load(base / "pointcloud2" / f"{idx:08d}.npy")[:, :3]
        init = np.eye(4)
        init[:3, 3] = p2.mean(0) - p1.mean(0)                                                           # icp.py:62-66
        T, _, _, _ = I.icp_yaw(p1, p2, init, radius=0.1, its=30)
        if abs(np.arctan2(T[1, 0], T[0, 0]) - gt[k]) > 0.02:
            continue            # the algorithm itself left the basin (a box seen from one side): its path is not reproducible
        compared += 1
        moved = p1 @ T[:3, :3].T + T[:3, 3]
        cs, sn = np.cos(a[k, 0]), np.sin(a[k, 0])
        R = np.array([[cs, -sn, 0], [sn, cs, 0], [0, 0, 1.0]])
        assert np.abs(p1 @ R.T + t[k] - moved).max() < 5e-3, k          # fp32 distances may pick another neighbour here and there
    assert compared >= 4
    assert np.median(np.abs(a[:, 0] - gt)) < 0.02                       # and the refinement finds the motion
    # --use_old_results re-evaluates the stored predictions (icp.py:177-180)
    C.reset_config()
    again = train.main(["eval_only", "--config", str(path), "--use_old_results"])
    assert again["eval"]["corr_levels"] == res["eval"]["corr_levels"]
    C.reset_config()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["bf16", "bf16x6"])
def test_default_architecture_trains_on_tensor_cores(tmp_path, precision):
    """configs/default.json's architecture (two five-layer conv stacks, models/tp8.py:49-59) through the drop-in driver in
    the tensor-core modes: the graph-replayed training loop, evaluation and the checkpoint round trip.  bf16x6 must
    reproduce the fp32 mode's evaluation of the same checkpoint to the parity tolerance."""
    import __graft_entry__ as ge
    ge.build()
    from alignnet_b200 import config as C, train
    cfg_path, logdir = _make_run(tmp_path)
    C.reset_config()
    np.random.seed(3)
    last = train.main(["train", "--config", cfg_path, "--precision", precision])
    assert last["eval"]["num"] == 6
    assert (logdir / "model-1.index").exists() and (logdir / "val/eval000001/eval.json").exists()
    pred = np.load(logdir / "val/eval000001/pred_translations.npy")
    assert pred.shape == (6, 3) and np.isfinite(pred).all()
    out = {}
    for prec in (precision, "fp32"):
        C.reset_config()
        np.random.seed(11)
        train.main(["eval_only", "--config", cfg_path, "--eval_epoch", "1", "--precision", prec])
        out[prec] = (np.load(logdir / "val/eval000001/pred_translations.npy"), np.load(logdir / "val/eval000001/pred_s2_pc1centers.npy"))
    d_t = float(np.abs(out[precision][0] - out["fp32"][0]).max())
    d_c = float(np.abs(out[precision][1] - out["fp32"][1]).max())
    print(precision, "vs fp32 on the same checkpoint: translations", d_t, "stage-2 centres", d_c)
    if precision == "bf16x6":
        assert d_c < 1e-4, d_c                         # upstream of the arg-max canonicalisation: no discontinuity
    else:
        assert d_c < 0.3, d_c                          # bf16 arithmetic on a model trained for twelve steps (measured 0.14)
