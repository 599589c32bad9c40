"""Drop-in for the reference driver `train.py` on the B200 engine.

    python -m alignnet_b200.train train     --config configs/SynthCars.json
    python -m alignnet_b200.train eval_only --config configs/SynthCars.json --eval_epoch 199

Same CLI (train.py:31-39), config schema (the reference's configs/*.json load unchanged), schedules
(train.py:133-174), epoch structure (train.py:296-326: train one epoch, evaluate, checkpoint on even / every fifth /
last epoch) and output files: `<logdir>/config.json`, `<logdir>/out.log`, `<logdir>/val/eval%06d/{eval.json,
eval_180.json, pred_*.npy}` (train.py:399-407,487-543, evaluation.py:274-287), `<logdir>/model.ckpt.*` and
`<logdir>/model-<epoch>.*` in TensorFlow's checkpoint format (`alignnet_b200.tf_checkpoint`, train.py:316-322).
Differences, all deliberate: there is no TensorBoard writer, `--refineICP` runs the batched device ICP of `alignnet_b200.icp` instead of the
authors' Open3D fork (row N4, parity unpinned), and the
val/test split of the synthetic sets (evaluation.py:161-162: idx >= 1000) is computed here and passed to the device
evaluation.  `evaluation.special.mode == "timings"` is the reference's own benchmark harness (train.py:553-559): batch 32,
no restore, ten evaluation passes, `Timing bs=32: <seconds per pair>`; `"icp"` with `icp.variant == "p2point"` runs the ICP
baseline over the validation split (train.py:548-551, icp.py:150-225) on the same device kernel.  `training.optimizer.optimizer`
selects Adam or the momentum optimiser (train.py:211-216).  `training.pretraining.model` is restored like the
reference does (all variables except the global step, then an initial evaluation, train.py:276-293).  One process per GPU;
under torchrun every rank reads only its shard of each batch and the gradient all-reduce is the only collective."""
from __future__ import annotations

import argparse
import datetime
import json
import logging
import os
import time
from typing import Dict, List

import numpy as np
import torch

from . import config as C
from . import dist as D
from . import engine as E
from . import evaluation, icp, provider, schedules, tf_checkpoint

logger = logging.getLogger("tp")


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("operation", choices=["train", "eval_only"], help="Operation to run")
    p.add_argument("--config", required=True, default="", help="Config file")
    p.add_argument("--refineICP", action="store_true", help="refine the predictions with yaw-constrained ICP (eval_only; row N4)")
    p.add_argument("--its", required=False, default=30)
    p.add_argument("--use_old_results", action="store_true")
    p.add_argument("--refineICPmethod", required=False, default="p2p", choices=["p2p"])
    p.add_argument("--eval_epoch", required=False, default="199", help="Epoch to eval in eval_only mode")
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16x3", "bf16x6"],
                   help="bf16 tensor-core mode, fp32 CUDA-core parity mode, or fp32-grade split-bf16 tensor-core modes")
    return p.parse_args(argv)


def _is_test(cfg, idxs: List[int]) -> np.ndarray:
    """evaluation.py:159-162."""
    if "KITTI_tracklets" in cfg.data.basepath:
        out = []
        for i in idxs:
            meta = json.load(open(f"{cfg.data.basepath}/meta/{str(i).zfill(8)}.json"))
            out.append("trackids" in meta and meta["trackids"][0] in [2, 6, 7, 8, 10])
        return np.array(out, bool)
    return np.arange(len(idxs)) >= 1000


def save_checkpoint(eng: E.Engine, prefix: str) -> None:
    """saver.save (train.py:316-322): a TensorFlow tensor-bundle checkpoint `<prefix>.index` + `<prefix>.data-*` with the
    reference graph's variable names (parameters, BN shadows, Adam slots, global step)."""
    tf_checkpoint.save_from_engine(eng, prefix)


def load_checkpoint(eng: E.Engine, prefix: str, restore_step: bool = True) -> None:
    """saver.restore (train.py:250-293); also reads checkpoints written by the reference itself.  With
    `restore_step=False` everything but the global step is taken (the pre-training restore, train.py:277-281)."""
    step = eng.step
    tf_checkpoint.load_into_engine(eng, prefix)
    if not restore_step:
        eng.step = step


def train_one_epoch(cfg, eng: E.Engine, train_idxs: List[int], epoch: int, rank: int, world: int) -> float:
    """train.py:337-393: shuffle, load + jitter each batch, one optimiser step per batch.  The step is replayed as a CUDA
    graph over static batch buffers (what bench.py measures); a new lr / bn_decay plateau captures a new graph.  Data
    parallel: every rank draws the SAME permutation (seeded with the epoch) and reads only its contiguous slice of each
    global batch, so an epoch is a partition of the training set and the host IO is not multiplied by the world size."""
    idxs = list(train_idxs)
    if world > 1:
        np.random.RandomState(1000003 * (epoch + 1)).shuffle(idxs)
    else:
        np.random.shuffle(idxs)                                  # the reference's global-RNG draw (train.py:341)
    bs = cfg.training.batch_size
    nb_epoch = len(train_idxs) // bs
    lo, hi = D.shard_bounds(bs, rank, world)
    mine = [i for b in range(nb_epoch) for i in idxs[b * bs + lo:b * bs + hi]]
    loss_sum, n = 0.0, 0
    allreduce = D.allreduce_grads if world > 1 else None
    loader = provider.Prefetcher(cfg.data.basepath, mine, hi - lo, cfg.model.num_points, jitter=True, device=str(eng.device))
    static = None
    for batch in loader:
        if static is None:
            static = {k: torch.empty_like(v) for k, v in batch.items()}
        for k, v in batch.items():
            static[k].copy_(v)
        lr = schedules.learning_rate(cfg, eng.step, nb_epoch)
        bn_d = schedules.bn_decay(cfg, eng.step, nb_epoch)
        loss = eng.train_step_graph(static, lr=lr, bn_decay=bn_d, allreduce=allreduce, seed_salt=rank)
        loss_sum += float(loss[0].cpu())
        n += 1
    mean = loss_sum / max(n, 1)
    logger.info("train mean loss: %f" % mean)
    return mean


def eval_one_epoch(cfg, eng: E.Engine, val_idxs: List[int], epoch, refine_icp: bool = False, its: int = 30,
                   icp_method: str = "p2p", use_old_results: bool = False, do_timings: bool = False) -> Dict:
    """train.py:396-545: eval-mode forward over the validation split, host decode of the angles (quirk Q1), the
    eval.json metrics with and without accepting the 180-degree flip."""
    eval_dir = f"{cfg.logging.logdir}/val/eval{str(epoch).zfill(6)}"
    base_eval_dir = eval_dir
    if refine_icp:                                               # train.py:401-402
        eval_dir = f'{eval_dir}/refined_{icp_method}{"_" + str(its) if int(its) != 30 else ""}'
    if os.path.isdir(eval_dir):                                  # keep earlier results (train.py:404-405)
        backup, n = f"{eval_dir}_backup_{int(time.time())}", 0
        while os.path.exists(backup if n == 0 else f"{backup}_{n}"):
            n += 1
        os.rename(eval_dir, backup if n == 0 else f"{backup}_{n}")
    os.makedirs(eval_dir, exist_ok=True)
    bs = cfg.training.batch_size
    num_batches = int(np.ceil(len(val_idxs) / bs))
    keep: Dict[str, list] = {k: [] for k in ("pred_translations", "pred_angles", "pred_s1_pc1centers", "pred_s1_pc2centers",
                                             "pred_s2_pc1centers", "pred_s2_pc2centers", "pred_s2_pc1angles",
                                             "pred_s2_pc2angles", "gt_translations", "gt_angles", "gt_pc1centers")}
    old = None
    if use_old_results:                                          # train.py:421-424: predictions of an earlier run
        old = {k: np.load(f"{base_eval_dir}/{k}.npy") for k in ("pred_translations", "pred_angles", "pred_s2_pc1centers")}
    loss_sum, t_exec, num_full_batches = 0.0, [], 0
    for b in range(num_batches):
        chunk = list(val_idxs[b * bs:(b + 1) * bs])
        valid = len(chunk)
        chunk = chunk + [chunk[-1]] * (bs - valid)              # the graph has a static batch size (train.py:190)
        batch = provider.load_batch(cfg.data.basepath, chunk, cfg.model.num_points, jitter=False, device=str(eng.device))
        torch.cuda.synchronize()
        t0 = time.time()
        ep = eng.forward(batch["pcs1"], batch["pcs2"], False)
        loss = eng.loss(batch, ep)
        pa = eng.pred_angles(ep)
        torch.cuda.synchronize()
        t_exec.append((time.time() - t0) / bs)
        if valid == bs:                                          # the padded last batch is not counted (train.py:459-460)
            loss_sum += float(loss[0].cpu())
            num_full_batches += 1
        pt, pang = ep["pred_translations"][:valid].cpu().numpy(), pa[:valid].cpu().numpy()
        a1 = eng.decode_angles(ep["pred_pc1angle_logits"], False)[:valid].cpu().numpy()
        a2 = eng.decode_angles(ep["pred_pc2angle_logits"], False)[:valid].cpu().numpy()
        centers = {k: ep[k][:valid].cpu().numpy() for k in ("pred_s1_pc1centers", "pred_s1_pc2centers", "pred_s2_pc1centers",
                                                            "pred_s2_pc2centers")}
        if refine_icp:                                           # train.py:463-484, the whole batch in one launch
            full1 = [np.load(f"{cfg.data.basepath}/pointcloud1/{str(i).zfill(8)}.npy")[:, :3] for i in chunk[:valid]]
            full2 = [np.load(f"{cfg.data.basepath}/pointcloud2/{str(i).zfill(8)}.npy")[:, :3] for i in chunk[:valid]]
            if old is not None:                                  # train.py:464-465
                sl = slice(b * bs, b * bs + valid)
                inits = np.stack([icp.get_mat_angle(old["pred_translations"][sl][i], float(np.ravel(old["pred_angles"][sl][i])[0]),
                                                    old["pred_s2_pc1centers"][sl][i]) for i in range(valid)])
            else:
                inits = np.stack([icp.get_mat_angle(pt[i], float(pang[i]), centers["pred_s2_pc1centers"][i]) for i in range(valid)])
            t1 = time.time()
            refined, _ = icp.refine(full1, full2, inits, radius=0.1, its=int(its), device=str(eng.device))
            t_exec[-1] += (time.time() - t1) / max(valid, 1)
            rt, ra = icp.to_translation_angle(refined)
            pt, pang = rt.astype(pt.dtype), ra.astype(pang.dtype)
            centers["pred_s2_pc1centers"] = np.zeros_like(centers["pred_s2_pc1centers"])       # rotation about the origin now
        keep["pred_translations"].append(pt)
        keep["pred_angles"].append(pang[:, None])
        keep["pred_s2_pc1angles"].append(a1[:, None])
        keep["pred_s2_pc2angles"].append(a2[:, None])
        for k, v in centers.items():
            keep[k].append(v)
        keep["gt_translations"].append(batch["translations"][:valid].cpu().numpy())
        keep["gt_angles"].append(batch["rel_angles"][:valid].cpu().numpy())
        keep["gt_pc1centers"].append(batch["pc1_centers"][:valid].cpu().numpy())
    arr = {k: np.concatenate(v, axis=0).astype(np.float32) for k, v in keep.items()}      # float32 like train.py:408-419
    mean_time = float(np.sum(t_exec) * bs / max(len(val_idxs), 1))                        # train.py:502
    if do_timings:                                               # train.py:504-505: the reference's benchmark harness
        print(f"Timing bs={bs}: {mean_time}")
        return {"mean_time": mean_time}
    is_test = _is_test(cfg, list(val_idxs))
    result = {}
    for inverted in (False, True):
        d = evaluation.evaluate(arr["pred_translations"], arr["pred_angles"], arr["gt_translations"], arr["gt_angles"],
                                arr["pred_s2_pc1centers"], arr["gt_pc1centers"], is_test, inverted, mean_time,
                                device=str(eng.device))
        with open(f'{eval_dir}/eval{"_180" if inverted else ""}.json', "w") as fh:
            json.dump(d, fh)
        result["eval_180" if inverted else "eval"] = d
    for k in ("pred_translations", "pred_angles", "pred_s1_pc1centers", "pred_s1_pc2centers", "pred_s2_pc1centers",
              "pred_s2_pc2centers", "pred_s2_pc1angles", "pred_s2_pc2angles"):                 # train.py:534-543
        np.save(f"{eval_dir}/{k}.npy", arr[k])
    logger.info("val mean loss: %f" % (loss_sum / num_full_batches if num_full_batches > 0 else 0.0))   # train.py:501
    logger.info("val corr_levels %s (180: %s)" % (result["eval"]["corr_levels"], result["eval_180"]["corr_levels"]))
    return result


def main(argv=None) -> Dict:
    flags = parse_args(argv)
    if flags.refineICP and flags.operation != "eval_only":
        raise ValueError("--refineICP only applies to eval_only (train.py:463)")
    cfg = C.load_config(flags.config)
    C.validate(cfg)
    rank, world, local = D.init() if "RANK" in os.environ else (0, 1, 0)
    os.makedirs(cfg.logging.logdir, exist_ok=True)
    if rank == 0:
        copy = f"{cfg.logging.logdir}/config.json"
        if os.path.exists(copy):
            copy = f'{copy[:-5]}_{datetime.datetime.today().strftime("%Y-%m-%d_%H-%M-%S")}.json'
        C.save_config(copy)
        logging.basicConfig(level=logging.INFO, handlers=[logging.FileHandler(f"{cfg.logging.logdir}/out.log"),
                                                          logging.StreamHandler()])
    if cfg.evaluation.has("special") and cfg.evaluation.special.mode == "icp":          # train.py:548-551: no model at all
        torch.cuda.set_device(local)
        return icp.evaluate(cfg, bool(flags.use_old_results), f"cuda:{local}") if rank == 0 else {}
    train_idxs = provider.get_data_files(f"{cfg.data.basepath}/split/train.txt")
    val_idxs = provider.get_data_files(f"{cfg.data.basepath}/split/val.txt")
    device = f"cuda:{local}"
    torch.cuda.set_device(local)
    timings = cfg.evaluation.has("special") and cfg.evaluation.special.mode == "timings"
    if timings:                                                  # train.py:553-559
        cfg.training.__dict__["batch_size"] = 32
    eng = E.Engine(C.arch_from_config(cfg), device, flags.precision, seed=0)
    eng.set_optimizer(*C.optimizer_from_config(cfg))             # train.py:211-216
    start_epoch = 0
    ckpt = f"{cfg.logging.logdir}/model.ckpt"
    eval_only = flags.operation == "eval_only" or timings
    nb_epoch = max(1, len(train_idxs) // cfg.training.batch_size)
    if eval_only:
        start_epoch = int(flags.eval_epoch)
        if not flags.use_old_results and not timings:            # train.py:250-264
            path = f"{cfg.logging.logdir}/model-{int(flags.eval_epoch)}"
            if not os.path.isfile(path + ".index"):
                raise FileNotFoundError(path + ".index")
            load_checkpoint(eng, path)
            if eng.step % nb_epoch != 0 or eng.step // nb_epoch - 1 != int(flags.eval_epoch):
                raise ValueError(f"{path}: global step {eng.step} is not the end of epoch {flags.eval_epoch} "
                                 f"({nb_epoch} batches per epoch)")
        logger.info(f"Evaluating at epoch {start_epoch}")
    elif os.path.isfile(ckpt + ".index"):                        # resume (train.py:267-275)
        load_checkpoint(eng, ckpt)
        if eng.step % nb_epoch != 0:
            raise ValueError(f"{ckpt}: global step {eng.step} is not a multiple of {nb_epoch} batches per epoch")
        start_epoch = eng.step // nb_epoch
        logger.info(f"Continuing training at epoch {start_epoch}")
    elif cfg.training.pretraining.model != "":                   # train.py:276-293
        pre = cfg.training.pretraining.model
        if not os.path.isfile(pre + ".index"):
            raise FileNotFoundError(pre + ".index")
        load_checkpoint(eng, pre, restore_step=False)
        assert eng.step == 0
        logger.info(f"Pre-trained weights loaded from {pre}, starting initial evaluation")
        if rank == 0:
            eval_one_epoch(cfg, eng, val_idxs, "pretr")
        logger.info("Initial evaluation finished")
    last = {}
    for epoch in range(start_epoch, cfg.training.num_epochs):
        logger.info("**** EPOCH %03d ****    lr: %.8f, bn_decay: %.8f" % (epoch, schedules.learning_rate(cfg, eng.step, nb_epoch),
                                                                           schedules.bn_decay(cfg, eng.step, nb_epoch)))
        if not eval_only:
            train_one_epoch(cfg, eng, train_idxs, epoch, rank, world)
        if rank == 0:
            for _ in range(10 if timings else 1):                # train.py:305-307
                last = eval_one_epoch(cfg, eng, val_idxs, epoch, refine_icp=bool(flags.refineICP), its=int(flags.its),
                                      icp_method=flags.refineICPmethod, use_old_results=bool(flags.use_old_results),
                                      do_timings=timings)
        if eval_only:
            break
        was_last = epoch == cfg.training.num_epochs - 1
        if rank == 0 and (epoch % 2 == 0 or was_last):
            save_checkpoint(eng, ckpt)
        if rank == 0 and (epoch % 5 == 0 or was_last or cfg.evaluation.save_every_epoch):
            save_checkpoint(eng, f"{cfg.logging.logdir}/model-{epoch}")
    logger.info("Finished Training")
    return last


if __name__ == "__main__":
    main()
