"""Drop-in for the reference driver `train.py` on the B200 engine.

    python -m alignnet_b200.train train     --config configs/SynthCars.json
    python -m alignnet_b200.train eval_only --config configs/SynthCars.json --eval_epoch 199

Same CLI (train.py:31-39), config schema (the reference's configs/*.json load unchanged), schedules
(train.py:133-174), epoch structure (train.py:296-326: train one epoch, evaluate, checkpoint on even / every fifth /
last epoch) and output files: `<logdir>/config.json`, `<logdir>/out.log`, `<logdir>/val/eval%06d/{eval.json,
eval_180.json, pred_*.npy}` (train.py:399-407,487-543, evaluation.py:274-287), `<logdir>/model.ckpt.*` and
`<logdir>/model-<epoch>.*` in TensorFlow's checkpoint format (`alignnet_b200.tf_checkpoint`, train.py:316-322).
Differences, all deliberate: there is no TensorBoard writer, `--refineICP` runs the batched device ICP of `alignnet_b200.icp` instead of the
authors' Open3D fork (row N4, parity unpinned; `--use_old_results` is accepted and ignored), and the
val/test split of the synthetic sets (evaluation.py:161-162: idx >= 1000) is computed here and passed to the device
evaluation.  One process per GPU; under torchrun the gradient all-reduce is the only collective."""
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
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32"], help="bf16 tensor-core mode or fp32 parity mode")
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


def load_checkpoint(eng: E.Engine, prefix: str) -> None:
    """saver.restore (train.py:250-293); also reads checkpoints written by the reference itself."""
    tf_checkpoint.load_into_engine(eng, prefix)


def train_one_epoch(cfg, eng: E.Engine, train_idxs: List[int], epoch: int, rank: int, world: int) -> float:
    """train.py:337-393: shuffle, load + jitter each batch, one optimiser step per batch."""
    idxs = list(train_idxs)
    np.random.shuffle(idxs)
    bs = cfg.training.batch_size
    nb_epoch = len(train_idxs) // bs
    lo, hi = D.shard_bounds(bs, rank, world)
    loss_sum, n = 0.0, 0
    allreduce = D.allreduce_grads if world > 1 else None
    loader = provider.Prefetcher(cfg.data.basepath, idxs, bs, cfg.model.num_points, jitter=True, device=str(eng.device))
    for batch in loader:
        if world > 1:
            batch = {k: v[lo:hi].contiguous() for k, v in batch.items()}
        lr = schedules.learning_rate(cfg, eng.step, nb_epoch)
        bn_d = schedules.bn_decay(cfg, eng.step, nb_epoch)
        loss = eng.train_step(batch, lr=lr, bn_decay=bn_d, allreduce=allreduce)
        loss_sum += float(loss[0].cpu())
        n += 1
    mean = loss_sum / max(n, 1)
    logger.info("train mean loss: %f" % mean)
    return mean


def eval_one_epoch(cfg, eng: E.Engine, val_idxs: List[int], epoch: int, refine_icp: bool = False, its: int = 30,
                   icp_method: str = "p2p") -> Dict:
    """train.py:396-545: eval-mode forward over the validation split, host decode of the angles (quirk Q1), the
    eval.json metrics with and without accepting the 180-degree flip."""
    eval_dir = f"{cfg.logging.logdir}/val/eval{str(epoch).zfill(6)}"
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
                                             "pred_s2_pc1centers", "pred_s2_pc2centers", "gt_translations", "gt_angles",
                                             "gt_pc1centers")}
    loss_sum, t_exec = 0.0, []
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
        loss_sum += float(loss[0].cpu())
        pt, pang = ep["pred_translations"][:valid].cpu().numpy(), pa[:valid].cpu().numpy()
        centers = {k: ep[k][:valid].cpu().numpy() for k in ("pred_s1_pc1centers", "pred_s1_pc2centers", "pred_s2_pc1centers",
                                                            "pred_s2_pc2centers")}
        if refine_icp:                                           # train.py:463-484, the whole batch in one launch
            full1 = [np.load(f"{cfg.data.basepath}/pointcloud1/{str(i).zfill(8)}.npy")[:, :3] for i in chunk[:valid]]
            full2 = [np.load(f"{cfg.data.basepath}/pointcloud2/{str(i).zfill(8)}.npy")[:, :3] for i in chunk[:valid]]
            inits = np.stack([icp.get_mat_angle(pt[i], float(pang[i]), centers["pred_s2_pc1centers"][i]) for i in range(valid)])
            t1 = time.time()
            refined, _ = icp.refine(full1, full2, inits, radius=0.1, its=int(its), device=str(eng.device))
            t_exec[-1] += (time.time() - t1) / max(valid, 1)
            rt, ra = icp.to_translation_angle(refined)
            pt, pang = rt.astype(pt.dtype), ra.astype(pang.dtype)
            centers["pred_s2_pc1centers"] = np.zeros_like(centers["pred_s2_pc1centers"])       # rotation about the origin now
        keep["pred_translations"].append(pt)
        keep["pred_angles"].append(pang[:, None])
        for k, v in centers.items():
            keep[k].append(v)
        keep["gt_translations"].append(batch["translations"][:valid].cpu().numpy())
        keep["gt_angles"].append(batch["rel_angles"][:valid].cpu().numpy())
        keep["gt_pc1centers"].append(batch["pc1_centers"][:valid].cpu().numpy())
    arr = {k: np.concatenate(v, axis=0).astype(np.float64) for k, v in keep.items()}
    is_test = _is_test(cfg, list(val_idxs))
    result = {}
    for inverted in (False, True):
        d = evaluation.evaluate(arr["pred_translations"], arr["pred_angles"], arr["gt_translations"], arr["gt_angles"],
                                arr["pred_s2_pc1centers"], arr["gt_pc1centers"], is_test, inverted, float(np.mean(t_exec)),
                                device=str(eng.device))
        with open(f'{eval_dir}/eval{"_180" if inverted else ""}.json', "w") as fh:
            json.dump(d, fh)
        result["eval_180" if inverted else "eval"] = d
    for k in ("pred_translations", "pred_angles", "pred_s1_pc1centers", "pred_s1_pc2centers", "pred_s2_pc1centers",
              "pred_s2_pc2centers"):
        np.save(f"{eval_dir}/{k}.npy", arr[k])
    logger.info("val mean loss: %f" % (loss_sum / max(num_batches, 1)))
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
    train_idxs = provider.get_data_files(f"{cfg.data.basepath}/split/train.txt")
    val_idxs = provider.get_data_files(f"{cfg.data.basepath}/split/val.txt")
    device = f"cuda:{local}"
    torch.cuda.set_device(local)
    eng = E.Engine(C.arch_from_config(cfg), device, flags.precision, seed=0)
    start_epoch = 0
    ckpt = f"{cfg.logging.logdir}/model.ckpt"
    eval_only = flags.operation == "eval_only"
    if eval_only:
        path = f"{cfg.logging.logdir}/model-{int(flags.eval_epoch)}"
        load_checkpoint(eng, path if os.path.isfile(path + ".index") else ckpt)
        start_epoch = int(flags.eval_epoch)
    elif os.path.isfile(ckpt + ".index"):                        # resume (train.py:267-270)
        load_checkpoint(eng, ckpt)
        start_epoch = eng.step // max(1, len(train_idxs) // cfg.training.batch_size)
    last = {}
    for epoch in range(start_epoch, cfg.training.num_epochs):
        nb_epoch = max(1, len(train_idxs) // cfg.training.batch_size)
        logger.info("**** EPOCH %03d ****    lr: %.8f, bn_decay: %.8f" % (epoch, schedules.learning_rate(cfg, eng.step, nb_epoch),
                                                                           schedules.bn_decay(cfg, eng.step, nb_epoch)))
        if not eval_only:
            train_one_epoch(cfg, eng, train_idxs, epoch, rank, world)
        if rank == 0:
            last = eval_one_epoch(cfg, eng, val_idxs, epoch, refine_icp=bool(flags.refineICP), its=int(flags.its),
                                  icp_method=flags.refineICPmethod)
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
