"""Convergence equivalence of the two precisions: the fp32 parity mode (held to 1e-4 of the oracle) and the bf16 fast
mode (the one bench.py times) are trained from the same initial parameters on the same seeded synthetic stream with
the same dropout seeds, and must arrive at the same place.

bf16 gradients are NOT close to fp32 gradients step by step (cosine ~0.8 at B=1024, the same distance the CPU
oracle's own bf16 rounding model shows -- tests/test_zz_fullsize_oracle.py): the loss is discontinuous in roundings
(arg-max bins, max-pool rows).  What matters for the reference's users is where training goes, so this test compares
loss trajectories and the reference's own evaluation metrics (evaluation.py:16-46 through `an3d_evaluate`) on a
held-out set.  Bands are ~4x the differences measured on B200 (profiles/r2_convergence.txt)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

STEPS, BATCH, POINTS, POOL = 1400, 256, 200, 16
EVAL_AT = (1000, 1200, 1400)        # the translation error is averaged over these checkpoints (it swings between them)


@pytest.fixture(scope="module", autouse=True)
def _build():
    import __graft_entry__ as ge
    ge.build()


def _train(precision):
    from alignnet_b200 import engine, evaluation, synth
    dev = lambda d: {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in d.items()}
    train = [dev(synth.make_batch_fast(BATCH, POINTS, seed=1000 + i)) for i in range(POOL)]
    val_host = synth.make_batch_fast(1024, POINTS, seed=999)
    val = dev(val_host)
    e = engine.Engine(engine.shipped_arch(), "cuda:0", precision, seed=3)

    def evaluate():
        ep = e.forward(val["pcs1"], val["pcs2"], False)
        ev = evaluation.evaluate(ep["pred_translations"], e.pred_angles(ep), val_host["translations"], val_host["rel_angles"],
                                 ep["pred_s2_pc1centers"], val_host["pc1_centers"], accept_inverted_angle=True)
        return float(e.loss(val, ep)[0].cpu()), ev

    losses, evs, val_loss = [], [], None
    for t in range(STEPS):
        losses.append(float(e.train_step(train[t % POOL], lr=1e-3, bn_decay=0.5, seed=t)[0].cpu()))
        if t + 1 in EVAL_AT:
            val_loss, ev = evaluate()
            evs.append(ev)
    ev = dict(evs[-1])
    ev["mean_dist_translation_per_checkpoint"] = [x["mean_dist_translation"] for x in evs]
    ev["mean_dist_translation"] = float(np.mean(ev["mean_dist_translation_per_checkpoint"]))
    return np.asarray(losses), val_loss, ev


def test_bf16_training_converges_like_fp32():
    """Measured (profiles/r2_convergence.txt, two dropout seeds per precision, checkpoints every 200 steps up to 1600):
    final training loss 0.0375 (fp32) vs 0.0378 (bf16), held-out loss 0.00931 vs 0.00936, mean translation error
    0.107-0.118 m vs 0.109-0.139 m; mid-training the bf16 runs trail the fp32 runs in translation error (step 1000: 0.12 vs
    0.155 m in the probe, 0.131 vs 0.199 m in one later run of this test) and catch up by step 1400, and a single run's
    error swings by +-20 % from one checkpoint to the next.  Training is not reproducible run to run (fp32 atomics in the
    backward's weight-gradient sums, Adam's sign-like early updates -- tests/test_gpu_determinism.py), so every band here
    is at least twice the largest difference seen over all recorded runs, and the translation error is the mean over the
    checkpoints at steps 1000 / 1200 / 1400 (probe: fp32 0.118-0.119, bf16 0.136-0.142).  The angle metrics of this early
    phase swing by +-15 degrees between checkpoints and between two fp32 dropout seeds, so they are only required to be
    in the same regime."""
    l32, v32, ev32 = _train("fp32")
    l16, v16, ev16 = _train("bf16")
    first32, last32, last16 = l32[:20].mean(), l32[-50:].mean(), l16[-50:].mean()
    print(f"train loss: first-20 {first32:.4f}; last-50 fp32 {last32:.4f} bf16 {last16:.4f}; val loss fp32 {v32:.5f} bf16 {v16:.5f}")
    keys = ("corr_levels_translation", "mean_dist_translation", "mean_dist_translation_per_checkpoint", "corr_levels_angles",
            "mean_dist_angle")
    print("fp32 eval:", {k: ev32[k] for k in keys})
    print("bf16 eval:", {k: ev16[k] for k in keys})
    assert np.isfinite(l32).all() and np.isfinite(l16).all()
    # both learn: the loss more than halves, the translation error falls from ~0.7 m (step 200) below 0.25 m
    assert last32 < 0.5 * first32 and last16 < 0.5 * l16[:20].mean(), (first32, last32, last16)
    assert ev32["mean_dist_translation"] < 0.25 and ev16["mean_dist_translation"] < 0.25
    # and they learn the same thing: loss trajectories (50-step windows; recorded runs: 1.7-3.5 %), final training
    # loss (recorded: 0.2-2.6 %), held-out loss (recorded: 0.2-1 %)
    w32, w16 = l32.reshape(-1, 50).mean(1), l16.reshape(-1, 50).mean(1)
    wdiff = np.abs(w16[2:] / w32[2:] - 1)
    print(f"windows: max |bf16 / fp32 - 1| = {wdiff.max():.4f}; last-50 {abs(last16 / last32 - 1):.4f}; val {abs(v16 / v32 - 1):.4f}; "
          f"translation ratio {ev16['mean_dist_translation'] / ev32['mean_dist_translation']:.3f}")
    assert wdiff.max() <= 0.10, (w32, w16)
    assert abs(last16 - last32) <= 0.06 * last32, (last16, last32)
    assert abs(v16 - v32) <= 0.04 * v32, (v16, v32)
    # translation error, averaged over the three checkpoints: recorded ratios 0.89-1.20
    assert ev16["mean_dist_translation"] <= 1.6 * ev32["mean_dist_translation"]
    # angles: same regime (see the docstring for why not tighter)
    assert abs(ev16["mean_dist_angle"] - ev32["mean_dist_angle"]) <= 30.0


def test_bf16x6_training_tracks_fp32_like_a_second_fp32_run():
    """The six-product split mode is held to the fp32 mode's single-step tolerances elsewhere
    (tests/test_gpu_parity.py, tests/test_zz_fullsize_oracle.py); here the engines take 60 optimiser steps from the same
    initial parameters on the same stream with the same dropout seeds.  Early training is chaotic -- Adam's first updates
    are sign-like, so a near-zero gradient element moves its weight by +lr in one run and -lr in another -- and two runs
    of the fp32 engine itself separate (its backward reduces with fp32 atomics).  The yardstick for bf16x6 is therefore a
    SECOND fp32 run: its loss trajectory must stay as close to fp32 run A as fp32 run B does (within a factor 3 on the
    mean and maximum relative distance, with a floor for the case that the two fp32 runs happen to coincide)."""
    from alignnet_b200 import engine, synth
    dev = lambda d: {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in d.items()}
    steps, B, N = 60, 128, 160
    train = [dev(synth.make_batch_fast(B, N, seed=2000 + i)) for i in range(8)]
    traj = {}
    for name, prec in (("fp32_a", "fp32"), ("fp32_b", "fp32"), ("bf16x6", "bf16x6")):
        e = engine.Engine(engine.shipped_arch(), "cuda:0", prec, seed=5)
        traj[name] = np.asarray([float(e.train_step(train[t % 8], lr=1e-3, bn_decay=0.5, seed=t)[0].cpu()) for t in range(steps)])
    rel_b = np.abs(traj["fp32_b"] / traj["fp32_a"] - 1)
    rel_x = np.abs(traj["bf16x6"] / traj["fp32_a"] - 1)
    print(f"|loss / loss(fp32 run A) - 1| over 60 steps: second fp32 run mean {rel_b.mean():.2e} max {rel_b.max():.2e} (first step "
          f"{rel_b[0]:.1e}); bf16x6 mean {rel_x.mean():.2e} max {rel_x.max():.2e} (first step {rel_x[0]:.1e})")
    assert np.isfinite(traj["bf16x6"]).all() and traj["bf16x6"][-10:].mean() < traj["bf16x6"][:10].mean()
    assert rel_x[0] < 1e-4                                  # same parameters, same batch: the single-step tolerance
    assert rel_x.mean() <= max(3 * rel_b.mean(), 2e-2) and rel_x.max() <= max(3 * rel_b.max(), 8e-2), (rel_x.mean(), rel_b.mean())
