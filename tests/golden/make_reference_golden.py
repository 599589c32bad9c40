"""Generates `tests/golden/reference_*.npz` by EXECUTING THE REFERENCE'S OWN CODE: the unmodified
`/root/reference/models/tp8.py` + `utils/tf_util.py` (get_model, get_loss, classLogits2angle) are
imported and run on the eager TF1 shim of `oracle/tf1_shim` (TensorFlow 1.8 itself is not
installable here), on the same seeded inputs / parameters / dropout masks as the oracle fixtures of
`make_golden.py`.  Run from the repo root in the build container (needs /root/reference):

    python tests/golden/make_reference_golden.py

Stored per case: the 8 end_points in eval and train mode from a float32 run (TF semantics) and a
float64 run (rounding-free), the host-decoded pred_angles (train.py:453-456), the loss, the updated
EMA shadows, every gradient's L2 norm and the gradients of all tensors <= 2048 elements (float64 run),
and the variable names/shapes the reference graph created.  The rigid-transform fixture comes from
`tp_utils/pointcloud.py` (get_mat_angle, transform_points, translate_transform_to_new_center_of_rotation)
and `utils/eulerangles.py` (euler2mat), imported likewise.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import reference_run as RR  # noqa: E402
from helpers import golden_case  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case(name, accept_inverted=True):
    g, arch, params, state, batch, masks = golden_case(name)
    if accept_inverted is None:
        accept_inverted = arch.accept_inverted_angle      # the architecture's own setting
    arch.accept_inverted_angle = accept_inverted
    out = {"source": "reference models/tp8.py executed on oracle/tf1_shim", "case": name,
           "accept_inverted_angle": accept_inverted}
    for tag, double in (("f32", False), ("f64", True)):
        ev = RR.run(batch, arch, params, state, False, double=double, with_loss=True)
        tr = RR.run(batch, arch, params, state, True, 0.5, masks, double=double, with_loss=True, with_grads=double)
        for k, v in ev["end_points"].items():
            out[f"{tag}/eval/{k}"] = v
        for k, v in tr["end_points"].items():
            out[f"{tag}/train/{k}"] = v
        out[f"{tag}/eval/pred_angles"] = ev["pred_angles"]
        out[f"{tag}/eval/loss"] = np.float64(ev["loss"])
        out[f"{tag}/train/loss"] = np.float64(tr["loss"])
        if double:
            for k, v in tr["grads"].items():
                out["gradnorm/" + k] = np.float64(np.sqrt((v.astype(np.float64) ** 2).sum()))
                if v.size <= 2048:
                    out["grad/" + k] = v
            for k, v in tr["new_state"].items():
                if v.size <= 256:
                    out["state/" + k] = v
            out["var_names"] = np.array(tr["var_names"])
            out["var_shapes"] = np.array([" ".join(map(str, tr["var_shapes"][n])) for n in tr["var_names"]])
            out["trainable"] = np.array(tr["trainable"])
            out["shadow_names"] = np.array(sorted(tr["new_state"].keys()))
    suffix = "" if (accept_inverted or name.startswith("default")) else "_noinv"
    np.savez_compressed(os.path.join(HERE, f"reference_{name}{suffix}.npz"), **out)
    print(name, suffix, "loss", out["f64/train/loss"])


def rigid():
    pc = RR.load_pointcloud_module()
    eu = RR.load_eulerangles_module()
    rng = np.random.Generator(np.random.PCG64(7))
    n, npts = 16, 40
    t, th, c = rng.normal(size=(n, 3)), rng.uniform(-np.pi, np.pi, size=n), rng.normal(size=(n, 3)) * 5
    pts = rng.normal(size=(n, npts, 3)) * 3
    hom = np.concatenate([pts, np.ones((n, npts, 1))], axis=2)
    mats = np.stack([pc.get_mat_angle(t[i], th[i], c[i]) for i in range(n)])
    moved = np.stack([pc.transform_points(hom[i].copy(), mats[i]) for i in range(n)])
    composed = np.stack([pc.transform_points(hom[i].copy(), [mats[i], mats[(i + 1) % n]]) for i in range(n)])
    new_c = rng.normal(size=(n, 3)) * 5
    t_new = pc.translate_transform_to_new_center_of_rotation(t, th[:, None], c, new_c)
    rz = np.stack([eu.euler2mat(z=a) for a in th])
    np.savez_compressed(os.path.join(HERE, "reference_rigid.npz"), t=t, theta=th, c=c, pts=pts, mats=mats, moved=moved,
                        composed=composed, new_c=new_c, t_new=np.asarray(t_new), rz=rz,
                        source="reference tp_utils/pointcloud.py + utils/eulerangles.py")
    print("rigid written")


if __name__ == "__main__":
    assert RR.available(), "needs /root/reference"
    case("tiny_B4_N16")
    case("tiny_B4_N16", accept_inverted=False)
    case("shipped_B4_N16")
    case("shipped_B32_N200")
    case("default_B32_N64", accept_inverted=None)                   # configs/default.json's architecture (five-layer stacks)
    rigid()
