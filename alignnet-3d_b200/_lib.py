"""ctypes binding of libalignnet_b200.so (the C ABI declared in include/alignnet_b200.h).

Fails loudly when the library is missing: there is no Python / CPU fallback for any entry point.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "csrc" / "libalignnet_b200.so"

MAX_LAYERS = 8

OK = 0
ERR_NAMES = {0: "AN3D_OK", -1: "AN3D_ERR_INVALID", -2: "AN3D_ERR_UNSUPPORTED", -3: "AN3D_ERR_NO_DEVICE",
             -4: "AN3D_ERR_ARCH", -5: "AN3D_ERR_CUDA", -6: "AN3D_ERR_ALIGN", -7: "AN3D_ERR_WORKSPACE"}

TRAINING = 1
PRECISION_FP32 = 0
PRECISION_BF16 = 2
PRECISION_BF16X3 = 16
PRECISION_BF16X6 = 32
WEIGHTS_PREPARED = 4
DETERMINISTIC = 8


class Arch(C.Structure):
    _fields_ = [
        ("num_bins", C.c_int32), ("accept_inverted_angle", C.c_int32),
        ("angle_factor", C.c_float), ("early_stage_factor", C.c_float),
        ("n_conv", C.c_int32 * 3), ("conv", (C.c_int32 * MAX_LAYERS) * 3),
        ("n_fc", C.c_int32 * 3), ("fc", (C.c_int32 * MAX_LAYERS) * 3),
        ("keep_prob", C.c_float * 3),
    ]


class Outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "pred_s1_pc1centers", "pred_s1_pc2centers", "pred_s2_pc1centers", "pred_s2_pc2centers",
        "pred_pc1angle_logits", "pred_pc2angle_logits", "pred_translations", "pred_remaining_angle_logits")]


class Labels(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "translations", "rel_angles", "pc1_centers", "pc2_centers", "pc1_angles", "pc2_angles")]


class Dropout(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("masks", C.c_void_p * 5), ("seed_dev", C.c_void_p)]


# symbol -> (restype, argtypes); mirrors include/alignnet_b200.h one to one
SIGNATURES = {
    "an3d_version": (C.c_int, []),
    "an3d_last_error": (C.c_char_p, []),
    "an3d_create": (C.c_int, [C.POINTER(Arch), C.POINTER(C.c_void_p)]),
    "an3d_destroy": (C.c_int, [C.c_void_p]),
    "an3d_num_elements": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]),
    "an3d_num_tensors": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]),
    "an3d_tensor_info": (C.c_int, [C.c_void_p, C.c_int, C.c_int32, C.c_char_p, C.c_int32, C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int64 * 4)]),
    "an3d_workspace_bytes": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
    "an3d_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                               C.c_int32, C.c_float, C.POINTER(Dropout), C.POINTER(Outputs), C.c_void_p, C.c_int64,
                               C.c_void_p]),
    "an3d_loss": (C.c_int, [C.c_void_p, C.POINTER(Labels), C.POINTER(Outputs), C.c_int32, C.c_void_p, C.c_void_p,
                            C.c_int64, C.c_void_p]),
    "an3d_loss_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Labels),
                                     C.POINTER(Outputs), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_void_p]),
    "an3d_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int64,
                                 C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "an3d_step_advance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "an3d_adam_step_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p,
                                     C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "an3d_momentum_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float,
                                     C.c_void_p]),
    "an3d_decode_angles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "an3d_rigid_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                   C.c_void_p]),
    "an3d_transform_pcs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_void_p]),
    "an3d_loss_p2p": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "an3d_launch_count": (C.c_uint64, []),
    "an3d_profile_begin": (C.c_int, []),
    "an3d_profile_end": (C.c_int, [C.POINTER(C.c_float * 8), C.POINTER(C.c_int32 * 8)]),
    "an3d_selftest_conv_stack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "an3d_selftest_fc_gemm": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                        C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "an3d_selftest_split_gemm": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                           C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_void_p]),
    "an3d_icp_yaw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                               C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "an3d_resample_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "an3d_evaluate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                C.c_int32, C.c_void_p, C.c_void_p]),
    "an3d_bench_umma": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "an3d_selftest_umma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p]),
    "an3d_recenter_translations": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                             C.c_void_p]),
}

_lib = None


class An3dError(RuntimeError):
    def __init__(self, code: int, where: str, msg: str):
        self.code = code
        super().__init__(f"{where}: {ERR_NAMES.get(code, code)}: {msg}")


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("ALIGNNET_B200_LIB", LIB_PATH))
    if not path.exists():
        raise ImportError(
            f"{path} not found: build it with `python alignnet-3d_b200/build.py` (or __graft_entry__.build()). "
            "alignnet_b200 has no CPU/PyTorch fallback.")
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, where: str) -> None:
    if code != OK:
        msg = load().an3d_last_error()
        raise An3dError(code, where, msg.decode() if msg else "")
