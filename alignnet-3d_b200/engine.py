"""Host-side engine: owns the flat parameter / state / optimiser buffers (torch CUDA tensors used
purely as device-memory containers) and drives libalignnet_b200.so through the C ABI.

Replaces, for the tp8 model, what the reference does with a tf.Session: variable storage
(utils/tf_util.py:19-49), `sess.run` of the forward / loss / train_op fetches (train.py:368,448)
and the Adam optimiser (train.py:212-217).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib

OUTPUT_KEYS = ("pred_s1_pc1centers", "pred_s1_pc2centers", "pred_s2_pc1centers", "pred_s2_pc2centers",
               "pred_pc1angle_logits", "pred_pc2angle_logits", "pred_translations", "pred_remaining_angle_logits")
LABEL_KEYS = ("translations", "rel_angles", "pc1_centers", "pc2_centers", "pc1_angles", "pc2_angles")
MASK_KEYS = ("s1_b0", "s1_b1", "s2_b0", "s2_b1", "head")
LOSS_PARTS = ("per_transform_loss", "translation", "angle", "stage1_pc1_transl_loss", "stage1_pc2_transl_loss",
              "stage2_pc1_transl_loss", "stage2_pc2_transl_loss", "stage3_transl_loss", "stage2_pc1_angle_loss",
              "stage2_pc1_angle_class_loss", "stage2_pc1_angle_residual_loss", "stage2_pc2_angle_loss",
              "stage2_pc2_angle_class_loss", "stage2_pc2_angle_residual_loss", "stage3_angle_loss",
              "stage3_angle_class_loss", "stage3_angle_residual_loss")


def shipped_arch(num_bins: int = 50, accept_inverted_angle: bool = True, angle_factor: float = 1.0,
                 early_stage_factor: float = 0.5) -> _lib.Arch:
    """The architecture every shipped reference config uses (configs/SynthCars.json:8-20)."""
    return make_arch(num_bins, (64, 128, 256), (512, 256), 0.7, (64, 128, 512), (512, 256), 0.7, (64, 128, 1024),
                     (512, 256), 0.7, angle_factor, early_stage_factor, accept_inverted_angle)


def default_arch() -> _lib.Arch:
    """The architecture of the reference's configs/default.json:8-22 (five-layer conv stacks, 36 bins): not of the
    fused kernels' [64,128,C] form -- the bf16 mode runs it layer by layer on the tensor cores."""
    return make_arch(36, (128, 128, 256), (512, 256), 0.7, (64, 64, 64, 128, 1024), (512, 256), 0.7,
                     (64, 64, 64, 128, 1024), (512, 256), 0.7, 1.0, 0.1, False)


def make_arch(num_bins, s1_conv, s1_fc, s1_keep, s2_conv, s2_fc, s2_keep, emb_conv, head_fc, head_keep,
              angle_factor=1.0, early_stage_factor=0.5, accept_inverted_angle=True) -> _lib.Arch:
    a = _lib.Arch()
    a.num_bins = int(num_bins)
    a.accept_inverted_angle = int(bool(accept_inverted_angle))
    a.angle_factor = float(angle_factor)
    a.early_stage_factor = float(early_stage_factor)
    for s, (conv, fc, keep) in enumerate(((s1_conv, s1_fc, s1_keep), (s2_conv, s2_fc, s2_keep),
                                          (emb_conv, head_fc, head_keep))):
        a.n_conv[s] = len(conv)
        for i, c in enumerate(conv):
            a.conv[s][i] = int(c)
        a.n_fc[s] = len(fc)
        for i, c in enumerate(fc):
            a.fc[s][i] = int(c)
        a.keep_prob[s] = 1.0 if keep is None else float(keep)
    return a


class Layout:
    """Flat-buffer layout reported by the library (TF variable names -> offset/shape)."""

    def __init__(self, lib, ctx, which: int):
        n = C.c_int32()
        _lib.check(lib.an3d_num_tensors(ctx, which, C.byref(n)), "an3d_num_tensors")
        total = C.c_int64()
        _lib.check(lib.an3d_num_elements(ctx, which, C.byref(total)), "an3d_num_elements")
        self.total = int(total.value)
        self.entries: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        self.order = []
        buf = C.create_string_buffer(256)
        for i in range(n.value):
            off, nd, shp = C.c_int64(), C.c_int32(), (C.c_int64 * 4)()
            _lib.check(lib.an3d_tensor_info(ctx, which, i, buf, 256, C.byref(off), C.byref(nd), C.byref(shp)),
                       "an3d_tensor_info")
            name = buf.value.decode()
            self.entries[name] = (int(off.value), tuple(int(shp[k]) for k in range(nd.value)))
            self.order.append(name)


class Engine:
    def __init__(self, arch: _lib.Arch, device: Optional[str] = None, precision: str = "fp32", seed: int = 0,
                 allocate: bool = True, cache_eval_weights: bool = True, deterministic: bool = False):
        """cache_eval_weights (on by default; AN3D_EVAL_CACHE=0 turns it off): bf16 inference calls after the first
        one on unchanged parameters pass AN3D_WEIGHTS_PREPARED and skip the ~40 launches that fold the BN layers and
        pack the weight images (c2 on B200: 0.588 -> 0.501 ms, profiles/r2_ab_switches.txt).  Every Engine method that
        changes parameters or BN state invalidates the cache; code that writes `engine.params` / `engine.bn_state`
        directly must call `params_changed()`."""
        self.lib = _lib.load()
        # AN3D_DETERMINISTIC: bit-reproducible inference too (training-mode forwards always are); see the C header
        self.deterministic = bool(deterministic) or os.environ.get("AN3D_DETERMINISTIC") == "1"
        self.cache_eval_weights = bool(cache_eval_weights) and os.environ.get("AN3D_EVAL_CACHE") != "0"
        self._pversion = 0                       # bumped whenever params / bn_state may have changed
        self._prepared: Dict[Tuple[int, int, int], int] = {}
        self.arch = arch
        ctx = C.c_void_p()
        _lib.check(self.lib.an3d_create(C.byref(arch), C.byref(ctx)), "an3d_create")
        self.ctx = ctx
        self.params_layout = Layout(self.lib, ctx, 0)
        self.state_layout = Layout(self.lib, ctx, 1)
        self.num_bins = int(arch.num_bins)
        self.set_precision(precision)
        self.step = 0          # global step (train.py:195)
        self._ws: Dict[Tuple[int, int, int], torch.Tensor] = {}
        self._out: Dict[int, Dict[str, torch.Tensor]] = {}
        self.device = None
        if allocate:
            if not torch.cuda.is_available():
                raise RuntimeError("alignnet_b200.Engine needs a CUDA (sm_100) device; there is no CPU fallback")
            self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
            n = self.params_layout.total
            self.params = torch.empty(n, dtype=torch.float32, device=self.device)
            self.grads = torch.zeros(n, dtype=torch.float32, device=self.device)
            self.adam_m = torch.zeros(n, dtype=torch.float32, device=self.device)
            self.adam_v = torch.zeros(n, dtype=torch.float32, device=self.device)
            self.bn_state = torch.zeros(self.state_layout.total, dtype=torch.float32, device=self.device)
            self.loss_buf = torch.zeros(20, dtype=torch.float32, device=self.device)
            # device-resident copies of the global step and the dropout seed: what a captured CUDA graph reads
            self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
            self.seed_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
            self._graphs: Dict[tuple, tuple] = {}
            self._ar_in_graph = None                  # None: untried, True / False: the collective can(not) be captured
            self.optimizer = "adam"                   # train.py:211-216: 'adam' | 'momentum' (set_optimizer)
            self.momentum = 0.0
            self.mom_accum = None                     # MomentumOptimizer's slot (`<var>/Momentum`), allocated on demand
            self.set_params(self.init_params(seed))

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.an3d_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ---- configuration --------------------------------------------------------------------
    PRECISIONS = {"fp32": _lib.PRECISION_FP32, "bf16": _lib.PRECISION_BF16, "bf16x3": _lib.PRECISION_BF16X3,
                  "bf16x6": _lib.PRECISION_BF16X6}

    def set_precision(self, precision: str) -> None:
        """'fp32': CUDA-core parity mode.  'bf16': tcgen05 fast mode (fused kernels for [64,128,C] conv stacks, the
        layer-by-layer tensor-core path for any other stack).  'bf16x3' / 'bf16x6': fp32-grade arithmetic on the tensor
        cores -- every GEMM operand split into two / three bf16 images (include/alignnet_b200.h)."""
        if precision not in self.PRECISIONS:
            raise ValueError("precision must be one of %s" % ", ".join(sorted(self.PRECISIONS)))
        self.precision = precision
        self.pflag = self.PRECISIONS[precision]

    # ---- parameters -----------------------------------------------------------------------
    def init_params(self, seed: int = 0) -> Dict[str, np.ndarray]:
        """Xavier-uniform weights (utils/tf_util.py:41-45; fans include the [1,3] window of the
        first conv), zero biases (:159), gamma=1 / beta=0 (:470-473)."""
        rng = np.random.Generator(np.random.PCG64(seed))
        out = {}
        for name in self.params_layout.order:
            _, shape = self.params_layout.entries[name]
            if name.endswith("/weights"):
                if len(shape) == 4:
                    receptive = shape[0] * shape[1]
                    fan_in, fan_out = shape[2] * receptive, shape[3] * receptive
                    mat = (shape[1] * shape[2], shape[3])
                else:
                    fan_in, fan_out = shape
                    mat = shape
                limit = math.sqrt(6.0 / (fan_in + fan_out))
                out[name] = rng.uniform(-limit, limit, size=mat).astype(np.float32)
            elif name.endswith("/gamma"):
                out[name] = np.ones(shape, np.float32)
            else:
                out[name] = np.zeros(shape, np.float32)
        return out

    def _flatten(self, layout: Layout, tensors: Dict[str, np.ndarray], strict: bool = True) -> np.ndarray:
        flat = np.zeros(layout.total, np.float32)
        for name, (off, shape) in layout.entries.items():
            if name not in tensors:
                if strict:
                    raise KeyError(f"missing tensor {name!r}")
                continue
            v = np.asarray(tensors[name], np.float32).reshape(-1)
            n = int(np.prod(shape))
            if v.size != n:
                raise ValueError(f"{name}: expected {n} elements ({shape}), got {v.size}")
            flat[off:off + n] = v
        return flat

    def _unflatten(self, layout: Layout, flat: np.ndarray) -> Dict[str, np.ndarray]:
        out = {}
        for name, (off, shape) in layout.entries.items():
            n = int(np.prod(shape))
            mat = (shape[1] * shape[2], shape[3]) if len(shape) == 4 else shape
            out[name] = flat[off:off + n].reshape(mat).copy()
        return out

    def params_changed(self) -> None:
        """Invalidate the cached inference-mode folds (see `cache_eval_weights`)."""
        self._pversion += 1

    def set_params(self, tensors: Dict[str, np.ndarray]) -> None:
        self._pversion += 1
        self.params.copy_(torch.from_numpy(self._flatten(self.params_layout, tensors)))

    def get_params(self) -> Dict[str, np.ndarray]:
        return self._unflatten(self.params_layout, self.params.cpu().numpy())

    def get_grads(self) -> Dict[str, np.ndarray]:
        return self._unflatten(self.params_layout, self.grads.cpu().numpy())

    def set_state(self, tensors: Dict[str, np.ndarray]) -> None:
        self._pversion += 1
        self.bn_state.copy_(torch.from_numpy(self._flatten(self.state_layout, tensors)))

    def get_state(self) -> Dict[str, np.ndarray]:
        return self._unflatten(self.state_layout, self.bn_state.cpu().numpy())

    # ---- buffers --------------------------------------------------------------------------
    def workspace_bytes(self, B: int, N: int, flags: int) -> int:
        out = C.c_int64()
        _lib.check(self.lib.an3d_workspace_bytes(self.ctx, B, N, flags, C.byref(out)), "an3d_workspace_bytes")
        return int(out.value)

    def _workspace(self, B: int, N: int, flags: int) -> torch.Tensor:
        key = (B, N, flags)
        ws = self._ws.get(key)
        if ws is None:
            # one live workspace: large shapes would otherwise pile up.  Captured graphs keep their own reference to
            # the workspace they baked in (see _capture), so eviction here never frees memory a replay touches; what
            # lives INSIDE the evicted workspace -- the cached inference folds -- is forgotten with it.
            self._ws.clear()
            self._prepared.clear()
            for k in [k for k in getattr(self, "_graphs", {}) if k[0] == "fwd-lean"]:
                del self._graphs[k]
            ws = torch.empty(self.workspace_bytes(B, N, flags) + 256, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def _outputs(self, B: int) -> Dict[str, torch.Tensor]:
        o = self._out.get(B)
        if o is None:
            nb2 = 2 * self.num_bins
            o = {k: torch.empty((B, nb2 if "logits" in k else 3), dtype=torch.float32, device=self.device)
                 for k in OUTPUT_KEYS}
            self._out = {B: o}
        return o

    @staticmethod
    def _out_struct(o: Dict[str, torch.Tensor]) -> _lib.Outputs:
        s = _lib.Outputs()
        for k in OUTPUT_KEYS:
            setattr(s, k, o[k].data_ptr())
        return s

    @staticmethod
    def _label_struct(labels: Dict[str, torch.Tensor]) -> _lib.Labels:
        s = _lib.Labels()
        for k in LABEL_KEYS:
            t = labels.get(k)
            setattr(s, k, None if t is None else t.data_ptr())
        return s

    def _check_input(self, t: torch.Tensor, shape) -> torch.Tensor:
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise TypeError("inputs must be contiguous float32 CUDA tensors")
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    # ---- compute --------------------------------------------------------------------------
    def forward(self, pcs1: torch.Tensor, pcs2: torch.Tensor, is_training: bool, bn_decay: Optional[float] = None,
                masks: Optional[Dict[str, torch.Tensor]] = None, seed: int = 0,
                seed_on_device: bool = False) -> Dict[str, torch.Tensor]:
        """get_model (models/tp8.py:135-158).  Returns the 8 end_points as CUDA tensors (buffers are
        reused between calls with the same batch size)."""
        B, N = int(pcs1.shape[0]), int(pcs1.shape[1])
        self._check_input(pcs1, (B, N, 3))
        self._check_input(pcs2, (B, N, 3))
        flags = self.pflag | (_lib.TRAINING if is_training else 0)
        ws = self._workspace(B, N, flags)
        call_flags = flags | (_lib.DETERMINISTIC if self.deterministic else 0)
        if is_training:
            self._pversion += 1                  # the moving averages change
        elif self.cache_eval_weights and self.pflag == _lib.PRECISION_BF16:
            if self._prepared.get((B, N, flags)) == self._pversion:
                call_flags |= _lib.WEIGHTS_PREPARED
            self._prepared[(B, N, flags)] = self._pversion
        out = self._outputs(B)
        ostruct = self._out_struct(out)
        d = _lib.Dropout()
        d.seed = int(seed)
        if seed_on_device:
            d.seed_dev = self.seed_dev.data_ptr()
        if masks is not None:
            for i, k in enumerate(MASK_KEYS):
                if k in masks and masks[k] is not None:
                    d.masks[i] = self._check_input(masks[k], masks[k].shape).data_ptr()
        decay = 0.9 if bn_decay is None else float(bn_decay)   # utils/tf_util.py:475
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.an3d_forward(self.ctx, self.params.data_ptr(), self.bn_state.data_ptr(), pcs1.data_ptr(),
                                         pcs2.data_ptr(), B, N, call_flags, decay, C.byref(d), C.byref(ostruct),
                                         ws.data_ptr(), ws.numel(), stream), "an3d_forward")
        self._last = (B, N, flags)
        return out

    def loss(self, labels: Dict[str, torch.Tensor], end_points: Dict[str, torch.Tensor]) -> torch.Tensor:
        """get_loss forward only (models/tp8.py:401-407).  Returns the 20-float loss vector."""
        B = int(end_points["pred_translations"].shape[0])
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ls, os_ = self._label_struct(labels), self._out_struct(end_points)
        scratch = torch.empty(64 * B + 1024, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.an3d_loss(self.ctx, C.byref(ls), C.byref(os_), B, self.loss_buf.data_ptr(),
                                      scratch.data_ptr(), scratch.numel() * 4, stream), "an3d_loss")
        return self.loss_buf

    def backward(self, pcs1, pcs2, labels: Dict[str, torch.Tensor], end_points: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Loss + gradients of the last training-mode forward into self.grads."""
        B, N, flags = self._last
        if not (flags & _lib.TRAINING):
            raise RuntimeError("backward() needs a preceding forward(is_training=True)")
        ws = self._workspace(B, N, flags)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ls, os_ = self._label_struct(labels), self._out_struct(end_points)
        _lib.check(self.lib.an3d_loss_backward(self.ctx, self.params.data_ptr(), pcs1.data_ptr(), pcs2.data_ptr(),
                                               C.byref(ls), C.byref(os_), B, N, flags, self.grads.data_ptr(),
                                               self.loss_buf.data_ptr(), ws.data_ptr(), ws.numel(), stream),
                   "an3d_loss_backward")
        return self.loss_buf

    def adam_step(self, lr: float, grad_scale: float = 1.0, beta1: float = 0.9, beta2: float = 0.999,
                  eps: float = 1e-8) -> None:
        """tf.train.AdamOptimizer(lr).minimize(..., global_step) (train.py:212-217)."""
        self.step += 1
        self._pversion += 1
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.an3d_adam_step(self.params.data_ptr(), self.grads.data_ptr(), self.adam_m.data_ptr(),
                                           self.adam_v.data_ptr(), self.params.numel(), lr, self.step, grad_scale,
                                           beta1, beta2, eps, stream), "an3d_adam_step")

    def set_optimizer(self, optimizer: str = "adam", momentum: Optional[float] = None) -> None:
        """cfg.training.optimizer.optimizer (train.py:211-216): 'adam' (tf.train.AdamOptimizer(lr)) or 'momentum'
        (tf.train.MomentumOptimizer(lr, momentum=cfg.training.optimizer.momentum)); anything else is the reference's
        `assert False, "Invalid optimizer"`."""
        if optimizer == "adam":
            self.optimizer = "adam"
            return
        if optimizer != "momentum":
            raise ValueError(f"Invalid optimizer {optimizer!r} (train.py:215)")
        if momentum is None:
            raise ValueError("the momentum optimiser needs cfg.training.optimizer.momentum (train.py:212)")
        self.optimizer, self.momentum = "momentum", float(momentum)
        if self.mom_accum is None:
            self.mom_accum = torch.zeros_like(self.params)

    def momentum_step(self, lr: float, grad_scale: float = 1.0, count_step: bool = True) -> None:
        """tf.train.MomentumOptimizer(lr, momentum).minimize(..., global_step) (train.py:211-212, 217)."""
        if self.mom_accum is None:
            raise RuntimeError("momentum_step before set_optimizer('momentum', momentum=...)")
        if count_step:
            self.step += 1
        self._pversion += 1
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.an3d_momentum_step(self.params.data_ptr(), self.grads.data_ptr(), self.mom_accum.data_ptr(),
                                               self.params.numel(), lr, self.momentum, grad_scale, stream),
                   "an3d_momentum_step")

    def optimizer_step(self, lr: float, grad_scale: float = 1.0) -> None:
        """train_op = optimizer.minimize(loss, global_step=batch) (train.py:217) with the configured optimiser."""
        if self.optimizer == "momentum":
            self.momentum_step(lr, grad_scale)
        else:
            self.adam_step(lr, grad_scale=grad_scale)

    def train_step(self, batch: Dict[str, torch.Tensor], lr: float, bn_decay: float, seed: Optional[int] = None,
                   masks=None, allreduce=None) -> torch.Tensor:
        """One `sess.run([train_op, loss, ...])` (train.py:368): forward, loss, backward,
        optional gradient all-reduce (callable), optimiser update (Adam unless set_optimizer chose momentum).  Returns the
        device loss vector."""
        ep = self.forward(batch["pcs1"], batch["pcs2"], True, bn_decay, masks, self.step if seed is None else seed)
        loss = self.backward(batch["pcs1"], batch["pcs2"], batch, ep)
        scale = 1.0
        if allreduce is not None:
            scale = allreduce(self.grads)
        self.optimizer_step(lr, grad_scale=scale)
        return loss

    # ---- CUDA-graph replay ----------------------------------------------------------------------
    # The step is ~250 dependent launches; replaying it as a CUDA graph removes the per-launch gaps on the GPU
    # timeline (c3: 8.3 -> 7.9 ms, c2: 0.83 -> 0.65 ms).  Everything that changes from step to step lives in
    # device memory: the batch (static buffers the caller refills), the global step and the dropout seed
    # (advanced by a one-thread kernel inside the graph).  lr and bn_decay are staircase schedules
    # (train.py:133-174): a new value simply captures a new graph.
    def _capture(self, key, fn):
        """Returns (graph, outputs, fresh).  The first call runs `fn` once for real (workspace allocation,
        kernel attributes) and then captures it; `fresh` tells the caller that this call's work is already done."""
        if key in self._graphs:
            g, out, _held = self._graphs[key]
            return g, out, False
        out = fn()
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        # the graph bakes in raw pointers of the workspace and output buffers live at capture time: hold them for
        # the graph's lifetime, whatever `_workspace` / `_outputs` evict later
        held = (list(self._ws.values()), list(self._out.values()))
        self._graphs[key] = (g, out, held)
        return g, out, True

    def forward_graph(self, pcs1: torch.Tensor, pcs2: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Eval-mode get_model replayed as a CUDA graph over the caller's STATIC input buffers (refill them in
        place between calls).  Returns the same 8 end_points buffers on every call."""
        lean = False
        if self.cache_eval_weights and self.pflag == _lib.PRECISION_BF16:
            B, N = int(pcs1.shape[0]), int(pcs1.shape[1])
            if self._prepared.get((B, N, self.pflag)) != self._pversion:
                return self.forward(pcs1, pcs2, False)        # parameters changed: one eager call re-derives the folds
            lean = True                                        # the replayed graph holds only the data-dependent launches
        key = ("fwd-lean" if lean else "fwd", pcs1.data_ptr(), pcs2.data_ptr(), tuple(pcs1.shape), self.pflag)
        g, out, fresh = self._capture(key, lambda: self.forward(pcs1, pcs2, False))
        if not fresh:
            g.replay()
        return out

    def _advance_step(self, seed_base: int = 0) -> None:
        """step_dev += 1; seed_dev = seed_base + step_dev (one thread, inside the captured graph).  `seed_base` is baked
        into the graph: data-parallel ranks pass rank << 40 so that the shards do not share dropout masks."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.an3d_step_advance(self.step_dev.data_ptr(), self.seed_dev.data_ptr(), int(seed_base), stream),
                   "an3d_step_advance")

    def _optimizer_step_dev(self, lr: float, grad_scale: float) -> None:
        """The update inside a captured step: Adam reads its step count from device memory; the momentum update carries
        none, so its ordinary entry point is captured as it is (the host counter is advanced by train_step_graph)."""
        if self.optimizer == "momentum":
            self.momentum_step(lr, grad_scale, count_step=False)
        else:
            self._adam_step_dev(lr, grad_scale)

    def _adam_step_dev(self, lr: float, grad_scale: float, beta1=0.9, beta2=0.999, eps=1e-8) -> None:
        self._pversion += 1
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.an3d_adam_step_dev(self.params.data_ptr(), self.grads.data_ptr(), self.adam_m.data_ptr(),
                                               self.adam_v.data_ptr(), self.params.numel(), lr, self.step_dev.data_ptr(),
                                               grad_scale, beta1, beta2, eps, stream), "an3d_adam_step_dev")

    def train_step_graph(self, batch: Dict[str, torch.Tensor], lr: float, bn_decay: float, allreduce=None,
                         seed_salt: int = 0) -> torch.Tensor:
        """`Engine.train_step` replayed as CUDA graphs over the caller's STATIC batch buffers.  Without an
        all-reduce the whole step is one graph; with one (data parallel) the graph is split around the collective:
        [advance step, forward, loss + backward] -> all-reduce (eager, NCCL) -> [Adam]."""
        if getattr(self, "_step_dev_shadow", None) != self.step:      # eager steps ran in between: resynchronise
            self.step_dev.fill_(self.step)
        ptrs = tuple(batch[k].data_ptr() for k in sorted(batch))
        shape = tuple(batch["pcs1"].shape)

        salt = int(seed_salt) << 40

        def fwd_bwd():
            self._advance_step(salt)
            ep = self.forward(batch["pcs1"], batch["pcs2"], True, bn_decay, None, seed_on_device=True)
            return self.backward(batch["pcs1"], batch["pcs2"], batch, ep)

        one_graph = allreduce is None or (self._ar_in_graph is not False and os.environ.get("AN3D_GRAPH_ALLREDUCE") != "0")
        if one_graph:
            # With a collective the NCCL kernel is captured into the step's graph as well (torch.distributed supports
            # capturing its NCCL collectives): one replay per step instead of graph / eager collective / graph, which
            # left two launch gaps on every rank's timeline (SCALE_r01: +0.09 ms at 2 ranks, +0.15 ms at 8).
            def whole():
                loss = fwd_bwd()
                scale = 1.0 if allreduce is None else float(allreduce(self.grads))
                self._optimizer_step_dev(lr, scale)
                return loss
            opt = (self.optimizer, self.momentum if self.optimizer == "momentum" else None)
            key = ("train" if allreduce is None else "train-ar", ptrs, shape, float(lr), float(bn_decay), self.pflag, salt, opt)
            try:
                g, loss, fresh = self._capture(key, whole)
                if allreduce is not None:
                    self._ar_in_graph = True
            except Exception:
                if allreduce is None or self._ar_in_graph:
                    raise
                self._ar_in_graph = False                     # this backend's collective cannot be captured: split path
                torch.cuda.synchronize(self.device)
                return self.train_step_graph(batch, lr, bn_decay, allreduce)
            if not fresh:
                g.replay()
        else:
            g1, loss, fresh = self._capture(("train-a", ptrs, shape, float(bn_decay), self.pflag, salt), fwd_bwd)
            if not fresh:
                g1.replay()
            scale = float(allreduce(self.grads))
            g2, _, fresh2 = self._capture(("train-b", float(lr), scale, self.optimizer, self.momentum),
                                          lambda: self._optimizer_step_dev(lr, scale))
            if not fresh2:
                g2.replay()
        self.step += 1
        self._pversion += 1
        self._step_dev_shadow = self.step
        return loss

    # ---- small utilities on the same ABI ----------------------------------------------------
    def decode_angles(self, logits: torch.Tensor, scaled) -> torch.Tensor:
        """scaled: True / 1 = tf_get_angles, False / 0 = classLogits2angle (host decoder), 2 = tf_classLogits2angle."""
        B = int(logits.shape[0])
        out = torch.empty(B, dtype=torch.float32, device=logits.device)
        stream = torch.cuda.current_stream(logits.device).cuda_stream
        _lib.check(self.lib.an3d_decode_angles(logits.data_ptr(), out.data_ptr(), B, self.num_bins, int(scaled), stream),
                   "an3d_decode_angles")
        return out

    def loss_p2p(self, pcs1: torch.Tensor, labels: Dict[str, torch.Tensor], end_points: Dict[str, torch.Tensor]) -> torch.Tensor:
        """a21: `_get_loss_p2p` (models/tp8.py:374-398) as the reference computes it (quirk Q6: the clouds collapse to
        the tiled rotation centres).  Returns the device vector [per_transform_loss, loss].  Forward value only."""
        B, N = int(pcs1.shape[0]), int(pcs1.shape[1])
        dec = lambda k: self.decode_angles(end_points[k], 2)              # tf_classLogits2angle
        ang = (dec("pred_pc2angle_logits") - dec("pred_pc1angle_logits") + dec("pred_remaining_angle_logits")).contiguous()
        rel = labels["rel_angles"].reshape(B, -1)[:, 0].contiguous()
        ws = torch.empty(2 * B * N * 3 + 4, dtype=torch.float32, device=pcs1.device)
        out = torch.empty(2, dtype=torch.float32, device=pcs1.device)
        stream = torch.cuda.current_stream(pcs1.device).cuda_stream
        _lib.check(self.lib.an3d_loss_p2p(pcs1.data_ptr(), end_points["pred_translations"].data_ptr(), ang.data_ptr(),
                                          end_points["pred_s2_pc1centers"].data_ptr(), labels["translations"].data_ptr(),
                                          rel.data_ptr(), labels["pc1_centers"].data_ptr(), B, N, out.data_ptr(),
                                          ws.data_ptr(), ws.numel() * 4, stream), "an3d_loss_p2p")
        return out

    def pred_angles(self, end_points: Dict[str, torch.Tensor]) -> torch.Tensor:
        """train.py:453-456: dec(pc2) - dec(pc1) + dec(remaining) with the host decoder's semantics."""
        a1 = self.decode_angles(end_points["pred_pc1angle_logits"], False)
        a2 = self.decode_angles(end_points["pred_pc2angle_logits"], False)
        ar = self.decode_angles(end_points["pred_remaining_angle_logits"], False)
        return a2 - a1 + ar


def rigid_apply(pts: torch.Tensor, translation=None, angle=None, center=None) -> torch.Tensor:
    """Batched get_mat_angle + transform_points (tp_utils/pointcloud.py:279-298) on the device."""
    lib = _lib.load()
    B, N = int(pts.shape[0]), int(pts.shape[1])
    out = torch.empty_like(pts)
    ptr = lambda t: None if t is None else t.data_ptr()
    stream = torch.cuda.current_stream(pts.device).cuda_stream
    _lib.check(lib.an3d_rigid_apply(pts.data_ptr(), ptr(translation), ptr(angle), ptr(center), out.data_ptr(), B, N,
                                    stream), "an3d_rigid_apply")
    return out


def transform_pcs(pcs: torch.Tensor, translations=None, angles=None, rotation_centers=None) -> torch.Tensor:
    """a21: `tf_transform_pcs` (models/tp8.py:361-371) exactly as coded, quirk Q6 included (every translate step REPLACES
    the cloud by the tiled translation).  The intended transform is `rigid_apply`."""
    lib = _lib.load()
    B, N = int(pcs.shape[0]), int(pcs.shape[1])
    out = torch.empty_like(pcs)
    ptr = lambda t: None if t is None else t.data_ptr()
    stream = torch.cuda.current_stream(pcs.device).cuda_stream
    _lib.check(lib.an3d_transform_pcs(pcs.data_ptr(), ptr(translations), ptr(angles), ptr(rotation_centers),
                                      out.data_ptr(), B, N, stream), "an3d_transform_pcs")
    return out


def recenter_translations(translations, angles, old_centers, new_centers) -> torch.Tensor:
    """translate_transform_to_new_center_of_rotation (tp_utils/pointcloud.py:309-318) on the device."""
    lib = _lib.load()
    out = torch.empty_like(translations)
    stream = torch.cuda.current_stream(translations.device).cuda_stream
    _lib.check(lib.an3d_recenter_translations(translations.data_ptr(), angles.data_ptr(), old_centers.data_ptr(),
                                              new_centers.data_ptr(), out.data_ptr(), int(translations.shape[0]),
                                              stream), "an3d_recenter_translations")
    return out
