/*
 * alignnet_b200.h -- C ABI of libalignnet_b200.so: the B200 (sm_100a) engine for the AlignNet-3D
 * `tp8` hot path (siamese 3-stage PointNet pose regressor: forward, loss, backward, TF-Adam, rigid
 * apply).
 *
 * The reference (grossjohannes/AlignNet-3D) is pure Python on TensorFlow 1.8 and has NO FFI; this
 * header is the boundary a maintainer would bind from `models/tp8.py` with ctypes (see
 * INTEGRATION.md).  Each entry point cites the reference interface it replaces (file:line into
 * the reference tree).
 *
 * Conventions
 *   - every function returns int: 0 = AN3D_OK, negative = AN3D_ERR_*; an3d_last_error() returns
 *     a thread-local message for the last failure.  No exceptions cross the boundary.
 *   - NO CPU FALLBACK: compute entry points fail with AN3D_ERR_NO_DEVICE / AN3D_ERR_ARCH when no
 *     sm_100 device is current.  Host-only queries (create, layout, workspace size) work anywhere.
 *   - the caller owns every buffer (device pointers unless stated otherwise): inputs, outputs,
 *     flat parameter / gradient / Adam buffers, BN shadow state, workspace.  All device pointers
 *     must be 16-byte aligned and contiguous (checked, AN3D_ERR_ALIGN).
 *   - calls are stream-ordered and asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - a ctx is host-only immutable model metadata; it is thread-compatible (no internal state is
 *     mutated by compute calls), one workspace per in-flight step.
 */
#ifndef ALIGNNET_B200_H_
#define ALIGNNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AN3D_VERSION 100

enum {
  AN3D_OK = 0,
  AN3D_ERR_INVALID = -1,      /* bad argument / shape / NULL pointer                       */
  AN3D_ERR_UNSUPPORTED = -2,  /* config the engine rejects (never a silent fallback)       */
  AN3D_ERR_NO_DEVICE = -3,    /* no CUDA device                                            */
  AN3D_ERR_ARCH = -4,         /* device is not sm_100 (B200)                               */
  AN3D_ERR_CUDA = -5,         /* CUDA runtime error (message in an3d_last_error)           */
  AN3D_ERR_ALIGN = -6,        /* pointer not 16-byte aligned                               */
  AN3D_ERR_WORKSPACE = -7     /* workspace too small                                       */
};

#define AN3D_MAX_LAYERS 8

/* Architecture = the `model` block of the reference's config JSON (configs/default.json:7-23;
 * read at models/tp8.py:98,108,115,130,154,307-308,318). */
typedef struct an3d_arch {
  int32_t num_bins;                       /* cfg.model.angles.num_bins                      */
  int32_t accept_inverted_angle;          /* cfg.model.angles.accept_inverted_angle         */
  float angle_factor;                     /* cfg.model.options.angle_factor                 */
  float early_stage_factor;               /* cfg.model.options.early_stage_factor           */
  int32_t n_conv[3];                      /* #conv layers of s1transformer / s2transformer / embedding */
  int32_t conv[3][AN3D_MAX_LAYERS];       /* their widths                                   */
  int32_t n_fc[3];                        /* #hidden FC layers of s1 / s2 / remaining_transform_prediction */
  int32_t fc[3][AN3D_MAX_LAYERS];         /* their widths (output layer is implied: 3 or 3+2*num_bins) */
  float keep_prob[3];                     /* dropout KEEP probability after the last hidden FC (tp8.py:81,86) */
} an3d_arch;

typedef struct an3d_ctx an3d_ctx;

/* Flags for an3d_forward / an3d_workspace_bytes. */
enum {
  AN3D_TRAINING = 1,        /* is_training=True: batch statistics, EMA update, dropout (tf_util.py:476-488,572) */
  AN3D_PRECISION_FP32 = 0,  /* CUDA-core fp32 everywhere: the <=1e-4 parity mode               */
  AN3D_PRECISION_BF16 = 2,  /* bf16 tcgen05 tensor-core GEMMs with fp32 accumulation (fast mode).  Conv stacks of the
                             * form [64, 128, C] run in the fused kernels; any other depth / width (models/tp8.py:49-59,
                             * e.g. configs/default.json:13-15) runs layer by layer through the split-operand GEMM
                             * below with ONE bf16 image per operand */
  /* fp32-grade arithmetic on the tensor cores: every GEMM operand enters as a sum of bf16 images (hi + lo, or
   * hi + mid + lo) and the tcgen05 products of the significant image pairs accumulate in one fp32 TMEM tile
   * (3 products: relative error ~2^-18 per product, held to the same tolerances as AN3D_PRECISION_FP32; 6 products:
   * ~2^-24).  Activations are materialised in fp32 as in the fp32 mode; statistics, pooling, BN and the loss are the
   * fp32 mode's kernels.  Any conv-stack depth. */
  AN3D_PRECISION_BF16X3 = 16,
  AN3D_PRECISION_BF16X6 = 32,
  /* Inference only (ignored with AN3D_TRAINING and in fp32 mode; not part of the workspace size).  The caller
   * asserts that the previous an3d_forward on THIS workspace ran in inference mode with the SAME params and
   * bn_state: the folded BN scales / shifts and the packed weight images it left in the workspace are reused and
   * the ~40 launches that derive them are skipped.  The library cannot check the assertion. */
  AN3D_WEIGHTS_PREPARED = 4,
  /* Bit-reproducible forward pass (not part of the workspace size).  Training-mode forwards always are: every
   * cross-CTA fp32 sum of the pass goes through per-CTA partial slots added in a fixed order, and no FC layer splits K
   * across CTAs.  Inference additionally splits K of under-filled FC launches with fp32 reductions (+9 % at B=1024);
   * this flag turns that off, so two inference calls on the same inputs return identical bits. */
  AN3D_DETERMINISTIC = 8
};

/* The eight tensors of end_points (models/tp8.py:146-156).  Centers/translations [B,3],
 * logits [B, 2*num_bins], fp32, row-major. */
typedef struct an3d_outputs {
  float* pred_s1_pc1centers;
  float* pred_s1_pc2centers;
  float* pred_s2_pc1centers;
  float* pred_s2_pc2centers;
  float* pred_pc1angle_logits;
  float* pred_pc2angle_logits;
  float* pred_translations;
  float* pred_remaining_angle_logits;
} an3d_outputs;

/* Labels = placeholders 3..8 of models/tp8.py:13-23 (translations [B,3], *_angles [B,1], ...). */
typedef struct an3d_labels {
  const float* translations;
  const float* rel_angles;    /* unused by the 'separate' loss, kept for signature parity */
  const float* pc1_centers;
  const float* pc2_centers;
  const float* pc1_angles;
  const float* pc2_angles;
} an3d_labels;

/* Dropout control.  masks[i] (optional, device, fp32 0/1 keep-masks [B, last hidden width]) for
 * i = s1/branch1, s1/branch2, s2/branch1, s2/branch2, head; NULL -> Philox-style counter RNG
 * keyed by `seed` (tf.nn.dropout at utils/tf_util.py:572-574). */
typedef struct an3d_dropout {
  uint64_t seed;
  const float* masks[5];
  const uint64_t* seed_dev; /* optional device pointer: when non-NULL the seed is read from device memory at
                               execution time (lets a captured CUDA graph draw fresh masks on every replay) */
} an3d_dropout;

/* ---- host-only ------------------------------------------------------------------------ */
int an3d_version(void);
const char* an3d_last_error(void);

/* Validates the architecture (rejects what v1 does not implement with AN3D_ERR_UNSUPPORTED) and
 * builds the flat parameter layout.  Replaces graph construction in get_model, models/tp8.py:135. */
int an3d_create(const an3d_arch* arch, an3d_ctx** out_ctx);
int an3d_destroy(an3d_ctx* ctx);

/* Flat-buffer layout.  which = 0: trainable parameters (fp32 buffer `params`, same layout for
 * `grads`, Adam `m`, `v`); which = 1: BN shadow state (`bn_state`).  Names are the TF variable
 * names of the reference graph (utils/tf_util.py:148-160,333-339,470-479).
 * Every tensor starts on a 16-byte boundary, so an3d_num_elements() may exceed the sum of the
 * tensor sizes by a few padding floats (which stay zero). */
int an3d_num_elements(const an3d_ctx* ctx, int which, int64_t* out_count);
int an3d_num_tensors(const an3d_ctx* ctx, int which, int32_t* out_count);
int an3d_tensor_info(const an3d_ctx* ctx, int which, int32_t index, char* name, int32_t name_capacity,
                     int64_t* offset, int32_t* ndim, int64_t shape[4]);

/* Bytes of caller-owned scratch needed for one step at (B, N) with the given flags. */
int an3d_workspace_bytes(const an3d_ctx* ctx, int32_t batch, int32_t num_points, int32_t flags, int64_t* out_bytes);

/* ---- device --------------------------------------------------------------------------- */

/* get_model (models/tp8.py:135-158) for both siamese branches + head.
 *   params    [num_elements(0)] fp32          bn_state [num_elements(1)] fp32 (updated in training)
 *   pcs1,pcs2 [B,N,3] fp32                    bn_decay: value of train.py:159-174 at this step
 *   dropout   may be NULL when !AN3D_TRAINING
 * In training mode the workspace keeps what an3d_loss_backward needs. */
int an3d_forward(const an3d_ctx* ctx, const float* params, float* bn_state, const float* pcs1, const float* pcs2,
                 int32_t batch, int32_t num_points, int32_t flags, float bn_decay, const an3d_dropout* dropout,
                 const an3d_outputs* out, void* workspace, int64_t workspace_bytes, void* stream);

/* get_loss -> _get_loss_separate (models/tp8.py:304-354, 401-407).  loss_out: device float[20]:
 * [0] per_transform_loss, [1] losses/translation, [2] losses/angle, [3..] the stage scalars of
 * tp8.py:339-353 in that order ([3..16]), rest zero. */
int an3d_loss(const an3d_ctx* ctx, const an3d_labels* labels, const an3d_outputs* out, int32_t batch,
              float* loss_out, void* workspace, int64_t workspace_bytes, void* stream);

/* Loss + full backward of the step last run by an3d_forward(AN3D_TRAINING) on `workspace`:
 * what optimizer.minimize(loss) differentiates (train.py:217).  grads: flat fp32, overwritten. */
int an3d_loss_backward(const an3d_ctx* ctx, const float* params, const float* pcs1, const float* pcs2,
                       const an3d_labels* labels, const an3d_outputs* out, int32_t batch, int32_t num_points,
                       int32_t flags, float* grads, float* loss_out, void* workspace, int64_t workspace_bytes,
                       void* stream);

/* tf.train.AdamOptimizer.apply_gradients (train.py:212-217) over the flat buffers:
 * g = grads*grad_scale; m,v EMAs; lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps).
 * step = t >= 1. */
int an3d_adam_step(float* params, const float* grads, float* m, float* v, int64_t count, float lr, int64_t step,
                   float grad_scale, float beta1, float beta2, float eps, void* stream);

/* tf_get_angles (models/tp8.py:294-301; scaled=1: residual * pi/nb, floor-mod wrap) or
 * classLogits2angle (tp8.py:229-244; scaled=0: unscaled residual, `if a > pi: a -= 2pi`). */
/* Device-resident step state for CUDA-graph replay of the training step (train.py:212-217 keeps the global
 * step in a tf.Variable for the same reason: it must advance inside the executed graph).
 *   an3d_step_advance: *step_dev += 1 ; *seed_dev = seed_base + *step_dev   (one thread)
 *   an3d_adam_step_dev: as an3d_adam_step with t = *step_dev read on the device (t >= 1). */
int an3d_step_advance(int64_t* step_dev, uint64_t* seed_dev, uint64_t seed_base, void* stream);
int an3d_adam_step_dev(float* params, const float* grads, float* m, float* v, int64_t count, float lr,
                       const int64_t* step_dev, float grad_scale, float beta1, float beta2, float eps, void* stream);

/* tf.train.MomentumOptimizer(lr, momentum=cfg.training.optimizer.momentum).apply_gradients (train.py:211-212; TF's
 * ApplyMomentum, use_nesterov = false): g = grads*grad_scale; accum = momentum*accum + g; p -= lr*accum.  `accum` is the
 * optimiser's one slot per variable (TF name `<var>/Momentum`), same flat layout as params.  Carries no step count, so the
 * same entry point serves eager steps and CUDA-graph replay. */
int an3d_momentum_step(float* params, const float* grads, float* accum, int64_t count, float lr, float momentum,
                       float grad_scale, void* stream);

/* scaled = 1: tf_get_angles (:294-301); 0: classLogits2angle (:229-244); 2: tf_classLogits2angle / tf_class2angle2
 * (:213-226, :248-251: unscaled residual, then tf.mod into [-pi, pi)) -- the decoder _get_loss_p2p uses. */
int an3d_decode_angles(const float* logits, float* angles, int32_t batch, int32_t num_bins, int32_t scaled,
                       void* stream);

/* get_mat_angle + transform_points (tp_utils/pointcloud.py:279-298), batched:
 * out[b,n,:] = Rz(angle[b]) (pts[b,n,:] - center[b]) + center[b] + translation[b].
 * translation / angle / center may be NULL (identity), as in the reference's defaults. */
int an3d_rigid_apply(const float* pts, const float* translation, const float* angle, const float* center,
                     float* out, int32_t batch, int32_t num_points, void* stream);

/* a21 -- tf_transform_pcs (models/tp8.py:361-371) EXACTLY as the reference codes it, quirk Q6 included: its helper
 * tf_translate_pcs (:357-358) RETURNS the tiled translation instead of adding it, so every translate step replaces the
 * cloud:   p = pcs ; if centers: p = -c ; if angles: p = p Rz(a)  (row vector times [[c,-s,0],[s,c,0],[0,0,1]], :26-27) ;
 *          if translations: p = -t ; if centers: p = c.
 * translations / angles / rotation_centers may be NULL (the reference's None).  The INTENDED rigid transform is
 * an3d_rigid_apply above; this entry point exists so that the unselected `p2p` loss is reproduced, not repaired. */
int an3d_transform_pcs(const float* pcs, const float* translations, const float* angles, const float* rotation_centers,
                       float* out, int32_t batch, int32_t num_points, void* stream);

/* a21 -- _get_loss_p2p (models/tp8.py:374-398) as the reference computes it: both clouds through an3d_transform_pcs,
 * tf.norm over the POINT axis (:386), squared, mean over [B,3]; the accept_inverted_angle variant (:388-393) is
 * identical to the plain one.  pred_angles = tf_classLogits2angle(pc2) - (pc1) + (remaining) (an3d_decode_angles with
 * scaled = 2).  loss_out[0] = per_transform_loss (= loss / B, what get_loss returns), loss_out[1] = loss.
 * workspace: at least 2*batch*num_points*3 floats + 16 bytes.  Forward value only: no shipped config selects this loss
 * (configs/default.json:52) and training with it is not implemented (an3d_loss_backward is the `separate` loss). */
int an3d_loss_p2p(const float* pcs1, const float* pred_translations, const float* pred_angles,
                  const float* pred_s2_pc1centers, const float* translations, const float* rel_angles,
                  const float* pc1_centers, int32_t batch, int32_t num_points, float* loss_out, void* workspace,
                  int64_t workspace_bytes, void* stream);

/* translate_transform_to_new_center_of_rotation (tp_utils/pointcloud.py:309-318):
 * out[i] = -d + Rz(angle[i]) d + t[i],  d = new_center[i] - old_center[i]. */
int an3d_recenter_translations(const float* translations, const float* angles, const float* old_centers,
                               const float* new_centers, float* out, int32_t count, void* stream);

/* Instrumentation for bench.py.  an3d_launch_count: kernels launched by this library since load.
 * an3d_profile_begin/end: while enabled, CUDA events are recorded on the launch stream around the
 * heavy kernels; _end synchronises and returns, per tag (host arrays of 8: 0 conv-stack layer-2
 * statistics pass, 1 conv-stack full pass, 2..4 backward kernels, 5 FC), the summed device
 * milliseconds and the number of launches. */
uint64_t an3d_launch_count(void);
int an3d_profile_begin(void);
int an3d_profile_end(float* ms_by_tag, int32_t* launches_by_tag);

/* Diagnostic (used by the test-suite, not by the hot path): one-CTA tcgen05 GEMM
 * d[128, n] = A * B^T over bf16 operands staged exactly like the production kernels stage them.
 * a_mn / b_mn = 0: operand given K-major (A [128,k], B [n,k] row-major); 1: MN-major (A [k,128],
 * B [k,n] row-major).  Pins the shared-memory / instruction descriptor conventions on hardware. */
int an3d_selftest_umma(const void* a_bf16, const void* b_bf16, float* d, int32_t n, int32_t k, int32_t a_mn,
                       int32_t b_mn, void* stream);

/* Batch assembly on the device (provider.py:60-71,85-136).  `points` holds the ragged
 * clouds of a batch back to back ([total, stride] floats, the first three columns are xyz, provider.py:125-126),
 * cloud b starting at row cloud_offset[b]; out[b, n, :] = points[cloud_offset[b] + sample_idx[b, n], :3]
 * (+ jitter[b, n, :] when given) -- the resample-with-replacement of provider.py:97-98 with the drawn indices, and
 * jitter_point_cloud (:60-71) with the drawn, already clipped noise.  sample_idx < 0 yields a zero point (the
 * reference substitutes zeros for an empty cloud, :97). */
int an3d_resample_gather(const float* points, const int64_t* cloud_offset, const int32_t* sample_idx, int32_t batch,
                         int32_t num_points, int32_t stride, const float* jitter, float* out, void* stream);

/* Yaw-constrained point-to-point ICP refinement (replaces o3.registration_icp with
 * TransformationEstimationPointToPoint(with_constraint=True) of icp.py:69-78, driven from train.py:463-484).  One CTA
 * per pair: nearest target point of every transformed source point within `radius` (brute force over the target
 * cloud staged through shared memory), closed-form yaw + translation update, `its` iterations or until fitness and
 * inlier RMSE both change by less than 1e-6.  src/tgt: ragged clouds back to back ([total,3] floats) with per-pair
 * row offsets and counts; init / out: [pairs,16] row-major 4x4 transforms; stats: [pairs,3] = fitness, inlier RMSE,
 * iterations run. */
int an3d_icp_yaw(const float* src, const int64_t* src_off, const int32_t* src_n, const float* tgt, const int64_t* tgt_off,
                 const int32_t* tgt_n, const float* init, int32_t pairs, float radius, int32_t its, float* out,
                 float* stats, void* stream);

/* Evaluation metrics on the device (evaluation.py:16-46,128-211).  For every predicted
 * transform: centre-of-rotation correction of the translation (pointcloud.py:309-318), xy translation error with the
 * 0.02 / 0.1 / 0.2 m levels, yaw error in degrees (optionally min with the 180-degree flip) with the 1 / 5 / 10
 * levels, joint levels; accumulated over the sets {all, val, test} (is_test: n bytes, NULL = all val) and the ranges
 * {all, 5 m, 10 m, 15 m, 20 m} of |gt_pc1center|.  Inputs are device double arrays ([n,3] / [n]); acc receives the
 * raw sums [3][5][14] (num, 3 translation levels, sum d_t, sum d_t^2, 3 angle levels, sum d_a, sum d_a^2, 3 joint
 * levels) and is zeroed by the call; the means / RMS of eval.json are a division on the host. */
int an3d_evaluate(const double* pred_translations, const double* pred_angles, const double* pred_centers,
                  const double* gt_translations, const double* gt_angles, const double* gt_pc1centers,
                  const uint8_t* is_test, int32_t n, int32_t accept_inverted_angle, double* acc, void* stream);

/* Diagnostic micro-benchmark: `ctas` CTAs each issue `iters` tcgen05 MMAs of shape 128 x n x 16 (bf16, operands in
 * the given majors, un-swizzled plane layout over a k-deep resident tile); cycles_per_mma_dev[cta] receives the
 * measured SM cycles per MMA.  Used to choose operand layouts, not on the hot path. */
int an3d_bench_umma(int32_t n, int32_t k, int32_t a_mn, int32_t b_mn, int32_t iters, int32_t ctas,
                    float* cycles_per_mma_dev, void* stream);

/* Diagnostic (test-suite only): training-mode forward + backward of ONE conv stack (stage 0..2) of
 * ONE branch on the bf16 tensor-core path with a caller-supplied upstream gradient dG [B, C3].
 * g_out [B, C3] receives the pooled feature, grads (flat, zeroed first) the parameter gradients of
 * that stack, dcenter [B,3] / dangle [B] the gradient w.r.t. the recentring / canonicalisation. */
int an3d_selftest_conv_stack(const an3d_ctx* ctx, const float* params, float* bn_state, int32_t stage, int32_t branch,
                             const float* pcs, const float* center, const float* angle, int32_t batch,
                             int32_t num_points, const float* dG, float* g_out, float* grads, float* dcenter,
                             float* dangle, void* workspace, int64_t workspace_bytes, void* stream);

/* Diagnostic (test-suite only): the tcgen05 GEMM that runs every FC layer of the bf16 mode (forward, wgrad and
 * dgrad forms; replaces tf.matmul of utils/tf_util.py:337 and its gradients).
 *   C[i,j] (+)= sum_k A(i,k) B(j,k) (+ bias[j]);  a_mn/b_mn = 0: operand[idx*ld + k], 1: operand[k*ld + idx];
 * optional prologue on A: relu(a*scale[ch] + shift[ch]) * mask * mask_scale; ksplit > 1 or accumulate != 0
 * reduces into C (pre-zeroed by the caller); stat_sum/stat_sq [N] (double, pre-zeroed) receive the column
 * sums of C and C^2 (ksplit must be 1).  Operands are fp32, rounded to bf16 on load, fp32 accumulation. */
int an3d_selftest_fc_gemm(const float* a, int64_t lda, int32_t a_mn, const float* b, int64_t ldb, int32_t b_mn, float* c,
                          int64_t ldc, int32_t m, int32_t n, int32_t k, const float* bias, const float* pro_scale,
                          const float* pro_shift, const float* pro_mask, float pro_mask_scale, int32_t ksplit,
                          int32_t accumulate, double* stat_sum, double* stat_sq, void* stream);

/* Diagnostic (test-suite only): the split-operand tcgen05 GEMM of the materialised path (AN3D_PRECISION_BF16X3 / _BF16X6,
 * and AN3D_PRECISION_BF16 for conv stacks the fused kernels do not cover; replaces tf.nn.conv2d of
 * utils/tf_util.py:157 / tf.matmul of :337 and their gradients).  Same operand conventions as an3d_selftest_fc_gemm;
 * nsplit = 1, 2 or 3 bf16 images per operand (1, 3 or 6 tensor-core products).  accumulate != 0 adds into C; K slicing
 * is the library's choice (C is cleared by the call when it slices and accumulate == 0). */
int an3d_selftest_split_gemm(const float* a, int64_t lda, int32_t a_mn, const float* b, int64_t ldb, int32_t b_mn, float* c,
                             int64_t ldc, int32_t m, int32_t n, int32_t k, const float* bias, const float* pro_scale,
                             const float* pro_shift, const float* pro_mask, float pro_mask_scale, int32_t nsplit,
                             int32_t accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ALIGNNET_B200_H_ */
