// extern "C" entry points for forward / loss / backward: argument validation and dispatch on the
// precision flag.  No CPU fallback anywhere: a missing or non-sm_100 device is an error.
#include "common.cuh"
#include "loss.cuh"

namespace an3d {
int forward_impl(const Model& m, const float* params, float* state, const float* pcs1, const float* pcs2, int B, int N,
                 int flags, float bn_decay, const an3d_dropout* dropout, const an3d_outputs* out, void* workspace,
                 int64_t workspace_bytes, cudaStream_t st);
int backward_impl(const Model& m, const float* params, const float* pcs1, const float* pcs2, const an3d_labels* labels,
                  const an3d_outputs* out, int B, int N, int flags, float* grads, float* loss_out, void* workspace,
                  int64_t workspace_bytes, cudaStream_t st);
}  // namespace an3d

using namespace an3d;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int check_outputs(const an3d_outputs* o) {
  if (!o) {
    set_error("outputs struct is NULL");
    return AN3D_ERR_INVALID;
  }
  const void* ptrs[8] = {o->pred_s1_pc1centers, o->pred_s1_pc2centers, o->pred_s2_pc1centers, o->pred_s2_pc2centers,
                         o->pred_pc1angle_logits, o->pred_pc2angle_logits, o->pred_translations,
                         o->pred_remaining_angle_logits};
  for (int i = 0; i < 8; ++i) {
    if (!ptrs[i]) {
      set_error("output pointer %d is NULL", i);
      return AN3D_ERR_INVALID;
    }
    if (!aligned16(ptrs[i])) {
      set_error("output pointer %d is not 16-byte aligned", i);
      return AN3D_ERR_ALIGN;
    }
  }
  return AN3D_OK;
}

static int check_labels(const an3d_labels* l) {
  if (!l || !l->translations || !l->pc1_centers || !l->pc2_centers || !l->pc1_angles || !l->pc2_angles) {
    set_error("labels struct has NULL members (rel_angles may be NULL, the others may not)");
    return AN3D_ERR_INVALID;
  }
  return AN3D_OK;
}

extern "C" {

int an3d_workspace_bytes(const an3d_ctx* ctx, int32_t batch, int32_t num_points, int32_t flags, int64_t* out_bytes) {
  if (!ctx || !out_bytes) {
    set_error("an3d_workspace_bytes: NULL argument");
    return AN3D_ERR_INVALID;
  }
  PlanF32 p;
  AN3D_TRY(plan_f32(ctx->impl.model, batch, num_points, flags, nullptr, &p));
  *out_bytes = p.bytes;
  return AN3D_OK;
}

int an3d_forward(const an3d_ctx* ctx, const float* params, float* bn_state, const float* pcs1, const float* pcs2,
                 int32_t batch, int32_t num_points, int32_t flags, float bn_decay, const an3d_dropout* dropout,
                 const an3d_outputs* out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!ctx || !params || !bn_state || !pcs1 || !pcs2 || !workspace) {
    set_error("an3d_forward: NULL argument");
    return AN3D_ERR_INVALID;
  }
  if (batch < 1 || num_points < 1) {
    set_error("an3d_forward: batch=%d num_points=%d must be >= 1", batch, num_points);
    return AN3D_ERR_INVALID;
  }
  if (!aligned16(params) || !aligned16(bn_state) || !aligned16(pcs1) || !aligned16(pcs2) || !aligned16(workspace)) {
    set_error("an3d_forward: params / bn_state / pcs / workspace must be 16-byte aligned");
    return AN3D_ERR_ALIGN;
  }
  AN3D_TRY(check_outputs(out));
  AN3D_TRY(check_device());
  cudaStream_t st = (cudaStream_t)stream;
  return forward_impl(ctx->impl.model, params, bn_state, pcs1, pcs2, batch, num_points, flags, bn_decay, dropout, out,
                     workspace, workspace_bytes, st);
}

int an3d_loss(const an3d_ctx* ctx, const an3d_labels* labels, const an3d_outputs* out, int32_t batch, float* loss_out,
              void* workspace, int64_t workspace_bytes, void* stream) {
  if (!ctx || !loss_out || !workspace || batch < 1) {
    set_error("an3d_loss: bad argument");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_labels(labels));
  AN3D_TRY(check_outputs(out));
  AN3D_TRY(check_device());
  const int64_t need = loss_scratch_floats(batch) * (int64_t)sizeof(float);
  if (workspace_bytes < need) {
    set_error("an3d_loss: workspace too small: need %lld bytes", (long long)need);
    return AN3D_ERR_WORKSPACE;
  }
  return run_loss(ctx->impl.model, labels, out, batch, loss_out, static_cast<float*>(workspace), nullptr,
                  (cudaStream_t)stream);
}

int an3d_loss_backward(const an3d_ctx* ctx, const float* params, const float* pcs1, const float* pcs2,
                       const an3d_labels* labels, const an3d_outputs* out, int32_t batch, int32_t num_points,
                       int32_t flags, float* grads, float* loss_out, void* workspace, int64_t workspace_bytes,
                       void* stream) {
  if (!ctx || !params || !pcs1 || !pcs2 || !grads || !loss_out || !workspace || batch < 1 || num_points < 1) {
    set_error("an3d_loss_backward: bad argument");
    return AN3D_ERR_INVALID;
  }
  if (!aligned16(params) || !aligned16(grads) || !aligned16(workspace)) {
    set_error("an3d_loss_backward: params / grads / workspace must be 16-byte aligned");
    return AN3D_ERR_ALIGN;
  }
  AN3D_TRY(check_labels(labels));
  AN3D_TRY(check_outputs(out));
  AN3D_TRY(check_device());
  cudaStream_t st = (cudaStream_t)stream;
  return backward_impl(ctx->impl.model, params, pcs1, pcs2, labels, out, batch, num_points, flags, grads, loss_out,
                       workspace, workspace_bytes, st);
}

}  // extern "C"
