// Diagnostic entry (test-suite only): forward + backward of ONE conv stack of ONE branch on the
// bf16 tensor-core path with a caller-supplied upstream gradient, so the backward kernels can be
// checked against autograd without the (discontinuous) rest of the model in between.
#include "bf16_path.cuh"

using namespace an3d;

extern "C" int an3d_selftest_conv_stack(const an3d_ctx* ctx, const float* params, float* bn_state, int32_t stage,
                                        int32_t branch, const float* pcs, const float* center, const float* angle,
                                        int32_t batch, int32_t num_points, const float* dG, float* g_out, float* grads,
                                        float* dcenter, float* dangle, void* workspace, int64_t workspace_bytes,
                                        void* stream) {
  if (!ctx || !params || !bn_state || !pcs || !center || !dG || !g_out || !grads || !dcenter || !workspace || stage < 0 ||
      stage > 2 || branch < 0 || branch > 1) {
    set_error("an3d_selftest_conv_stack: bad argument");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_device());
  const Model& m = ctx->impl.model;
  cudaStream_t st = (cudaStream_t)stream;
  const int flags = AN3D_TRAINING | AN3D_PRECISION_BF16;
  PlanF32 p;
  AN3D_TRY(plan_f32(m, batch, num_points, flags, workspace, &p));
  if (p.bytes > workspace_bytes) {
    set_error("workspace too small: need %lld bytes", (long long)p.bytes);
    return AN3D_ERR_WORKSPACE;
  }
  AN3D_TRY(pack_weights_bf16(m, p, params, st));
  AN3D_TRY(pack_weights_bf16_bwd(m, p, params, st));
  AN3D_TRY(conv_stack_forward_bf16(m, p, stage, branch, pcs, center, angle, params, bn_state, true, 0.5f, st));
  const int C3 = m.conv[stage].back().cout;
  const int64_t ldg = stage == EMB ? 2 * C3 : C3;
  AN3D_CUDA_CHECK(cudaMemcpy2DAsync(g_out, C3 * sizeof(float), p.g[stage][branch], ldg * sizeof(float), C3 * sizeof(float),
                                    batch, cudaMemcpyDeviceToDevice, st));
  AN3D_CUDA_CHECK(cudaMemsetAsync(grads, 0, sizeof(float) * m.n_trainable, st));
  AN3D_CUDA_CHECK(cudaMemsetAsync(dcenter, 0, sizeof(float) * 3 * batch, st));
  if (dangle) AN3D_CUDA_CHECK(cudaMemsetAsync(dangle, 0, sizeof(float) * batch, st));
  return conv_stack_backward_bf16(m, p, stage, branch, pcs, center, angle, dG, C3, params, grads, true, dcenter,
                                  angle ? dangle : nullptr, st);
}
