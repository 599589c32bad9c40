// Diagnostic entry (test-suite only): forward + backward of ONE conv stack of ONE branch on the
// bf16 tensor-core path with a caller-supplied upstream gradient, so the backward kernels can be
// checked against autograd without the (discontinuous) rest of the model in between.
#include "bf16_path.cuh"
#include "fc2_gemm.cuh"
#include "gemm_tc.cuh"

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
  AN3D_TRY(bf16_supported(m));           // this entry exercises the FUSED kernels: [64, 128, C] stacks only
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

extern "C" int an3d_selftest_fc_gemm(const float* a, int64_t lda, int32_t a_mn, const float* b, int64_t ldb, int32_t b_mn,
                                     float* c, int64_t ldc, int32_t m, int32_t n, int32_t k, const float* bias,
                                     const float* pro_scale, const float* pro_shift, const float* pro_mask,
                                     float pro_mask_scale, int32_t ksplit, int32_t accumulate, double* stat_sum,
                                     double* stat_sq, void* stream) {
  if (!a || !b || !c || m <= 0 || n <= 0 || k <= 0 || ksplit < 1 || (stat_sum && ksplit > 1) || (!stat_sum != !stat_sq) ||
      (!pro_scale != !pro_shift)) {
    set_error("an3d_selftest_fc_gemm: bad argument");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_device());
  // the caller's fp32 operands (either major, optional BN + ReLU + mask prologue on A) are packed into the bf16 plane-major
  // images the GEMM consumes, exactly as the model code does; the temporaries live for the duration of this call
  cudaStream_t st = (cudaStream_t)stream;
  const int a_rows = a_mn ? k : m, a_cols = a_mn ? m : k, b_rows = b_mn ? k : n, b_cols = b_mn ? n : k;
  __nv_bfloat16 *ia = nullptr, *ib = nullptr;
  AN3D_CUDA_CHECK(cudaMalloc(&ia, sizeof(__nv_bfloat16) * fc_image_elems(a_rows, a_cols)));
  AN3D_CUDA_CHECK(cudaMalloc(&ib, sizeof(__nv_bfloat16) * fc_image_elems(b_rows, b_cols)));
  fc2::PackArgs pa, pb;
  pa.src = a; pa.ld = lda; pa.rows = a_rows; pa.cols = a_cols; pa.scale = pro_scale; pa.shift = pro_shift; pa.mask = pro_mask;
  pa.mask_scale = pro_mask_scale; pa.dst = ia;
  pb.src = b; pb.ld = ldb; pb.rows = b_rows; pb.cols = b_cols; pb.dst = ib;
  int rc = fc2::pack(pa, st);
  if (rc == AN3D_OK) rc = fc2::pack(pb, st);
  fc2::Params f;
  f.A.g = ia; f.A.rows = a_rows; f.A.cols = a_cols; f.a_mn = a_mn; f.B.g = ib; f.B.rows = b_rows; f.B.cols = b_cols; f.b_mn = b_mn;
  f.C = c; f.ldc = ldc; f.M = m; f.N = n; f.K = k; f.bias = bias; f.ksplit = ksplit; f.accumulate = accumulate;
  f.stat_sum = stat_sum; f.stat_sq = stat_sq;
  if (rc == AN3D_OK) rc = fc2::launch(f, st);
  cudaStreamSynchronize(st);
  cudaFree(ia); cudaFree(ib);
  return rc;
}

extern "C" int an3d_selftest_split_gemm(const float* a, int64_t lda, int32_t a_mn, const float* b, int64_t ldb, int32_t b_mn,
                                        float* c, int64_t ldc, int32_t m, int32_t n, int32_t k, const float* bias,
                                        const float* pro_scale, const float* pro_shift, const float* pro_mask,
                                        float pro_mask_scale, int32_t nsplit, int32_t accumulate, void* stream) {
  if (!a || !b || !c || m <= 0 || n <= 0 || k <= 0 || nsplit < 1 || nsplit > tcg::kMaxSplit || (!pro_scale != !pro_shift)) {
    set_error("an3d_selftest_split_gemm: bad argument");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_device());
  cudaStream_t st = (cudaStream_t)stream;
  const int a_rows = a_mn ? k : m, a_cols = a_mn ? m : k, b_rows = b_mn ? k : n, b_cols = b_mn ? n : k;
  const int64_t ea = fc_image_elems(a_rows, a_cols), eb = fc_image_elems(b_rows, b_cols);
  __nv_bfloat16 *ia = nullptr, *ib = nullptr;
  AN3D_CUDA_CHECK(cudaMalloc(&ia, sizeof(__nv_bfloat16) * ea * nsplit));
  AN3D_CUDA_CHECK(cudaMalloc(&ib, sizeof(__nv_bfloat16) * eb * nsplit));
  tcg::PackArgs pa, pb;
  pa.src = a; pa.ld = lda; pa.rows = a_rows; pa.cols = a_cols; pa.scale = pro_scale; pa.shift = pro_shift; pa.mask = pro_mask;
  pa.mask_scale = pro_mask_scale; pa.nsplit = nsplit;
  pb.src = b; pb.ld = ldb; pb.rows = b_rows; pb.cols = b_cols; pb.nsplit = nsplit;
  for (int s = 0; s < nsplit; ++s) { pa.dst[s] = ia + s * ea; pb.dst[s] = ib + s * eb; }
  tcg::Params f;
  int rc = tcg::pack(pa, st, &f.A);
  if (rc == AN3D_OK) rc = tcg::pack(pb, st, &f.B);
  f.a_mn = a_mn; f.b_mn = b_mn; f.C = c; f.ldc = ldc; f.M = m; f.N = n; f.K = k; f.bias = bias; f.accumulate = accumulate;
  if (rc == AN3D_OK) rc = tcg::launch(f, st);
  cudaStreamSynchronize(st);
  cudaFree(ia); cudaFree(ib);
  return rc;
}
