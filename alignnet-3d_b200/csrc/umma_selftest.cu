// Single-CTA tcgen05 GEMM used by the test-suite to pin the descriptor conventions of umma.cuh on
// real hardware: D[128, N] = A * B^T with bf16 operands staged in the plane layout, for both
// operand majors.  Not on the hot path.
#include "common.cuh"
#include "umma.cuh"

namespace an3d {
namespace {

using namespace umma;

// A_g: a_mn ? [K,128] row-major : [128,K] row-major.   B_g: b_mn ? [K,N] row-major : [N,K] row-major.
__global__ void __launch_bounds__(160) umma_selftest_kernel(const __nv_bfloat16* A_g, const __nv_bfloat16* B_g,
                                                            float* D_g, int N, int K, int a_mn, int b_mn) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  // planes: operand X has `cols/8` planes of `rows` 16-byte chunks (rows = global row count of the tile)
  const int a_rows = a_mn ? K : 128, a_planes = (a_mn ? 128 : K) / 8;
  const int b_rows = b_mn ? K : N, b_planes = (b_mn ? N : K) / 8;
  const uint32_t a_plane = a_rows * 16 + 16, b_plane = b_rows * 16 + 16;  // +16: odd multiple of 16B to spread banks
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_planes * a_plane + 127) & ~127u);
  const int tid = threadIdx.x, warp = uniform_warp_idx();
  if (tid == 0) {
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&tmem_base_s, 256);
  // stage operands: 16-byte chunk (row r, chunk column c) -> plane c, row r
  const int a_cols = a_mn ? 128 : K, b_cols = b_mn ? N : K;
  for (int i = tid; i < a_rows * a_planes; i += blockDim.x) {
    const int r = i / a_planes, c = i % a_planes;
    *reinterpret_cast<uint4*>(sa + c * a_plane + r * 16) =
        *reinterpret_cast<const uint4*>(A_g + (size_t)r * a_cols + c * 8);
  }
  for (int i = tid; i < b_rows * b_planes; i += blockDim.x) {
    const int r = i / b_planes, c = i % b_planes;
    *reinterpret_cast<uint4*>(sb + c * b_plane + r * 16) =
        *reinterpret_cast<const uint4*>(B_g + (size_t)r * b_cols + c * 8);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 4) {
    const uint32_t idesc = make_idesc(128, N, a_mn, b_mn);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t ad, bd;
      if (a_mn) ad = make_desc(smem_u32(sa) + ks * 256, 128, a_plane);        // 16 K rows = 256 B
      else ad = make_desc(smem_u32(sa) + ks * 2 * a_plane, a_plane, 128);     // 2 K chunks = 2 planes
      if (b_mn) bd = make_desc(smem_u32(sb) + ks * 256, 128, b_plane);
      else bd = make_desc(smem_u32(sb) + ks * 2 * b_plane, b_plane, 128);
      mma_bf16(tmem, ad, bd, idesc, ks > 0);
    }
    mma_commit(&bar_done);
  }
  if (warp < 4) {
    mbar_wait(&bar_done, 0);
    tc_fence_after();
    const int row = tid;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D_g[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

// Micro-benchmark: every CTA issues `iters` MMAs of shape 128 x N x 16 over a resident tile (operands in either
// major, un-swizzled plane layout, plane pitch as in the production kernels) and reports the tensor-pipe cycles
// per MMA.  Used to choose operand layouts (profiles/r1_umma_microbench.txt).
__global__ void __launch_bounds__(64) umma_bench_kernel(int N, int K, int a_mn, int b_mn, int iters, float* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  const int a_rows = a_mn ? K : 128, a_planes = (a_mn ? 128 : K) / 8;
  const int b_rows = b_mn ? K : N, b_planes = (b_mn ? N : K) / 8;
  const uint32_t a_plane = a_rows * 16 + 16, b_plane = b_rows * 16 + 16;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_planes * a_plane + 127) & ~127u);
  const uint32_t total = ((a_planes * a_plane + 127) & ~127u) + b_planes * b_plane;
  for (uint32_t i = threadIdx.x * 16; i < total; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  const int warp = uniform_warp_idx();
  if (threadIdx.x == 0) { mbar_init(&bar_done, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(&tmem_base_s, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(128, N, a_mn, b_mn);
    const int nks = K / 16;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      // descriptors precomputed: the loop body is the MMAs only (an address computation per MMA would be timed instead)
      const uint64_t a0 = a_mn ? make_desc(smem_u32(sa), 128, a_plane) : make_desc(smem_u32(sa), a_plane, 128);
      const uint64_t b0 = b_mn ? make_desc(smem_u32(sb), 128, b_plane) : make_desc(smem_u32(sb), b_plane, 128);
      const uint32_t astep = a_mn ? 256 : 2 * a_plane, bstep = b_mn ? 256 : 2 * b_plane;
      uint64_t ad[8], bd[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { ad[j] = desc_advance(a0, (j % nks) * astep); bd[j] = desc_advance(b0, (j % nks) * bstep); }
      t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mma_bf16_raw(tmem, ad[j], bd[j], idesc, (i + j) > 0);
      }
      mma_commit_raw(&bar_done);
    }
    __syncwarp();
    mbar_wait(&bar_done, 0);
    t1 = clock64();
    t0 = __shfl_sync(0xffffffffu, t0, 0);   // elected lane is lane 0 of a converged warp
    if (threadIdx.x == 0) out[blockIdx.x] = (float)(t1 - t0) / (float)iters;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

}  // namespace
}  // namespace an3d

extern "C" int an3d_bench_umma(int32_t n, int32_t k, int32_t a_mn, int32_t b_mn, int32_t iters, int32_t ctas,
                               float* cycles_per_mma_dev, void* stream) {
  using namespace an3d;
  if (!cycles_per_mma_dev || n < 16 || n > 256 || (n % 16) || k < 16 || (k % 16) || k > 256 || iters < 1 || ctas < 1) {
    set_error("an3d_bench_umma: bad argument");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_device());
  const int a_rows = a_mn ? k : 128, a_planes = (a_mn ? 128 : k) / 8;
  const int b_rows = b_mn ? k : n, b_planes = (b_mn ? n : k) / 8;
  const size_t bytes = ((a_planes * (a_rows * 16 + 16) + 127) & ~127u) + b_planes * (b_rows * 16 + 16) + 256;
  AN3D_CUDA_CHECK(cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  umma_bench_kernel<<<ctas, 64, bytes, (cudaStream_t)stream>>>(n, k, a_mn, b_mn, iters, cycles_per_mma_dev);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

extern "C" int an3d_selftest_umma(const void* a_bf16, const void* b_bf16, float* d, int32_t n, int32_t k, int32_t a_mn,
                                  int32_t b_mn, void* stream) {
  using namespace an3d;
  if (!a_bf16 || !b_bf16 || !d || n < 16 || n > 256 || (n % 16) || k < 16 || (k % 16) || k > 256) {
    set_error("an3d_selftest_umma: need 16 <= n <= 256 (multiple of 16) and 16 <= k <= 256 (multiple of 16)");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_device());
  const int a_rows = a_mn ? k : 128, a_planes = (a_mn ? 128 : k) / 8;
  const int b_rows = b_mn ? k : n, b_planes = (b_mn ? n : k) / 8;
  const size_t bytes = ((a_planes * (a_rows * 16 + 16) + 127) & ~127u) + b_planes * (b_rows * 16 + 16) + 256;
  AN3D_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  umma_selftest_kernel<<<1, 160, bytes, (cudaStream_t)stream>>>((const __nv_bfloat16*)a_bf16,
                                                                (const __nv_bfloat16*)b_bf16, d, n, k, a_mn, b_mn);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}
