// Generic tcgen05 GEMM for the FC layers of the bf16 path.
//
//   C[i, j] (+)= sum_k A(i,k) * B(j,k)  (+ bias[j])        tile 128 x 128, K in blocks of 64
//
// Operands are fp32 row-major matrices in global memory; loader warps convert them to bf16 on the
// fly (optionally applying the producing layer's BN affine + ReLU + dropout mask to A, so FC
// activations are never materialised) and stage them in the un-swizzled plane layout of umma.cuh.
// Each operand is either "K-major" (global rows = M/N index, contiguous dim = K) or "MN-major"
// (global rows = K index, contiguous dim = M/N) -- the same loader serves both, only the
// descriptors differ -- which covers forward (X W), wgrad (X^T dZ, split-K with reductions) and
// dgrad (dZ W^T) without any transposed copies.
#pragma once
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace an3d {
namespace fcgemm {

using namespace umma;

struct Params {
  const float* A; int64_t lda; int a_mn;    // a_mn = 0: A[i*lda + k]   1: A[k*lda + i]
  const float* B; int64_t ldb; int b_mn;    // b_mn = 0: B[j*ldb + k]   1: B[k*ldb + j]
  float* C; int64_t ldc;
  int M, N, K;                              // extents of i, j, k (N and the contiguous dims multiples of 8)
  const float* bias;                        // [N] or nullptr
  const float* pro_scale;                   // optional prologue on A: relu(a*scale[ch] + shift[ch]) (* mask * mask_scale)
  const float* pro_shift;
  const float* pro_mask;                    // same layout as A
  float pro_mask_scale;
  int ksplit;                               // gridDim.z; > 1 -> accumulate with reductions into pre-zeroed C
  int accumulate;                           // reductions even with ksplit == 1
};

constexpr int kThreads = 288;               // 4 epilogue warps, 4 loader warps, 1 MMA warp
constexpr int kStages = 4;
constexpr uint32_t kPlaneK = 128 * 16 + 16;   // K-major tile: 8 planes x 128 rows
constexpr uint32_t kPlaneMN = 64 * 16 + 16;   // MN-major tile: 16 planes x 64 rows
constexpr uint32_t kTileBytes = 16 * kPlaneMN > 8 * kPlaneK ? 16 * kPlaneMN : 8 * kPlaneK;
constexpr uint32_t kTileStride = (kTileBytes + 127) & ~127u;
constexpr size_t kSmemBytes = 2 * kStages * (size_t)kTileStride + 256;

struct Bars {
  uint64_t full[kStages], empty[kStages], done;
  uint32_t tmem_base;
};

// Staging of one 64-K-block of an operand, split in two phases so that the global loads of BOTH
// operands are in flight together (one latency period per K block instead of two).
struct TileRegs { float4 v[16]; };

__device__ __forceinline__ void tile_issue(TileRegs& R, const float* src, int64_t ld, int mn_major, int mn0, int mn_ext,
                                           int k0, int k_end, int t) {
  const int ncc = mn_major ? 16 : 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = t + 128 * i;
    const int row = q / ncc, cc = q - row * ncc;
    const int g_row = mn_major ? k0 + row : mn0 + row;
    const int g_col = mn_major ? mn0 + cc * 8 : k0 + cc * 8;
    const bool ok = mn_major ? (g_row < k_end && g_col < mn_ext) : (g_row < mn_ext && g_col < k_end);
    if (ok) {
      const float* p = src + (int64_t)g_row * ld + g_col;
      R.v[2 * i] = *reinterpret_cast<const float4*>(p);
      R.v[2 * i + 1] = *reinterpret_cast<const float4*>(p + 4);
    } else {
      R.v[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
      R.v[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

__device__ __forceinline__ void tile_finish(const TileRegs& R, uint8_t* dst, int64_t ld, int mn_major, int mn0, int mn_ext,
                                            int k0, int k_end, const float* pro_scale, const float* pro_shift,
                                            const float* pro_mask, float mask_scale, int t) {
  const int ncc = mn_major ? 16 : 8;
  const uint32_t plane = mn_major ? kPlaneMN : kPlaneK;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = t + 128 * i;
    const int row = q / ncc, cc = q - row * ncc;
    const int g_row = mn_major ? k0 + row : mn0 + row;
    const int g_col = mn_major ? mn0 + cc * 8 : k0 + cc * 8;
    const bool ok = mn_major ? (g_row < k_end && g_col < mn_ext) : (g_row < mn_ext && g_col < k_end);
    float v[8] = {R.v[2 * i].x, R.v[2 * i].y, R.v[2 * i].z, R.v[2 * i].w,
                  R.v[2 * i + 1].x, R.v[2 * i + 1].y, R.v[2 * i + 1].z, R.v[2 * i + 1].w};
    if (ok && pro_scale) {
      const float4 s0 = *reinterpret_cast<const float4*>(pro_scale + g_col), s1 = *reinterpret_cast<const float4*>(pro_scale + g_col + 4);
      const float4 h0 = *reinterpret_cast<const float4*>(pro_shift + g_col), h1 = *reinterpret_cast<const float4*>(pro_shift + g_col + 4);
      const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaxf(fmaf(v[e], sc[e], sh[e]), 0.f);
    }
    if (ok && pro_mask) {
      const float* mp = pro_mask + (int64_t)g_row * ld + g_col;
      const float4 m0 = *reinterpret_cast<const float4*>(mp), m1 = *reinterpret_cast<const float4*>(mp + 4);
      const float mk[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= mk[e] * mask_scale;
    }
    __nv_bfloat162 b0 = __floats2bfloat162_rn(v[0], v[1]), b1 = __floats2bfloat162_rn(v[2], v[3]),
                   b2 = __floats2bfloat162_rn(v[4], v[5]), b3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 out;
    out.x = *reinterpret_cast<uint32_t*>(&b0); out.y = *reinterpret_cast<uint32_t*>(&b1);
    out.z = *reinterpret_cast<uint32_t*>(&b2); out.w = *reinterpret_cast<uint32_t*>(&b3);
    *reinterpret_cast<uint4*>(dst + cc * plane + row * 16) = out;
  }
}

static __global__ void __launch_bounds__(kThreads, 1) fc_gemm_kernel(const Params P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kTileStride;
  Bars* bars = reinterpret_cast<Bars*>(smem + 2 * kStages * kTileStride);
  const int tid = threadIdx.x, warp = uniform_warp_idx();
  const int i0 = blockIdx.x * 128, j0 = blockIdx.y * 128;
  int kchunk = (P.K + P.ksplit - 1) / P.ksplit;
  kchunk = (kchunk + 63) & ~63;
  const int kbeg = blockIdx.z * kchunk, kend = min(P.K, kbeg + kchunk);
  const int nkb = kend > kbeg ? (kend - kbeg + 63) / 64 : 0;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars->full[i], 128); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->done, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(&bars->tmem_base, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp >= 4 && warp < 8) {
    const int t = tid - 128;
    uint32_t ph_e[kStages];
#pragma unroll
    for (int i = 0; i < kStages; ++i) ph_e[i] = 1;
    TileRegs ra, rb;
    if (nkb > 0) {
      tile_issue(ra, P.A, P.lda, P.a_mn, i0, P.M, kbeg, kend, t);
      tile_issue(rb, P.B, P.ldb, P.b_mn, j0, P.N, kbeg, kend, t);
    }
    for (int kb = 0; kb < nkb; ++kb) {
      const int st = kb % kStages;
      mbar_wait(&bars->empty[st], ph_e[st]); ph_e[st] ^= 1;
      const int k0 = kbeg + kb * 64;
      tile_finish(ra, sA + st * kTileStride, P.lda, P.a_mn, i0, P.M, k0, kend, P.pro_scale, P.pro_shift, P.pro_mask,
                  P.pro_mask_scale, t);
      tile_finish(rb, sB + st * kTileStride, P.ldb, P.b_mn, j0, P.N, k0, kend, nullptr, nullptr, nullptr, 1.f, t);
      if (kb + 1 < nkb) {   // next block's loads are issued before this block is handed to the MMA thread
        tile_issue(ra, P.A, P.lda, P.a_mn, i0, P.M, k0 + 64, kend, t);
        tile_issue(rb, P.B, P.ldb, P.b_mn, j0, P.N, k0 + 64, kend, t);
      }
      fence_proxy_async_smem();
      mbar_arrive(&bars->full[st]);
    }
  } else if (warp == 8) {
    {
      uint32_t ph_f[kStages];
#pragma unroll
      for (int i = 0; i < kStages; ++i) ph_f[i] = 0;
      const uint32_t idesc = make_idesc(128, 128, P.a_mn, P.b_mn);
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % kStages;
        mbar_wait(&bars->full[st], ph_f[st]); ph_f[st] ^= 1;
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + st * kTileStride), b_base = smem_u32(sB + st * kTileStride);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = P.a_mn ? make_desc(a_base + ks * 256, 128, kPlaneMN) : make_desc(a_base + ks * 2 * kPlaneK, kPlaneK, 128);
          const uint64_t bd = P.b_mn ? make_desc(b_base + ks * 256, 128, kPlaneMN) : make_desc(b_base + ks * 2 * kPlaneK, kPlaneK, 128);
          mma_bf16(tmem, ad, bd, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
        }
        mma_commit(&bars->empty[st]);
      }
      mma_commit(&bars->done);
    }
  } else if (nkb > 0) {
    mbar_wait(&bars->done, 0);
    tc_fence_after();
    const int i = i0 + tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool reduce = P.ksplit > 1 || P.accumulate;
    for (int g16 = 0; g16 < 128; g16 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + lane_base + g16, r);
      tmem_ld_wait();
      if (i < P.M) {
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
          const int j = j0 + g16 + j4;
          if (j >= P.N) continue;
          float v[4] = {__uint_as_float(r[j4]), __uint_as_float(r[j4 + 1]), __uint_as_float(r[j4 + 2]),
                        __uint_as_float(r[j4 + 3])};
          if (P.bias && blockIdx.z == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] += P.bias[j + e];
          }
          float* dst = P.C + (int64_t)i * P.ldc + j;
          if (reduce) red_add_v4(dst, v[0], v[1], v[2], v[3]);
          else *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 128);
}

inline double min_flop() {
  const char* e = getenv("AN3D_FC_TENSOR_MIN_FLOP");
  return e ? atof(e) : 1.0e9;
}

// usable when every 16-byte access of the kernel is aligned and the extents are chunkable
inline bool usable(const Params& p) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int a_contig = p.a_mn ? p.M : p.K, b_contig = p.b_mn ? p.N : p.K;
  return al(p.A) && al(p.B) && al(p.C) && (p.lda % 4 == 0) && (p.ldb % 4 == 0) && (p.ldc % 4 == 0) && (a_contig % 8 == 0) &&
         (b_contig % 8 == 0) && (p.N % 4 == 0) && (!p.pro_scale || (al(p.pro_scale) && al(p.pro_shift))) &&
         (!p.pro_mask || al(p.pro_mask)) && p.K >= 64 && p.M >= 64 && p.N >= 64 &&
         // below ~1 GFLOP the launch is latency-bound and the SIMT fp32 GEMM (more, smaller tiles) is faster;
         // AN3D_FC_TENSOR_MIN_FLOP overrides the threshold (the test-suite sets 0 to exercise this kernel)
         2.0 * p.M * p.N * p.K >= min_flop();
}

static int launch(const Params& p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    AN3D_CUDA_CHECK(cudaFuncSetAttribute(fc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set = true;
  }
  dim3 grid((p.M + 127) / 128, (p.N + 127) / 128, p.ksplit);
  prof_mark(PROF_FC, true, st);
  fc_gemm_kernel<<<grid, kThreads, kSmemBytes, st>>>(p);
  prof_mark(PROF_FC, false, st);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

}  // namespace fcgemm
}  // namespace an3d
