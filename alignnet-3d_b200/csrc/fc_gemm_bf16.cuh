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
  const float* A = nullptr; int64_t lda = 0; int a_mn = 0;    // a_mn = 0: A[i*lda + k]   1: A[k*lda + i]
  const float* B = nullptr; int64_t ldb = 0; int b_mn = 0;    // b_mn = 0: B[j*ldb + k]   1: B[k*ldb + j]
  float* C = nullptr; int64_t ldc = 0;
  int M = 0, N = 0, K = 0;                  // extents of i, j, k
  const float* bias = nullptr;              // [N] or nullptr
  const float* pro_scale = nullptr;         // optional prologue on A: relu(a*scale[ch] + shift[ch]) (* mask * mask_scale)
  const float* pro_shift = nullptr;
  const float* pro_mask = nullptr;          // same layout as A
  float pro_mask_scale = 1.f;
  int ksplit = 1;                           // gridDim.z; > 1 -> accumulate with reductions into pre-zeroed C
  int accumulate = 0;                       // reductions even with ksplit == 1
  double* stat_sum = nullptr;               // optional [N]: += column sums of C (bias included); needs ksplit == 1
  double* stat_sq = nullptr;                // optional [N]: += column sums of C^2
  int a_vec = 1, b_vec = 1, c_vec = 1;      // set by launch(): 16-byte vector access legal for A / B / C
  int nbatch = 1;                           // set by launch(): 2 = two problems of identical shape in one launch (grid.z)
};

constexpr int kLoaderThreads = 256;
constexpr int kThreads = 128 + kLoaderThreads + 32;   // 4 epilogue warps, 8 loader warps, 1 MMA warp
constexpr int kMmaWarp = (128 + kLoaderThreads) / 32;
constexpr int kStages = 4;
constexpr int kChunksPerThread = 1024 / kLoaderThreads;   // 16-byte chunks of one 128x64 (or 64x128) tile per thread
constexpr uint32_t kPlaneK = 128 * 16 + 16;   // K-major tile: 8 planes x 128 rows
constexpr uint32_t kPlaneMN = 64 * 16 + 16;   // MN-major tile: 16 planes x 64 rows
constexpr uint32_t kTileBytes = 16 * kPlaneMN > 8 * kPlaneK ? 16 * kPlaneMN : 8 * kPlaneK;
constexpr uint32_t kTileStride = (kTileBytes + 127) & ~127u;
constexpr size_t kSmemBytes = 2 * kStages * (size_t)kTileStride + 256;
static_assert(kStages * kTileStride >= 128 * 129 * 4, "the statistics transpose tile reuses the A ring");

struct Bars {
  uint64_t full[kStages], empty[kStages], done;
  uint32_t tmem_base;
};

// Staging of one 64-K-block of an operand, split in two phases so that the global loads of BOTH
// operands are in flight together (one latency period per K block instead of two).
struct TileRegs { float4 v[2 * kChunksPerThread]; };

__device__ __forceinline__ void tile_issue(TileRegs& R, const float* src, int64_t ld, int mn_major, int vec, int mn0,
                                           int mn_ext, int k0, int k_end, int t) {
  const int ncc = mn_major ? 16 : 8;
#pragma unroll
  for (int i = 0; i < kChunksPerThread; ++i) {
    const int q = t + kLoaderThreads * i;
    const int row = q / ncc, cc = q - row * ncc;
    const int g_row = mn_major ? k0 + row : mn0 + row;
    const int g_col = mn_major ? mn0 + cc * 8 : k0 + cc * 8;
    const int row_ext = mn_major ? k_end : mn_ext, col_ext = mn_major ? mn_ext : k_end;
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    if (g_row < row_ext && g_col < col_ext) {
      const float* p = src + (int64_t)g_row * ld + g_col;
      if (vec) {
        lo = *reinterpret_cast<const float4*>(p);
        hi = *reinterpret_cast<const float4*>(p + 4);
      } else {   // unaligned rows / ragged extent (e.g. the 103-wide output layers): element-wise, bounds-checked
        const int n = col_ext - g_col;
        lo.x = p[0];
        if (n > 1) lo.y = p[1];
        if (n > 2) lo.z = p[2];
        if (n > 3) lo.w = p[3];
        if (n > 4) hi.x = p[4];
        if (n > 5) hi.y = p[5];
        if (n > 6) hi.z = p[6];
        if (n > 7) hi.w = p[7];
      }
    }
    R.v[2 * i] = lo;
    R.v[2 * i + 1] = hi;
  }
}

__device__ __forceinline__ void tile_finish(const TileRegs& R, uint8_t* dst, int64_t ld, int mn_major, int vec, int mn0,
                                            int mn_ext, int k0, int k_end, const float* pro_scale, const float* pro_shift,
                                            const float* pro_mask, float mask_scale, int t) {
  const int ncc = mn_major ? 16 : 8;
  const uint32_t plane = mn_major ? kPlaneMN : kPlaneK;
#pragma unroll
  for (int i = 0; i < kChunksPerThread; ++i) {
    const int q = t + kLoaderThreads * i;
    const int row = q / ncc, cc = q - row * ncc;
    const int g_row = mn_major ? k0 + row : mn0 + row;
    const int g_col = mn_major ? mn0 + cc * 8 : k0 + cc * 8;
    const int row_ext = mn_major ? k_end : mn_ext, col_ext = mn_major ? mn_ext : k_end;
    const bool ok = g_row < row_ext && g_col < col_ext;
    const int n = col_ext - g_col;
    float v[8] = {R.v[2 * i].x, R.v[2 * i].y, R.v[2 * i].z, R.v[2 * i].w,
                  R.v[2 * i + 1].x, R.v[2 * i + 1].y, R.v[2 * i + 1].z, R.v[2 * i + 1].w};
    if (ok && pro_scale) {
      if (vec) {
        const float4 s0 = *reinterpret_cast<const float4*>(pro_scale + g_col), s1 = *reinterpret_cast<const float4*>(pro_scale + g_col + 4);
        const float4 h0 = *reinterpret_cast<const float4*>(pro_shift + g_col), h1 = *reinterpret_cast<const float4*>(pro_shift + g_col + 4);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaxf(fmaf(v[e], sc[e], sh[e]), 0.f);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = e < n ? fmaxf(fmaf(v[e], pro_scale[g_col + e], pro_shift[g_col + e]), 0.f) : 0.f;
      }
    }
    if (ok && pro_mask) {
      const float* mp = pro_mask + (int64_t)g_row * ld + g_col;
      if (vec) {
        const float4 m0 = *reinterpret_cast<const float4*>(mp), m1 = *reinterpret_cast<const float4*>(mp + 4);
        const float mk[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= mk[e] * mask_scale;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = e < n ? v[e] * mp[e] * mask_scale : 0.f;
      }
    }
    __nv_bfloat162 b0 = __floats2bfloat162_rn(v[0], v[1]), b1 = __floats2bfloat162_rn(v[2], v[3]),
                   b2 = __floats2bfloat162_rn(v[4], v[5]), b3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 out;
    out.x = *reinterpret_cast<uint32_t*>(&b0); out.y = *reinterpret_cast<uint32_t*>(&b1);
    out.z = *reinterpret_cast<uint32_t*>(&b2); out.w = *reinterpret_cast<uint32_t*>(&b3);
    *reinterpret_cast<uint4*>(dst + cc * plane + row * 16) = out;
  }
}

// Two problems of identical shape (the two siamese branches of one FC layer: same weights, their own activations,
// BN prologue, statistics and output) can share a launch: blockIdx.z = batch * ksplit + k-slice.
static __global__ void __launch_bounds__(kThreads, 1) fc_gemm_kernel(const Params P0, const Params P1) {
  const int bz = (int)blockIdx.z / P0.ksplit, kz = (int)blockIdx.z - bz * P0.ksplit;
  const Params P = bz ? P1 : P0;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kTileStride;
  Bars* bars = reinterpret_cast<Bars*>(smem + 2 * kStages * kTileStride);
  const int tid = threadIdx.x, warp = uniform_warp_idx();
  const int i0 = blockIdx.x * 128, j0 = blockIdx.y * 128;
  int kchunk = (P.K + P.ksplit - 1) / P.ksplit;
  kchunk = (kchunk + 63) & ~63;
  const int kbeg = kz * kchunk, kend = min(P.K, kbeg + kchunk);
  const int nkb = kend > kbeg ? (kend - kbeg + 63) / 64 : 0;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars->full[i], kLoaderThreads); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->done, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&bars->tmem_base, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp >= 4 && warp < kMmaWarp) {
    const int t = tid - 128;
    uint32_t ph_e[kStages];
#pragma unroll
    for (int i = 0; i < kStages; ++i) ph_e[i] = 1;
    TileRegs ra, rb;
    if (nkb > 0) {
      tile_issue(ra, P.A, P.lda, P.a_mn, P.a_vec, i0, P.M, kbeg, kend, t);
      tile_issue(rb, P.B, P.ldb, P.b_mn, P.b_vec, j0, P.N, kbeg, kend, t);
    }
    for (int kb = 0; kb < nkb; ++kb) {
      const int st = kb % kStages;
      mbar_wait(&bars->empty[st], ph_e[st]); ph_e[st] ^= 1;
      const int k0 = kbeg + kb * 64;
      tile_finish(ra, sA + st * kTileStride, P.lda, P.a_mn, P.a_vec, i0, P.M, k0, kend, P.pro_scale, P.pro_shift, P.pro_mask,
                  P.pro_mask_scale, t);
      tile_finish(rb, sB + st * kTileStride, P.ldb, P.b_mn, P.b_vec, j0, P.N, k0, kend, nullptr, nullptr, nullptr, 1.f, t);
      if (kb + 1 < nkb) {   // next block's loads are issued before this block is handed to the MMA thread
        tile_issue(ra, P.A, P.lda, P.a_mn, P.a_vec, i0, P.M, k0 + 64, kend, t);
        tile_issue(rb, P.B, P.ldb, P.b_mn, P.b_vec, j0, P.N, k0 + 64, kend, t);
      }
      fence_proxy_async_smem();
      mbar_arrive(&bars->full[st]);
    }
  } else if (warp == kMmaWarp) {
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
  } else if (nkb > 0) {
    mbar_wait_relaxed(&bars->done, 0);
    tc_fence_after();
    const int i = i0 + tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool reduce = P.ksplit > 1 || P.accumulate;
    const bool stats = P.stat_sum != nullptr;
    float* sT = reinterpret_cast<float*>(sA);      // [128 cols][129]: the operand ring is idle once `done` fired
    for (int g16 = 0; g16 < 128; g16 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + lane_base + g16, r);
      tmem_ld_wait();
#pragma unroll
      for (int j4 = 0; j4 < 16; j4 += 4) {
        const int j = j0 + g16 + j4;
        float v[4] = {__uint_as_float(r[j4]), __uint_as_float(r[j4 + 1]), __uint_as_float(r[j4 + 2]),
                      __uint_as_float(r[j4 + 3])};
        if (P.bias && kz == 0) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (j + e < P.N) v[e] += P.bias[j + e];
        }
        if (stats) {
#pragma unroll
          for (int e = 0; e < 4; ++e) sT[(g16 + j4 + e) * 129 + tid] = (i < P.M && j + e < P.N) ? v[e] : 0.f;
        }
        if (i < P.M && j < P.N) {
          float* dst = P.C + (int64_t)i * P.ldc + j;
          if (P.c_vec) {
            if (reduce) red_add_v4(dst, v[0], v[1], v[2], v[3]);
            else *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (j + e < P.N) {
                if (reduce) atomicAdd(dst + e, v[e]);
                else dst[e] = v[e];
              }
            }
          }
        }
      }
    }
    tc_fence_before();
    if (stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int j = j0 + tid;
      if (j < P.N) {
        const float* col = sT + tid * 129;
        float s = 0.f, ss = 0.f;
#pragma unroll 8
        for (int r2 = 0; r2 < 128; ++r2) { const float z = col[r2]; s += z; ss = fmaf(z, z, ss); }
        atomicAdd(P.stat_sum + j, (double)s);
        atomicAdd(P.stat_sq + j, (double)ss);
      }
    }
  }
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, 128);
}

inline double min_flop() {
  const char* e = getenv("AN3D_FC_TENSOR_MIN_FLOP");
  return e ? atof(e) : 0.0;
}

// every shape is supported (ragged / unaligned operands take the element-wise loader); AN3D_FC_TENSOR_MIN_FLOP
// lets a test send small products to the SIMT fp32 GEMM instead
inline bool usable(const Params& p) {
  return p.M > 0 && p.N > 0 && p.K > 0 && 2.0 * p.M * p.N * p.K >= min_flop() && !(p.stat_sum && p.ksplit > 1);
}

inline void finish_params(Params& p) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int a_contig = p.a_mn ? p.M : p.K, b_contig = p.b_mn ? p.N : p.K;
  p.a_vec = al(p.A) && p.lda % 4 == 0 && a_contig % 8 == 0 && (!p.pro_scale || (al(p.pro_scale) && al(p.pro_shift))) &&
            (!p.pro_mask || al(p.pro_mask));
  p.b_vec = al(p.B) && p.ldb % 4 == 0 && b_contig % 8 == 0;
  p.c_vec = al(p.C) && p.ldc % 4 == 0 && p.N % 4 == 0;
}

// p1 == nullptr: one problem; otherwise two problems of identical shape / majors / ksplit in one launch
static int launch(Params p, cudaStream_t st, const Params* p1 = nullptr) {
  static bool attr_set = false;
  if (!attr_set) {
    AN3D_CUDA_CHECK(cudaFuncSetAttribute(fc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set = true;
  }
  finish_params(p);
  Params q = p;
  if (p1) {
    q = *p1;
    if (q.M != p.M || q.N != p.N || q.K != p.K || q.a_mn != p.a_mn || q.b_mn != p.b_mn || q.ksplit != p.ksplit) {
      set_error("fcgemm::launch: batched problems must have identical shapes");
      return AN3D_ERR_INVALID;
    }
    finish_params(q);
  }
  p.nbatch = q.nbatch = p1 ? 2 : 1;
  dim3 grid((p.M + 127) / 128, (p.N + 127) / 128, p.ksplit * p.nbatch);
  prof_mark(PROF_FC, true, st);
  fc_gemm_kernel<<<grid, kThreads, kSmemBytes, st>>>(p, q);
  prof_mark(PROF_FC, false, st);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

}  // namespace fcgemm
}  // namespace an3d
