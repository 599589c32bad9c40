// FC layers of the bf16 path, second generation: bf16 operand IMAGES in global memory, fetched by the TMA unit.
//
//   C[i, j] (+)= sum_k A(i,k) * B(j,k)  (+ bias[j])        tile 128 x 128, K in blocks of 128
//
// Round 1's FC GEMM converted fp32 operands on loader warps: every CTA re-read and re-converted its fp32 A and B tiles
// (268 MB through L2 for the 4096 x 2048 x 512 head layer, 66 us = 0.05 of the tensor peak) and the register-staged
// pipeline was one K block deep (10-30 us for GEMMs with a microsecond of math).  Here every matrix that enters an FC
// GEMM is packed ONCE into a bf16 "plane-major" image -- by the kernel that produces it or by `pack_kernel`, which
// also applies the BN affine + ReLU + dropout mask of the producing layer -- and each image serves every role of that
// matrix (activations: forward + wgrad; weights: forward + dgrad; gradients: wgrad + dgrad):
//
//   image of X[R rows][C cols]:  blocks of 128 rows x 128 columns, block (rb, cb) at (rb * ncb + cb) * 32 KB; inside a block
//                                16 planes (8 columns each) of 128 rows x 16 bytes:  ((c8 % 16) * 128 + r % 128) * 16.
//                                Zero padded to whole blocks.
//
// A block IS the un-swizzled shared-memory operand tile of umma.cuh -- K-major (rows = M / N index, planes = K chunks) and
// MN-major (rows = K index, planes = M / N chunks) alike -- so every operand tile of a 128-wide K block is ONE 32 KB bulk
// copy (cp.async.bulk, TMA unit) issued by one thread through a 3-stage mbarrier ring.  No loader warps, no
// conversions, no bounds logic.  (A first version with 1-2 KB pieces per plane spent 0.7 us per K block just issuing
// its 24 copies: profiles/r2_fc_launches.txt.)
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace an3d {
namespace fc2 {

using namespace umma;

struct Image {
  const __nv_bfloat16* g = nullptr;
  int rows = 0, cols = 0;
  __host__ __device__ int row_blocks() const { return (rows + 127) >> 7; }
  __host__ __device__ int col_blocks() const { return (cols + 127) >> 7; }
  __host__ __device__ const __nv_bfloat16* block(int rb, int cb) const { return g + ((int64_t)rb * col_blocks() + cb) * 16384; }
};

// ---------------------------------------------------------------------------------------------
// packing: fp32 matrix (+ optional BN affine + ReLU, + optional dropout mask) -> bf16 image.  One thread per
// (row, chunk); rows >= R and columns >= C are written as zeros, so an image never needs a memset.
// ---------------------------------------------------------------------------------------------
struct PackArgs {
  const float* src = nullptr; int64_t ld = 0;
  int rows = 0, cols = 0;
  const float* scale = nullptr;      // [cols] or nullptr:  relu(x * scale + shift)
  const float* shift = nullptr;
  const float* mask = nullptr;       // same layout as src, or nullptr
  float mask_scale = 1.f;
  __nv_bfloat16* dst = nullptr;
};

static __global__ void __launch_bounds__(256) pack_kernel(const PackArgs a0, const PackArgs a1) {
  const PackArgs a = blockIdx.z ? a1 : a0;
  const int rows_pad = (a.rows + 127) & ~127, chunks = ((a.cols + 127) & ~127) >> 3;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows_pad * chunks) return;
  // consecutive threads walk the rows of one chunk: every lane reads one full 32-byte sector of its row, and the warp's
  // 16-byte writes are contiguous (the other order wrote 16 bytes per 2 KB: 15 us for a 4096 x 512 pair)
  const int c8 = (int)(i / rows_pad), r = (int)(i - (int64_t)c8 * rows_pad);
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  if (r < a.rows && c8 * 8 < a.cols) {
    const float* p = a.src + (int64_t)r * a.ld + c8 * 8;
    const int n = min(8, a.cols - c8 * 8);
    const bool vec = n == 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
    if (vec) {
      const float4 lo = *reinterpret_cast<const float4*>(p), hi = *reinterpret_cast<const float4*>(p + 4);
      v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (e < n) v[e] = p[e];
    }
    if (a.scale) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (e < n) v[e] = fmaxf(fmaf(v[e], a.scale[c8 * 8 + e], a.shift[c8 * 8 + e]), 0.f);
    }
    if (a.mask) {
      const float* mp = a.mask + (int64_t)r * a.ld + c8 * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (e < n) v[e] *= mp[e] * a.mask_scale;
    }
  }
  __nv_bfloat162 b0 = __floats2bfloat162_rn(v[0], v[1]), b1 = __floats2bfloat162_rn(v[2], v[3]),
                 b2 = __floats2bfloat162_rn(v[4], v[5]), b3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 out;
  out.x = *reinterpret_cast<uint32_t*>(&b0); out.y = *reinterpret_cast<uint32_t*>(&b1);
  out.z = *reinterpret_cast<uint32_t*>(&b2); out.w = *reinterpret_cast<uint32_t*>(&b3);
  const int ncb = chunks >> 4;
  const int64_t blk = (int64_t)(r >> 7) * ncb + (c8 >> 4);
  *reinterpret_cast<uint4*>(a.dst + blk * 16384 + ((c8 & 15) * 128 + (r & 127)) * 8) = out;
}

// one launch packs one matrix, or two of identical shape (the two siamese branches)
static int pack(const PackArgs& a, cudaStream_t st, const PackArgs* b = nullptr) {
  if (b && (b->rows != a.rows || b->cols != a.cols)) {
    set_error("fc2::pack: batched matrices must have identical shapes");
    return AN3D_ERR_INVALID;
  }
  const int64_t total = (int64_t)((a.rows + 127) & ~127) * (((a.cols + 127) & ~127) >> 3);
  dim3 grid((unsigned)((total + 255) / 256), 1, b ? 2 : 1);
  pack_kernel<<<grid, 256, 0, st>>>(a, b ? *b : a);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

// ---------------------------------------------------------------------------------------------
// the GEMM
// ---------------------------------------------------------------------------------------------
struct Params {
  Image A; int a_mn = 0;      // a_mn = 0: image rows = i (M index), image cols = k      1: image rows = k, image cols = i
  Image B; int b_mn = 0;      // b_mn = 0: image rows = j (N index), image cols = k      1: image rows = k, image cols = j
  float* C = nullptr; int64_t ldc = 0;
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;              // [N] or nullptr
  int ksplit = 1;                           // gridDim.z; > 1 -> accumulate with reductions into pre-zeroed C
  int accumulate = 0;                       // reductions even with ksplit == 1
  double* stat_sum = nullptr;               // optional [N]: += column sums of C (bias included); needs ksplit == 1
  double* stat_sq = nullptr;                // optional [N]: += column sums of C^2
  int c_vec = 1;                            // set by launch(): 16-byte vector access legal for C
};

constexpr int kThreads = 192;                 // warps 0-3 epilogue, 4 TMA issuer, 5 MMA
constexpr int kStages = 3;
constexpr int kKBlock = 128;
constexpr uint32_t kPlane = 128 * 16;         // plane stride inside a block (both majors)
constexpr uint32_t kTileBytes = 16 * kPlane;  // 32 KB
constexpr uint32_t kTxBytes = 2 * kTileBytes; // both operand tiles of a K block
constexpr size_t kSmemBytes = 2 * kStages * (size_t)kTileBytes + 1024;
static_assert(kStages * kTileBytes >= 128 * 129 * 4, "the statistics transpose tile reuses the A ring");

struct Bars {
  uint64_t full[kStages], empty[kStages], done;
  uint32_t tmem_base;
  float bias[128];      // this tile's bias values, staged while the MMAs run
};

// Two problems of identical shape (the two siamese branches of one FC layer: same weights, their own activations,
// statistics and output) can share a launch: blockIdx.z = batch * ksplit + k-slice.
static __global__ void __launch_bounds__(kThreads, 1) fc2_gemm_kernel(const Params P0, const Params P1) {
  const int bz = (int)blockIdx.z / P0.ksplit, kz = (int)blockIdx.z - bz * P0.ksplit;
  const Params& P = bz ? P1 : P0;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kTileBytes;
  Bars* bars = reinterpret_cast<Bars*>(smem + 2 * kStages * kTileBytes);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int i0 = blockIdx.x * 128, j0 = blockIdx.y * 128;
  int kchunk = (P.K + P.ksplit - 1) / P.ksplit;
  kchunk = (kchunk + kKBlock - 1) & ~(kKBlock - 1);
  const int kbeg = kz * kchunk, kend = min(P.K, kbeg + kchunk);
  const int nkb = kend > kbeg ? (kend - kbeg + kKBlock - 1) / kKBlock : 0;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->done, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(&bars->tmem_base, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 4) {
    if (lane == 0) {
      uint32_t ph = (1u << kStages) - 1u;           // first pass over the ring falls through (parity trick)
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % kStages;
        mbar_wait(&bars->empty[st], (ph >> st) & 1u); ph ^= 1u << st;
        mbar_arrive_expect_tx(&bars->full[st], kTxBytes);
        const int kblk = (kbeg >> 7) + kb, ib = i0 >> 7, jb = j0 >> 7;
        // K-major: block (row block = M / N tile, column block = K block); MN-major: block (K block, M / N tile)
        bulk_copy_g2s(sA + st * kTileBytes, P.a_mn ? P.A.block(kblk, ib) : P.A.block(ib, kblk), kTileBytes, &bars->full[st]);
        bulk_copy_g2s(sB + st * kTileBytes, P.b_mn ? P.B.block(kblk, jb) : P.B.block(jb, kblk), kTileBytes, &bars->full[st]);
      }
    }
  } else if (warp == 5) {
    uint32_t ph = 0;
    const uint32_t idesc = make_idesc(128, 128, P.a_mn, P.b_mn);
    for (int kb = 0; kb < nkb; ++kb) {
      const int st = kb % kStages;
      mbar_wait(&bars->full[st], (ph >> st) & 1u); ph ^= 1u << st;
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA + st * kTileBytes), b_base = smem_u32(sB + st * kTileBytes);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = P.a_mn ? make_desc(a_base + ks * 256, 128, kPlane) : make_desc(a_base + ks * 2 * kPlane, kPlane, 128);
          const uint64_t bd = P.b_mn ? make_desc(b_base + ks * 256, 128, kPlane) : make_desc(b_base + ks * 2 * kPlane, kPlane, 128);
          mma_bf16_raw(tmem, ad, bd, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
        }
        mma_commit_raw(&bars->empty[st]);
        if (kb == nkb - 1) mma_commit_raw(&bars->done);
      }
      __syncwarp();
    }
  } else if (nkb > 0) {
    // (the bias used to be read from global memory inside the store loop: 16 % of this kernel's stall samples)
    bars->bias[tid] = (P.bias && kz == 0 && j0 + tid < P.N) ? P.bias[j0 + tid] : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    mbar_wait_relaxed(&bars->done, 0);
    tc_fence_after();
    const int i = i0 + tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool reduce = P.ksplit > 1 || P.accumulate;
    const bool stats = P.stat_sum != nullptr;
    float* sT = reinterpret_cast<float*>(sA);      // [128 cols][129]: the operand ring is idle once `done` fired
    for (int g32 = 0; g32 < 128; g32 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem + lane_base + g32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j4 = 0; j4 < 32; j4 += 4) {
        const int j = j0 + g32 + j4;
        float v[4] = {__uint_as_float(r[j4]), __uint_as_float(r[j4 + 1]), __uint_as_float(r[j4 + 2]),
                      __uint_as_float(r[j4 + 3])};
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] += bars->bias[g32 + j4 + e];
        if (stats) {
#pragma unroll
          for (int e = 0; e < 4; ++e) sT[(g32 + j4 + e) * 129 + tid] = (i < P.M && j + e < P.N) ? v[e] : 0.f;
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
  if (warp == 5) tmem_dealloc(tmem, 128);
}

// p1 == nullptr: one problem; otherwise two problems of identical shape / majors / ksplit in one launch
static int launch(Params p, cudaStream_t st, const Params* p1 = nullptr) {
  static bool attr_set = false;
  if (!attr_set) {
    AN3D_CUDA_CHECK(cudaFuncSetAttribute(fc2_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set = true;
  }
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  auto check = [&](const Params& q) {
    const int a_rows = q.a_mn ? q.K : q.M, a_cols = q.a_mn ? q.M : q.K, b_rows = q.b_mn ? q.K : q.N, b_cols = q.b_mn ? q.N : q.K;
    return q.M > 0 && q.N > 0 && q.K > 0 && q.A.g && q.B.g && q.C && q.A.rows == a_rows && q.A.cols == a_cols &&
           q.B.rows == b_rows && q.B.cols == b_cols && !(q.stat_sum && q.ksplit > 1);
  };
  Params q = p1 ? *p1 : p;
  if (!check(p) || !check(q) || q.M != p.M || q.N != p.N || q.K != p.K || q.a_mn != p.a_mn || q.b_mn != p.b_mn ||
      q.ksplit != p.ksplit) {
    set_error("fc2::launch: bad operands (shape %d x %d x %d)", p.M, p.N, p.K);
    return AN3D_ERR_INVALID;
  }
  p.c_vec = al(p.C) && p.ldc % 4 == 0 && p.N % 4 == 0;
  q.c_vec = al(q.C) && q.ldc % 4 == 0 && q.N % 4 == 0;
  dim3 grid((p.M + 127) / 128, (p.N + 127) / 128, p.ksplit * (p1 ? 2 : 1));
  prof_mark(PROF_FC, true, st);
  fc2_gemm_kernel<<<grid, kThreads, kSmemBytes, st>>>(p, q);
  prof_mark(PROF_FC, false, st);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

}  // namespace fc2
}  // namespace an3d
