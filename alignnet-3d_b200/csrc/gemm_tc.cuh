// Split-operand tensor-core GEMM of the materialised path (AN3D_PRECISION_BF16X3 / _BF16X6, and AN3D_PRECISION_BF16
// for conv stacks the fused kernels do not cover): fp32 matrices enter as a SUM of bf16 images
//
//     X = X0 + X1 (+ X2),   X0 = bf16(X), X1 = bf16(X - X0), X2 = bf16(X - X0 - X1)
//
// and C = A * B is the sum of the bf16 x bf16 -> fp32 tcgen05 products of the image pairs whose weight is above the
// target precision, all accumulated in ONE TMEM tile:
//     1 image  (bf16):    A0B0                                     (8 significant bits)
//     2 images (bf16x3):  A0B0 + A0B1 + A1B0                       (dropped A1B1 ~ 2^-18 |a||b|)
//     3 images (bf16x6):  A0B0 + A0B1 + A1B0 + A1B1 + A0B2 + A2B0  (~ 2^-24: fp32 grade)
// Every image is the plane-major block image of fc2_gemm.cuh, so an operand tile is one 32 KB bulk copy and the same
// image serves as K-major or MN-major operand (forward, wgrad, dgrad of a layer all read the images packed once).
// The kernel is fc2_gemm_kernel's structure (one TMA-issuing thread, one MMA-issuing warp, four epilogue warps) with the
// K loop running over (K block, term) pairs, plus a single-stage variant (three CTAs per SM) for the conv layers'
// forward GEMMs, whose K is one block.  The accumulation length of one TMEM tile is capped (kMaxKBlocksPerCta): the
// tensor core's fp32 accumulate is not round-to-nearest, and millions of rows per reduction (wgrad over all points)
// would let that bias grow; the K slices meet in fp32 reductions.
#pragma once
#include <algorithm>

#include "common.cuh"
#include "fc2_gemm.cuh"
#include "kernels_f32.cuh"
#include "umma.cuh"

namespace an3d {
namespace tcg {

using namespace umma;
using fc2::Image;

constexpr int kMaxSplit = 3;
constexpr int kMaxTerms = 6;
constexpr int kMaxKBlocksPerCta = 32;

// one fp32 matrix as 1-3 bf16 images of identical geometry
struct SplitMat {
  Image img[kMaxSplit];
  int n = 0;
};

__host__ __device__ inline int num_terms(int nsplit) { return nsplit == 1 ? 1 : (nsplit == 2 ? 3 : 6); }

// ---------------------------------------------------------------------------------------------
// packing: fp32 matrix (+ BN affine + ReLU, + dropout mask of the producing layer) -> nsplit bf16 images
// ---------------------------------------------------------------------------------------------
struct PackArgs {
  const float* src = nullptr; int64_t ld = 0;
  int rows = 0, cols = 0;
  const float* scale = nullptr;      // [cols] or nullptr:  relu(x * scale + shift)
  const float* shift = nullptr;
  const float* mask = nullptr;       // same layout as src, or nullptr
  float mask_scale = 1.f;
  __nv_bfloat16* dst[kMaxSplit] = {nullptr, nullptr, nullptr};
  int nsplit = 1;
};

static __global__ void __launch_bounds__(256) pack_split_kernel(const PackArgs a) {
  const int rows_pad = (a.rows + 127) & ~127, chunks = ((a.cols + 127) & ~127) >> 3;
  // a warp covers 8 rows x 4 chunks of 8 columns: 128 contiguous bytes of each of its rows in, 128 contiguous bytes of
  // each of its 4 image planes out (8 rows x 16 B).  (One chunk per lane along the rows -- fc2::pack_kernel's mapping,
  // fine for FC-sized matrices -- moved 2.7 TB/s on the [819200, 1024] gradient: 32-byte pieces at a 4 KB stride.)
  const int cgroups = chunks >> 2;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t rg = w / cgroups;
  if (rg * 8 >= rows_pad) return;
  const int r = (int)(rg * 8) + (lane & 7), c8 = (int)(w - rg * cgroups) * 4 + (lane >> 3);
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
  const int ncb = chunks >> 4;
  const int64_t off = ((int64_t)(r >> 7) * ncb + (c8 >> 4)) * 16384 + ((c8 & 15) * 128 + (r & 127)) * 8;
#pragma unroll
  for (int s = 0; s < kMaxSplit; ++s) {
    if (s < a.nsplit) {
      __nv_bfloat162 b[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        b[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        // the residual is exact in fp32 (v and its bf16 rounding share the exponent range)
        v[2 * e] -= __bfloat162float(b[e].x);
        v[2 * e + 1] -= __bfloat162float(b[e].y);
      }
      uint4 out;
      out.x = *reinterpret_cast<uint32_t*>(&b[0]); out.y = *reinterpret_cast<uint32_t*>(&b[1]);
      out.z = *reinterpret_cast<uint32_t*>(&b[2]); out.w = *reinterpret_cast<uint32_t*>(&b[3]);
      *reinterpret_cast<uint4*>(a.dst[s] + off) = out;
    }
  }
}

static int pack(const PackArgs& a, cudaStream_t st, SplitMat* out) {
  if (a.nsplit < 1 || a.nsplit > kMaxSplit || a.rows < 1 || a.cols < 1) {
    set_error("tcg::pack: bad arguments (%d x %d, %d images)", a.rows, a.cols, a.nsplit);
    return AN3D_ERR_INVALID;
  }
  const int64_t total = (int64_t)((a.rows + 127) & ~127) * (((a.cols + 127) & ~127) >> 3);   // one thread per (row, chunk)
  pack_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
  AN3D_LAUNCH_CHECK();
  out->n = a.nsplit;
  for (int s = 0; s < a.nsplit; ++s) { out->img[s].g = a.dst[s]; out->img[s].rows = a.rows; out->img[s].cols = a.cols; }
  return AN3D_OK;
}

// ---------------------------------------------------------------------------------------------
// the GEMM:  C[i, j] (+)= sum_terms sum_k A_t(i,k) * B_t(j,k)  (+ bias[j])
// ---------------------------------------------------------------------------------------------
struct Params {
  SplitMat A; int a_mn = 0;   // a_mn = 0: image rows = i, image cols = k      1: image rows = k, image cols = i
  SplitMat B; int b_mn = 0;   // b_mn = 0: image rows = j, image cols = k      1: image rows = k, image cols = j
  float* C = nullptr; int64_t ldc = 0;
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;
  int accumulate = 0;         // reductions into C even with one K slice
  double* stat_sum = nullptr; // optional [N] (pre-zeroed): += column sums of C (bias included) ...
  double* stat_sq = nullptr;  // ... and of C^2.  Honoured by the A-stationary kernel only: launch() reports it in *stats_done
  int ksplit = 1;             // set by launch()
  int c_vec = 1;              // set by launch()
  int nterms = 1;             // set by launch()
};

constexpr int kThreads = 192;                 // warps 0-3 epilogue, 4 TMA issuer, 5 MMA
constexpr uint32_t kPlane = 128 * 16;
constexpr uint32_t kTileBytes = 16 * kPlane;  // 32 KB
constexpr uint32_t kTxBytes = 2 * kTileBytes;
template <int STAGES> constexpr size_t smem_bytes() { return 2 * STAGES * (size_t)kTileBytes + 1024; }

template <int STAGES>
struct Bars {
  uint64_t full[STAGES], empty[STAGES], done;
  uint32_t tmem_base;
  float bias[128];
};

// term t multiplies image ia of A with image ib of B; smallest contributions first.  Six-term order:
// (2,0) (0,2) (1,1) (1,0) (0,1) (0,0); three terms = the last three, one term = the last.
__device__ __forceinline__ void term_images(int nterms, int t, int* ia, int* ib) {
  const int q = 6 - nterms + t;
  *ia = (0x001102u >> (4 * q)) & 15;
  *ib = (0x010120u >> (4 * q)) & 15;
}

// Epilogue store of one warp's 32 rows x 32 columns: the TMEM load leaves a ROW per lane, so a direct store writes 32
// rows x 16 B per instruction (32 half-sector transactions at the row stride; the LSU, not HBM, paced the forward conv
// GEMMs: 9 k cycles per 128 x 128 tile).  Through a per-warp staging tile (32 x 36 floats) every store instruction
// writes 4 rows x 128 contiguous bytes instead.  `stage` is this warp's 4.5 KB; C must allow 16-byte accesses.
constexpr int kStageLd = 36;
constexpr int kStageBytesPerWarp = 32 * kStageLd * 4;
__device__ __forceinline__ void store_group_coalesced(float* stage, const float (&v)[32], float* C, int64_t ldc, int row0,
                                                      int col0, int M, int N, int lane) {
  __syncwarp();                                              // the previous group's reads of the staging tile are done
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(stage + lane * kStageLd + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  __syncwarp();
  const int rr = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rr;
    const float4 x = *reinterpret_cast<const float4*>(stage + r * kStageLd + c4);
    if (row0 + r < M && col0 + c4 < N) *reinterpret_cast<float4*>(C + (int64_t)(row0 + r) * ldc + col0 + c4) = x;
  }
}

// The same store, plus the column sums of the 32 x 32 block for the BN statistics of the layer: after the transposed
// read a lane holds 8 rows of 4 columns; two shuffle rounds fold the 4 lanes that share the columns.  On return lanes
// 0-7 hold sum / sum of squares of columns col0 + 4 * lane .. + 3 over this warp's (valid) rows, in s / q.
__device__ __forceinline__ void store_group_coalesced_stats(float* stage, const float (&v)[32], float* C, int64_t ldc, int row0,
                                                            int col0, int M, int N, int lane, float (&s)[4], float (&q)[4]) {
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8; ++k)
    *reinterpret_cast<float4*>(stage + lane * kStageLd + 4 * k) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  __syncwarp();
  const int rr = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
  for (int e = 0; e < 4; ++e) { s[e] = 0.f; q[e] = 0.f; }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rr;
    const float4 x = *reinterpret_cast<const float4*>(stage + r * kStageLd + c4);
    if (row0 + r < M && col0 + c4 < N) {
      *reinterpret_cast<float4*>(C + (int64_t)(row0 + r) * ldc + col0 + c4) = x;
      s[0] += x.x; s[1] += x.y; s[2] += x.z; s[3] += x.w;
      q[0] = fmaf(x.x, x.x, q[0]); q[1] = fmaf(x.y, x.y, q[1]); q[2] = fmaf(x.z, x.z, q[2]); q[3] = fmaf(x.w, x.w, q[3]);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    s[e] += __shfl_xor_sync(0xffffffffu, s[e], 8);  q[e] += __shfl_xor_sync(0xffffffffu, q[e], 8);
    s[e] += __shfl_xor_sync(0xffffffffu, s[e], 16); q[e] += __shfl_xor_sync(0xffffffffu, q[e], 16);
  }
}

template <int STAGES>
static __global__ void __launch_bounds__(kThreads) tc_gemm_kernel(const Params P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kTileBytes;
  Bars<STAGES>* bars = reinterpret_cast<Bars<STAGES>*>(smem + 2 * STAGES * kTileBytes);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int i0 = blockIdx.x * 128, j0 = blockIdx.y * 128, kz = blockIdx.z;
  const int nkb_total = (P.K + 127) >> 7;
  const int per = (nkb_total + P.ksplit - 1) / P.ksplit;
  const int kb0 = kz * per, nkb = max(0, min(nkb_total, kb0 + per) - kb0);
  const int nit = nkb * P.nterms;

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
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
      uint32_t ph = (1u << STAGES) - 1u;            // first pass over the ring falls through
      int st = 0;
      for (int it = 0; it < nit; ++it) {
        const int kb = it / P.nterms, t = it - kb * P.nterms;
        int ia, ib;
        term_images(P.nterms, t, &ia, &ib);
        mbar_wait(&bars->empty[st], (ph >> st) & 1u); ph ^= 1u << st;
        mbar_arrive_expect_tx(&bars->full[st], kTxBytes);
        const int kblk = kb0 + kb, ibk = i0 >> 7, jbk = j0 >> 7;
        const Image& A = P.A.img[ia];
        const Image& B = P.B.img[ib];
        bulk_copy_g2s(sA + st * kTileBytes, P.a_mn ? A.block(kblk, ibk) : A.block(ibk, kblk), kTileBytes, &bars->full[st]);
        bulk_copy_g2s(sB + st * kTileBytes, P.b_mn ? B.block(kblk, jbk) : B.block(jbk, kblk), kTileBytes, &bars->full[st]);
        st = st + 1 == STAGES ? 0 : st + 1;
      }
    }
  } else if (warp == 5) {
    uint32_t ph = 0;
    const uint32_t idesc = make_idesc(128, 128, P.a_mn, P.b_mn);
    int st = 0;
    for (int it = 0; it < nit; ++it) {
      mbar_wait(&bars->full[st], (ph >> st) & 1u); ph ^= 1u << st;
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA + st * kTileBytes), b_base = smem_u32(sB + st * kTileBytes);
      // K steps of this block that hold data (the image pads K to 128 with zeros: a 64-wide layer needs 4 of the 8)
      const int nks = min(8, (P.K - (kb0 + it / P.nterms) * 128 + 15) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (ks < nks) {
            const uint64_t ad = P.a_mn ? make_desc(a_base + ks * 256, 128, kPlane) : make_desc(a_base + ks * 2 * kPlane, kPlane, 128);
            const uint64_t bd = P.b_mn ? make_desc(b_base + ks * 256, 128, kPlane) : make_desc(b_base + ks * 2 * kPlane, kPlane, 128);
            mma_bf16_raw(tmem, ad, bd, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          }
        }
        mma_commit_raw(&bars->empty[st]);
        if (it == nit - 1) mma_commit_raw(&bars->done);
      }
      __syncwarp();
      st = st + 1 == STAGES ? 0 : st + 1;
    }
  } else if (nit > 0) {
    bars->bias[tid] = (P.bias && kz == 0 && j0 + tid < P.N) ? P.bias[j0 + tid] : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    mbar_wait_relaxed(&bars->done, 0);
    tc_fence_after();
    const int i = i0 + tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool reduce = P.ksplit > 1 || P.accumulate;
    float* stage = reinterpret_cast<float*>(sA) + warp * (kStageBytesPerWarp / 4);   // the operand ring is idle once `done` fired
    for (int g32 = 0; g32 < 128; g32 += 32) {
      if (j0 + g32 >= P.N) break;
      uint32_t r[32];
      tmem_ld32(tmem + lane_base + g32, r);
      tmem_ld_wait();
      if (!reduce && P.c_vec) {
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]) + bars->bias[g32 + e];
        store_group_coalesced(stage, v, P.C, P.ldc, i0 + warp * 32, j0 + g32, P.M, P.N, lane);
        continue;
      }
#pragma unroll
      for (int j4 = 0; j4 < 32; j4 += 4) {
        const int j = j0 + g32 + j4;
        float v[4] = {__uint_as_float(r[j4]), __uint_as_float(r[j4 + 1]), __uint_as_float(r[j4 + 2]),
                      __uint_as_float(r[j4 + 3])};
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] += bars->bias[g32 + j4 + e];
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
  }
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------
// A-stationary variant for the conv layers' forward GEMMs: M = all points of a branch (thousands of row tiles), K one
// block, N up to 8 column tiles.  The generic kernel above re-reads both operand tiles for every (output tile, term):
// 6 x 64 KB per 128 x 128 tile in the six-product mode -- 19.7 GB through L2 for one [819200, 128] x [128, 1024] layer,
// 2.9 ms against 0.6 ms of HBM time for its input and output.  Here a CTA owns a row tile: the NSPLIT images of its A
// tile are fetched once and stay in shared memory, the B image tiles (weights: L2-resident) stream through a ring --
// each is fetched once per (row tile, column tile) and meets every A image it pairs with -- and two TMEM accumulators
// alternate so the epilogue of column tile j runs under the MMAs of j + 1.
// ---------------------------------------------------------------------------------------------
constexpr int kAstatRing = 3;
constexpr int kAstatMaxN = 2048;
constexpr int kAstatStatBytes = 4 * 128 * 2 * 4;             // per epilogue warp: sum / sum of squares of a tile's 128 columns
template <int NSPLIT> constexpr size_t astat_smem_bytes() {
  return (size_t)(NSPLIT + kAstatRing) * kTileBytes + kAstatMaxN * 4 + 4 * kStageBytesPerWarp + kAstatStatBytes + 256;
}

struct AstatBars {
  uint64_t a_full, b_full[kAstatRing], b_empty[kAstatRing], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

template <int NSPLIT>
static __global__ void __launch_bounds__(kThreads, 1) tc_gemm_astat_kernel(const Params P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;                                        // NSPLIT tiles
  uint8_t* sB = smem + NSPLIT * kTileBytes;                  // ring
  float* sBias = reinterpret_cast<float*>(smem + (NSPLIT + kAstatRing) * kTileBytes);
  float* sStage = sBias + kAstatMaxN;                        // 4 warps x 32 x 36 floats
  float* sStat = sStage + 4 * (kStageBytesPerWarp / 4);      // [4 warps][128 cols][2]
  AstatBars* bars = reinterpret_cast<AstatBars*>(smem + (NSPLIT + kAstatRing) * kTileBytes + kAstatMaxN * 4 + 4 * kStageBytesPerWarp +
                                                 kAstatStatBytes);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int ib0 = blockIdx.x, i0 = ib0 * 128;
  const int ntn = (P.N + 127) >> 7;

  if (tid == 0) {
    mbar_init(&bars->a_full, 1);
    for (int i = 0; i < kAstatRing; ++i) { mbar_init(&bars->b_full[i], 1); mbar_init(&bars->b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(&bars->tmem_base, 256);
  for (int j = tid; j < ntn * 128; j += kThreads) sBias[j] = (P.bias && j < P.N) ? P.bias[j] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->a_full, NSPLIT * kTileBytes);
      for (int s = 0; s < NSPLIT; ++s) bulk_copy_g2s(sA + s * kTileBytes, P.A.img[s].block(ib0, 0), kTileBytes, &bars->a_full);
      uint32_t ph = (1u << kAstatRing) - 1u;
      int st = 0;
      for (int j = 0; j < ntn; ++j)
        for (int ib = NSPLIT - 1; ib >= 0; --ib) {           // smallest contributions first
          mbar_wait(&bars->b_empty[st], (ph >> st) & 1u); ph ^= 1u << st;
          mbar_arrive_expect_tx(&bars->b_full[st], kTileBytes);
          bulk_copy_g2s(sB + st * kTileBytes, P.B.img[ib].block(0, j), kTileBytes, &bars->b_full[st]);
          st = st + 1 == kAstatRing ? 0 : st + 1;
        }
    }
  } else if (warp == 5) {
    const uint32_t idesc = make_idesc(128, 128, 0, 1);
    const int nks = min(8, (P.K + 15) >> 4);                 // K steps that hold data (K <= 128 here)
    mbar_wait(&bars->a_full, 0);
    uint32_t phb = 0, phe = 3u;                              // accumulators start out free
    int st = 0;
    for (int j = 0; j < ntn; ++j) {
      const int acc = j & 1;
      mbar_wait(&bars->acc_empty[acc], (phe >> acc) & 1u); phe ^= 1u << acc;
      tc_fence_after();
      const uint32_t d = tmem + (uint32_t)acc * 128u;
      bool first = true;
      for (int ib = NSPLIT - 1; ib >= 0; --ib) {
        mbar_wait(&bars->b_full[st], (phb >> st) & 1u); phb ^= 1u << st;
        tc_fence_after();
        const uint32_t b_base = smem_u32(sB + st * kTileBytes);
        if (elect_one()) {
          for (int ia = NSPLIT - 1 - ib; ia >= 0; --ia) {    // the pairs (ia, ib) with ia + ib < NSPLIT
            const uint32_t a_base = smem_u32(sA + ia * kTileBytes);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              if (ks < nks)
                mma_bf16_raw(d, make_desc(a_base + ks * 2 * kPlane, kPlane, 128), make_desc(b_base + ks * 256, 128, kPlane), idesc,
                             (first && ks == 0) ? 0u : 1u);
            }
            first = false;
          }
          mma_commit_raw(&bars->b_empty[st]);
          if (ib == 0) mma_commit_raw(&bars->acc_full[acc]);
        }
        __syncwarp();
        first = false;
        st = st + 1 == kAstatRing ? 0 : st + 1;
      }
    }
  } else {
    const int i = i0 + tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t phf = 0;
    for (int j = 0; j < ntn; ++j) {
      const int acc = j & 1, j0 = j * 128;
      mbar_wait_relaxed(&bars->acc_full[acc], (phf >> acc) & 1u); phf ^= 1u << acc;
      tc_fence_after();
      for (int g32 = 0; g32 < 128; g32 += 32) {
        if (j0 + g32 >= P.N) break;
        uint32_t r[32];
        tmem_ld32(tmem + lane_base + (uint32_t)acc * 128u + g32, r);
        tmem_ld_wait();
        if (P.c_vec) {
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]) + sBias[j0 + g32 + e];
          if (P.stat_sum) {
            float cs[4], cq[4];
            store_group_coalesced_stats(sStage + warp * (kStageBytesPerWarp / 4), v, P.C, P.ldc, i0 + warp * 32, j0 + g32, P.M, P.N,
                                        lane, cs, cq);
            if (lane < 8) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                sStat[(warp * 128 + g32 + lane * 4 + e) * 2] = cs[e];
                sStat[(warp * 128 + g32 + lane * 4 + e) * 2 + 1] = cq[e];
              }
            }
          } else {
            store_group_coalesced(sStage + warp * (kStageBytesPerWarp / 4), v, P.C, P.ldc, i0 + warp * 32, j0 + g32, P.M, P.N, lane);
          }
          continue;
        }
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          const int jj = j0 + g32 + j4;
          float v[4] = {__uint_as_float(r[j4]), __uint_as_float(r[j4 + 1]), __uint_as_float(r[j4 + 2]),
                        __uint_as_float(r[j4 + 3])};
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] += sBias[jj + e];
          if (i < P.M && jj < P.N) {
            float* dst = P.C + (int64_t)i * P.ldc + jj;
            if (P.c_vec) {
              *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (jj + e < P.N) dst[e] = v[e];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
      if (P.stat_sum && P.c_vec) {
        // fold the four warps' partial sums of this column tile: one thread per column, two fp64 reductions per column
        // and row tile (fp32 partials over 128 rows; their rounding errors are independent across the thousands of row
        // tiles, so the totals carry ~1e-9 relative error)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int jj = j0 + tid;
        if (jj < P.N) {
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) { a0 += sStat[(w4 * 128 + tid) * 2]; a1 += sStat[(w4 * 128 + tid) * 2 + 1]; }
          atomicAdd(P.stat_sum + jj, (double)a0);
          atomicAdd(P.stat_sq + jj, (double)a1);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 256);
}

// C must be pre-zeroed by the caller when the launch reduces into it (accumulate, or K longer than one CTA's share:
// launch() clears it itself in the second case unless `accumulate` says C already holds a value to add to).
static int launch(Params p, cudaStream_t st, bool* stats_done = nullptr) {
  if (stats_done) *stats_done = false;
  static bool attr_set = false;
  if (!attr_set) {
    AN3D_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<3>()));
    AN3D_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<1>()));
    attr_set = true;
  }
  const int a_rows = p.a_mn ? p.K : p.M, a_cols = p.a_mn ? p.M : p.K, b_rows = p.b_mn ? p.K : p.N, b_cols = p.b_mn ? p.N : p.K;
  bool ok = p.M > 0 && p.N > 0 && p.K > 0 && p.C && p.A.n >= 1 && p.A.n == p.B.n && p.A.n <= kMaxSplit;
  for (int s = 0; ok && s < p.A.n; ++s)
    ok = p.A.img[s].g && p.B.img[s].g && p.A.img[s].rows == a_rows && p.A.img[s].cols == a_cols && p.B.img[s].rows == b_rows &&
         p.B.img[s].cols == b_cols;
  if (!ok) {
    set_error("tcg::launch: bad operands (shape %d x %d x %d)", p.M, p.N, p.K);
    return AN3D_ERR_INVALID;
  }
  p.nterms = num_terms(p.A.n);
  p.c_vec = (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && p.ldc % 4 == 0 && p.N % 4 == 0;
  // the conv layers' forward form: many row tiles, one K block, several column tiles -> A-stationary kernel
  if (p.a_mn == 0 && p.b_mn == 1 && p.K <= 128 && !p.accumulate && p.N > 128 && p.N <= kAstatMaxN && (p.M + 127) / 128 >= 148) {
    static bool astat_attr = false;
    if (!astat_attr) {
      AN3D_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_astat_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)astat_smem_bytes<1>()));
      AN3D_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_astat_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)astat_smem_bytes<2>()));
      AN3D_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_astat_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)astat_smem_bytes<3>()));
      astat_attr = true;
    }
    p.ksplit = 1;
    if (!p.c_vec || !p.stat_sq) p.stat_sum = p.stat_sq = nullptr;
    if (stats_done) *stats_done = p.stat_sum != nullptr;
    const unsigned grid = (unsigned)((p.M + 127) / 128);
    prof_mark(PROF_FC, true, st);
    if (p.A.n == 1) tc_gemm_astat_kernel<1><<<grid, kThreads, astat_smem_bytes<1>(), st>>>(p);
    else if (p.A.n == 2) tc_gemm_astat_kernel<2><<<grid, kThreads, astat_smem_bytes<2>(), st>>>(p);
    else tc_gemm_astat_kernel<3><<<grid, kThreads, astat_smem_bytes<3>(), st>>>(p);
    prof_mark(PROF_FC, false, st);
    AN3D_LAUNCH_CHECK();
    return AN3D_OK;
  }
  const int nkb = (p.K + 127) >> 7;
  const int tiles = ((p.M + 127) / 128) * ((p.N + 127) / 128);
  int ks = (nkb + kMaxKBlocksPerCta - 1) / kMaxKBlocksPerCta;          // accumulation length cap
  // few output tiles: slice K so the launch fills the SMs; the slices meet in fp32 reductions.
  //  * split modes (two / three images): always -- besides the fill, short accumulations are what keeps them at fp32
  //    grade: the tensor core's accumulate truncates, its error grows linearly with the chain (measured against fp64, six
  //    products: 5e-8 of max |A||B| with one K block per tile, 6e-7 with two, 2.6e-6 with sixteen);
  //  * one image (bf16 mode): only the accumulating form (wgrad).  The order of the reductions is not reproducible, and
  //    with bf16 roundings downstream a forward pass whose last bits change from run to run flips different arg-max
  //    bins (gradient cosine against the oracle wandered 0.60-0.67 on the default architecture; 0.688 every run now).
  if ((p.accumulate || p.A.n >= 2) && tiles < 148 && nkb > 1) ks = std::max(ks, std::min(nkb, (296 + tiles - 1) / tiles));
  int per = (nkb + ks - 1) / ks;
  ks = (nkb + per - 1) / per;                                           // no empty slices
  p.ksplit = ks;
  if (ks > 1 && !p.accumulate)
    AN3D_CUDA_CHECK(cudaMemset2DAsync(p.C, sizeof(float) * (size_t)p.ldc, 0, sizeof(float) * (size_t)p.N, (size_t)p.M, st));
  dim3 grid((p.M + 127) / 128, (p.N + 127) / 128, ks);
  prof_mark(PROF_FC, true, st);
  if (per * p.nterms <= 6) tc_gemm_kernel<1><<<grid, kThreads, smem_bytes<1>(), st>>>(p);
  else tc_gemm_kernel<3><<<grid, kThreads, smem_bytes<3>(), st>>>(p);
  prof_mark(PROF_FC, false, st);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

// ---------------------------------------------------------------------------------------------
// glue to the materialised path's plan (PlanF32::tc_split images per operand, scratch slots PlanF32::tcbuf)
// ---------------------------------------------------------------------------------------------
enum Slot { SLOT_X = 0, SLOT_DZ = 1, SLOT_W = 2 };

// layers with a dimension below 8 (the 3-wide first conv layer, 3-wide outputs) stay on the CUDA cores
inline bool use_tensor_cores(const PlanF32& p, int M, int N, int K) { return p.tc_split > 0 && std::min(M, std::min(N, K)) >= 8; }

// `keep`: pack into this buffer (p.tc_split images of fc_image_elems(rows, cols) elements) instead of the scratch slot --
// the training forward keeps the conv layers' input images for the backward pass
static int pack_slot(const PlanF32& p, int slot, const float* src, int64_t ld, int rows, int cols, const float* scale,
                     const float* shift, const float* mask, float mask_scale, SplitMat* out, cudaStream_t st,
                     __nv_bfloat16* keep = nullptr) {
  const int64_t elems = fc_image_elems(rows, cols);
  if (!keep && (elems * p.tc_split > p.tcbuf_elems[slot] || !p.tcbuf[slot])) {
    set_error("tcg::pack_slot: %d x %d does not fit image scratch %d", rows, cols, slot);
    return AN3D_ERR_WORKSPACE;
  }
  PackArgs a;
  a.src = src; a.ld = ld; a.rows = rows; a.cols = cols; a.scale = scale; a.shift = shift; a.mask = mask; a.mask_scale = mask_scale;
  a.nsplit = p.tc_split;
  for (int s = 0; s < p.tc_split; ++s) a.dst[s] = (keep ? keep : p.tcbuf[slot]) + s * elems;
  return pack(a, st, out);
}

// the images pack_slot(..., keep) left in `buf`
inline SplitMat kept_images(const PlanF32& p, const __nv_bfloat16* buf, int rows, int cols) {
  SplitMat m;
  m.n = p.tc_split;
  const int64_t elems = fc_image_elems(rows, cols);
  for (int s = 0; s < p.tc_split; ++s) { m.img[s].g = buf + s * elems; m.img[s].rows = rows; m.img[s].cols = cols; }
  return m;
}

}  // namespace tcg

// C[M,N] (+)= pro(A) * B (+ bias) with the operand conventions of GemmArgs (kernels_f32.cuh): on the tensor cores when the
// plan says so, else the CUDA-core SGEMM.  A goes to the input slot, B to the weight slot.
// stat_sum / stat_sq (optional, pre-zeroed [N] doubles): column sums of C and C^2 fused into the GEMM's epilogue where the
// kernel that runs supports it; *stats_done tells the caller whether they were produced.
static int gemm_mat(const PlanF32& p, const GemmArgs& g, bool ta, bool tb, cudaStream_t st, double* stat_sum = nullptr,
                    double* stat_sq = nullptr, bool* stats_done = nullptr, __nv_bfloat16* keep_a = nullptr) {
  if (stats_done) *stats_done = false;
  if (!tcg::use_tensor_cores(p, g.M, g.N, g.K)) return launch_gemm(g, ta, tb, st);
  tcg::Params q;
  AN3D_TRY(tcg::pack_slot(p, tcg::SLOT_X, g.A, g.lda, ta ? g.K : g.M, ta ? g.M : g.K, g.pro_scale, g.pro_shift, g.pro_mask,
                          g.pro_mask_scale, &q.A, st, keep_a));
  AN3D_TRY(tcg::pack_slot(p, tcg::SLOT_W, g.B, g.ldb, tb ? g.N : g.K, tb ? g.K : g.N, nullptr, nullptr, nullptr, 1.f, &q.B, st));
  q.a_mn = ta ? 1 : 0;
  q.b_mn = tb ? 0 : 1;
  q.C = g.C; q.ldc = g.ldc; q.M = g.M; q.N = g.N; q.K = g.K; q.bias = g.bias; q.accumulate = g.accumulate;
  q.stat_sum = stat_sum; q.stat_sq = stat_sq;
  return tcg::launch(q, st, stats_done);
}

}  // namespace an3d
