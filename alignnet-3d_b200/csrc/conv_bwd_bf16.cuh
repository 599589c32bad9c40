// Backward of the fused conv stack on the tensor cores (sm_100a, tcgen05).
//
// The max-pool makes the gradient w.r.t. the widest activation sparse (one row per cloud and
// channel) and training-mode BN makes it "sparse + affine in z3":
//     dz3[m,c] = s3[c] dy3[m,c] + p'[c] + q[c] r3[m,c],      r3 = a2 W3 (raw accumulator)
// so neither z3 nor dz3 ([M, C3], gigabytes) is ever materialised:
//     wgrad3 = A2^T S + sa2 (x) p' + (G2 W3) diag(q)          S = sparse s3*dy3, G2 = A2^T A2 (Gram)
//     da2    = S (W3)^T + u + A2 Gq                           Gq = W3 diag(q) W3^T, u = W3 p'
// Kernels here:
//   gram2_kernel  : the Gram matrix A2^T A2 and the column sums, contraction over points (MN-major operands)
//   t1_sparse_kernel : T1 = A2^T S, the sparse part of wgrad3
//   dgrad3_kernel : da2 -> dy2 (ReLU mask) + BN2 backward sums, per cloud
//   bwd_l2_kernel : dz2 -> wgrad2, da1 -> dy1 + BN1 backward sums
// A2 tiles travel between kernels as the exact shared-memory plane images (bulk copies).  dy2 travels TRANSPOSED
// (dy2_img_bytes below): the accumulators of dgrad3 / bwd_l2 are channel-major (thread = channel, registers = consecutive
// points), so an image whose 16-byte chunks hold 8 consecutive POINTS of one channel is written and re-read with 16-byte
// accesses, where the forward layout (chunk = 8 channels of one point) costs a 2-byte access per element.  Both kernels
// are bound by shared-memory bandwidth (the MMAs' operand reads alone take ~45 % of it), not by issue slots or the tensor
// pipe (profiles/r2_timeline_bwd.txt), so the epilogues' access width is what matters.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "umma.cuh"
#include "conv_fwd_bf16.cuh"

namespace an3d {
namespace convbwd {

using namespace umma;
using convfwd::plane_stride;

// -DAN3D_TIMELINE: CTA 0 of dgrad3 / bwd_l2 stamps clock64() at its roles' hand-over points and prints one line per item
// (tools/build_variants.sh ... -DAN3D_TIMELINE; read with tools/prof_step.py).  Never defined in the shipped build.
#ifdef AN3D_TIMELINE
static __device__ long long g_tl[12][64];
#define TL(slot, li) do { if (blockIdx.x == 0 && (li) < 64) g_tl[slot][li] = clock64(); } while (0)
#else
#define TL(slot, li) do { } while (0)
#endif
constexpr uint32_t kWHalfBytes = 128 * 64 * 2;   // one [128 rows][64 k] weight image
// Transposed tile image [point block of 8][128 channels][8 points] bf16: element (pt, k) at (pt / 8) * 2048 + k * 16 +
// (pt % 8) * 2.  As an MMA operand it is K-major for a contraction over points (rows = channels; LBO = 2048, SBO = 128)
// and MN-major for a contraction over channels (rows = K = channels, MN = points; LBO = 128, SBO = 2048).
constexpr uint32_t kTPlane = 2048;
__host__ __device__ inline uint32_t dy2_img_bytes(int PC) { return (uint32_t)(PC / 8) * kTPlane; }
constexpr uint32_t kPlaneW = 2048;

// =============================================================================================
// gram2: Gram[k, k'] = sum_pt a2[pt,k] a2[pt,k']  and  sa2[k] = sum_pt a2[pt,k]  over the saved A2 tile images of one
// (stage, branch): the layer-3 BN statistics of the forward pass and the dense part of wgrad3.  Contraction over
// points with MN-major operands straight from the images; the column sums ride along as a 129th output column (a
// constant 'ones' plane appended to every image buffer in shared memory), so nothing but the tensor cores ever
// touches the activations.  HBM-bound (it streams every image once): three images in flight per SM.
// =============================================================================================
struct Gram2Params {
  const uint8_t* a2_img;
  uint32_t img_bytes;
  int N, PC, npc, n_items, items_per_cta;
  float* parts;          // [CTA][128][132]: this CTA's Gram matrix (columns 0..127) and column sums (column 128)
};
constexpr int kGram2PartCols = 132;
constexpr int kGram2Threads = 192;       // warps 0-3 flush, warp 4 MMA, warp 5 loader
constexpr int kGram2Bufs = 3;
constexpr uint32_t kGram2Cols = 144;     // 128 Gram columns + the column sums + 15 zero columns

inline size_t gram2_smem_bytes(int PC) { return kGram2Bufs * 18 * (size_t)plane_stride(PC) + 256; }

struct Gram2Bars {
  uint64_t full[kGram2Bufs], empty[kGram2Bufs], done;
  uint32_t tmem_base;
};

static __global__ void __launch_bounds__(kGram2Threads, 1) gram2_kernel(const Gram2Params P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t plane = plane_stride(P.PC);
  const uint32_t buf_bytes = 18 * plane;
  Gram2Bars* bars = reinterpret_cast<Gram2Bars*>(smem + kGram2Bufs * buf_bytes);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int it_begin = min(P.n_items, (int)blockIdx.x * P.items_per_cta);
  const int it_end = min(P.n_items, it_begin + P.items_per_cta);
  const int n_local = it_end - it_begin;

  if (tid == 0) {
    for (int i = 0; i < kGram2Bufs; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->done, 1);
    fence_barrier_init();
  }
  // planes 16 / 17 of every buffer: channel 128 = 1 for every row, channels 129..143 = 0 (padding rows of an image
  // hold zero activations, so a one there adds nothing)
  for (int i = tid; i < kGram2Bufs * 2 * P.PC; i += kGram2Threads) {
    const int bidx = i / (2 * P.PC), r = i - bidx * 2 * P.PC;
    const int pl = r / P.PC, row = r - pl * P.PC;
    *reinterpret_cast<uint4*>(smem + bidx * buf_bytes + (16 + pl) * plane + row * 16) =
        make_uint4(pl == 0 ? 0x00003f80u : 0u, 0, 0, 0);
  }
  if (warp == 4) tmem_alloc(&bars->tmem_base, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 4) {
    if (n_local > 0) {
      const uint32_t idesc = make_idesc(128, kGram2Cols, 1, 1);
      uint32_t ph = 0;
      for (int li = 0; li < n_local; ++li) {
        // the range is walked BACKWARDS: the forward kernel (same item partition) wrote the images in ascending order a
        // moment ago, so the tail of every CTA's range is what may still sit in L2
        const int it = it_end - 1 - li;
        const int cloud = it / P.npc, pchunk = it - cloud * P.npc;
        const int nvalid = min(P.PC, P.N - pchunk * P.PC);
        const int NT = (nvalid + 15) & ~15;
        const int b = li % kGram2Bufs;
        mbar_wait(&bars->full[b], (ph >> b) & 1u); ph ^= 1u << b;
        tc_fence_after();
        if (elect_one()) {
          const uint64_t d = make_desc(smem_u32(smem + b * buf_bytes), 128, plane);
          for (int ks = 0; ks < NT / 16; ++ks)
            mma_bf16_raw(tmem, desc_advance(d, ks * 256), desc_advance(d, ks * 256), idesc, (li > 0 || ks > 0) ? 1u : 0u);
          mma_commit_raw(&bars->empty[b]);
          if (li == n_local - 1) mma_commit_raw(&bars->done);
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      uint32_t ph_e = (1u << kGram2Bufs) - 1u;
      for (int li = 0; li < n_local; ++li) {
        const int b = li % kGram2Bufs;
        mbar_wait_sleep(&bars->empty[b], (ph_e >> b) & 1u, 128u); ph_e ^= 1u << b;
        mbar_arrive_expect_tx(&bars->full[b], P.img_bytes);
        bulk_copy_g2s(smem + b * buf_bytes, P.a2_img + (size_t)(it_end - 1 - li) * P.img_bytes, P.img_bytes, &bars->full[b]);
      }
    }
  } else if (n_local > 0) {
    // ---- flush: TMEM -> vector reductions (lanes = k) ----
    mbar_wait_relaxed(&bars->done, 0);
    tc_fence_after();
    const int k = tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float* row = P.parts + ((size_t)blockIdx.x * 128 + k) * kGram2PartCols;
    for (int g16 = 0; g16 < 128; g16 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + lane_base + g16, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(row + g16 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    }
    uint32_t r[16];
    tmem_ld16(tmem + lane_base + 128, r);
    tmem_ld_wait();
    *reinterpret_cast<float4*>(row + 128) = make_float4(__uint_as_float(r[0]), 0.f, 0.f, 0.f);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

// =============================================================================================
// t1_sparse: T1[k, c] = sum over clouds of w[b,c] * a2[b, r*(b,c), k] on CUDA cores.
// S has exactly one non-zero per (cloud, channel), so the "GEMM" A2^T S is a gather-scale-accumulate
// with 1/N of the dense FLOPs.  CTA = (cloud range, quarter of the 128 k rows); it needs only 4 of the
// 16 planes of each saved A2 image (one bulk copy per item into a kT1Stages-deep ring).  THREAD = one
// channel: its (arg row, weight) pair comes straight from global memory into registers (coalesced,
// prefetched two items ahead), the 32 k-values of the selected row are four 16-byte shared-memory
// loads, and T1[32 k][c] lives in 32 registers until one flush with reductions at the end.
// =============================================================================================
struct T1Params {
  const uint8_t* a2_img;
  uint32_t img_bytes;
  const int32_t* gidx;
  const float* dyext;
  const float* s3;
  int B, N, PC, npc, C3, n_items, items_per_cta;
  float* t1;             // [128][C3], accumulated with reductions
};
constexpr int kT1Stages = 6;
inline int t1_threads(int C3) { return C3 < 1024 ? C3 : 1024; }
inline size_t t1_smem_bytes(int PC) { return (size_t)kT1Stages * 4 * plane_stride(PC) + 2 * kT1Stages * 8 + 64; }

static __global__ void __launch_bounds__(1024, 1) t1_sparse_kernel(const T1Params P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t plane = plane_stride(P.PC);
  const uint32_t qbytes = 4 * plane;
  uint8_t* sA = smem;                                                               // [kT1Stages][qbytes]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kT1Stages * qbytes);  // [kT1Stages]
  uint64_t* empty = full + kT1Stages;                                               // [kT1Stages]
  const int tid = threadIdx.x, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;
  const int kq = blockIdx.y;
  const int it_begin = min(P.n_items, (int)blockIdx.x * P.items_per_cta);
  const int it_end = min(P.n_items, it_begin + P.items_per_cta);
  const int n_local = it_end - it_begin;
  if (tid == 0) {
    for (int i = 0; i < kT1Stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], nwarps); }
    fence_barrier_init();
  }
  __syncthreads();
  if (n_local == 0) return;
  auto load_img = [&](int li) {
    const int st = li % kT1Stages;
    mbar_arrive_expect_tx(&full[st], qbytes);
    bulk_copy_g2s(sA + (size_t)st * qbytes, P.a2_img + (size_t)(it_begin + li) * P.img_bytes + (size_t)kq * qbytes, qbytes,
                  &full[st]);
  };
  if (tid == 0)
    for (int li = 0; li < min(n_local, kT1Stages); ++li) load_img(li);

  const int c = tid;                       // blockDim.x == C3 (<= 1024)
  const float s3c = P.s3[c];
  const bool one_item_per_cloud = P.npc == 1;      // the common case: no integer division per item
  // The look-ahead ring holds the RAW loaded values (gradient at the arg row, arg row): anything computed from them at
  // fetch time -- the validity test used to be -- makes the thread wait for the load it has just issued, and the ring
  // buys nothing (ncu: 45 % of this kernel's stall samples sat on these two loads).
  auto fetch = [&](int li, float& dy, int& idx) {
    dy = 0.f; idx = -1;
    if (li < n_local) {
      const int it = it_begin + li;
      const int cloud = one_item_per_cloud ? it : it / P.npc;
      dy = P.dyext[(size_t)cloud * P.C3 + c];
      idx = P.gidx[(size_t)cloud * P.C3 + c];
    }
  };
  auto resolve = [&](int li, float dy, int idx, float& w, int& row) {
    const int it = it_begin + li;
    const int pchunk = one_item_per_cloud ? 0 : it - (it / P.npc) * P.npc;
    const int p0 = pchunk * P.PC, nvalid = min(P.PC, P.N - p0);
    const int r = idx - p0;
    w = s3c * dy;
    row = (r >= 0 && r < nvalid && w != 0.f) ? r : -1;
  };
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  // (arg row, weight) pairs are fetched kT1Look items ahead into a register ring with compile-time slots (the item
  // loop is unrolled by the ring size): one or two items of look-ahead left the kernel waiting on these loads
  constexpr int kT1Look = 4;
  float wq[kT1Look];
  int rq[kT1Look];
#pragma unroll
  for (int u = 0; u < kT1Look; ++u) fetch(u, wq[u], rq[u]);
  for (int li0 = 0; li0 < n_local; li0 += kT1Look) {
#pragma unroll
    for (int u = 0; u < kT1Look; ++u) {
      const int li = li0 + u;
      if (li >= n_local) break;
      const int st = li % kT1Stages;
      float w0;
      int r0;
      resolve(li, wq[u], rq[u], w0, r0);
      fetch(li + kT1Look, wq[u], rq[u]);
      // refill the stage drained one iteration ago (every warp has arrived on its `empty` barrier by now, or will shortly)
      if (tid == 0 && li >= 1 && li - 1 + kT1Stages < n_local) {
        const int sp = (li - 1) % kT1Stages;
        mbar_wait(&empty[sp], (uint32_t)(((li - 1) / kT1Stages) & 1));
        load_img(li - 1 + kT1Stages);
      }
      mbar_wait(&full[st], (uint32_t)((li / kT1Stages) & 1));
      if (r0 >= 0) {
        const uint8_t* src = sA + (size_t)st * qbytes + (uint32_t)r0 * 16u;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          const uint4 v = *reinterpret_cast<const uint4*>(src + kc * plane);
          const uint32_t uu[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            acc[kc * 8 + 2 * h] = fmaf(w0, __uint_as_float(uu[h] << 16), acc[kc * 8 + 2 * h]);
            acc[kc * 8 + 2 * h + 1] = fmaf(w0, __uint_as_float(uu[h] & 0xffff0000u), acc[kc * 8 + 2 * h + 1]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
  }
  float* dst = P.t1 + (size_t)(kq * 32) * P.C3 + c;
#pragma unroll
  for (int j = 0; j < 32; ++j) atomicAdd(dst + (size_t)j * P.C3, acc[j]);
}

// =============================================================================================
// dgrad3: da2^T[k, pt] = Gq a2^T + u + W3 S^T ; dy2 = da2 * [a2 > 0] (written as the transposed image) ; sums for the
// BN2 backward
// =============================================================================================
struct Dg3Params {
  const uint8_t* a2_img;
  uint8_t* dy2_img;
  uint32_t img_bytes;
  const int32_t* gidx;
  const float* dyext;
  const float* s3;
  const __nv_bfloat16* gq_img;   // 2 halves [128 k][64 k']
  const __nv_bfloat16* w3n_img;  // C3/64 half-chunks [128 k][64 c]
  const float* uvec;             // [128]
  const float* gamma2;           // [128]
  const float* beta2;            // [128]
  int B, N, PC, npc, C3, n_items, items_per_cta;
  int wstages;                   // depth of the weight ring (3 or 4)
  double* red2;                  // [128][2] sum dy2, sum dy2 * xhat2
};
// warps 0-7 epilogue (two per TMEM lane quarter, each takes half of the point columns), warps 8-11 two scatter
// teams (team t fills S buffer t with the odd / even 64-channel half-chunks), 12 MMA, 13 weight loader, 14 tile loader
constexpr int kDg3Threads = 480;
constexpr int kDg3EpiThreads = 256;
constexpr int kDg3MmaWarp = 12;

// weight ring depth: the kernel is bound by the latency of this ring (every cloud re-streams all W3 half-chunks), and a
// fourth 16 KB stage is what still fits next to two A2 tiles and two scatter tiles at 208 points per item
inline int dg3_wstages(int PC) {
  return 2 * 16 * (size_t)plane_stride(PC) + 2 * 8 * (size_t)plane_stride(PC) + 4 * kWHalfBytes + 3 * 128 * 4 + 256 <= 227 * 1024 ? 4 : 3;
}
inline size_t dg3_smem_bytes(int PC) {
  return 2 * 16 * (size_t)plane_stride(PC) + 2 * 8 * (size_t)plane_stride(PC) + dg3_wstages(PC) * kWHalfBytes + 3 * 128 * 4 + 256;
}

struct Dg3Bars {
  uint64_t a2_full[2], a2_free[2], w_full[4], w_empty[4], sd_full[2], sd_empty[4], d_full[2], d_empty[2];
  uint32_t tmem_base;
};

static __global__ void __launch_bounds__(kDg3Threads, 1) dgrad3_kernel(const Dg3Params P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t plane = plane_stride(P.PC);
  const uint32_t sd_bytes = 8 * plane;
  auto sA2 = [&](int i) { return smem + (size_t)i * P.img_bytes; };        // (not pointer arrays: see bwd_l2_kernel)
  auto sSd = [&](int i) { return smem + 2 * (size_t)P.img_bytes + (size_t)i * sd_bytes; };
  uint8_t* sW = smem + 2 * P.img_bytes + 2 * sd_bytes;
  float* sU = reinterpret_cast<float*>(sW + (size_t)P.wstages * kWHalfBytes);
  float* sBeta = sU + 128;
  float* sIg = sBeta + 128;
  Dg3Bars* bars = reinterpret_cast<Dg3Bars*>(sIg + 128);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int it_begin = min(P.n_items, (int)blockIdx.x * P.items_per_cta);
  const int it_end = min(P.n_items, it_begin + P.items_per_cta);
  const int n_local = it_end - it_begin;
  const int nhc = P.C3 / 64;           // even: C3 is a multiple of 128
  const int nring = 2 + nhc;
  const uint32_t dy2_bytes = dy2_img_bytes(P.PC);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->a2_full[i], 1); mbar_init(&bars->a2_free[i], kDg3EpiThreads / 32);
      mbar_init(&bars->sd_full[i], 2);                                       // one arrival per warp of the team
      mbar_init(&bars->sd_empty[2 * i], 1); mbar_init(&bars->sd_empty[2 * i + 1], 1);
      mbar_init(&bars->d_full[i], 1); mbar_init(&bars->d_empty[i], kDg3EpiThreads / 32);
    }
    for (int i = 0; i < 4; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
    fence_barrier_init();
  }
  for (uint32_t i = tid * 16; i < 2 * sd_bytes; i += kDg3Threads * 16) *reinterpret_cast<uint4*>(sSd(0) + i) = make_uint4(0, 0, 0, 0);
  if (tid < 128) {
    sU[tid] = P.uvec[tid];
    sBeta[tid] = P.beta2[tid];
    const float gm = P.gamma2[tid];
    sIg[tid] = gm != 0.f ? 1.0f / gm : 0.f;
  }
  if (warp == kDg3MmaWarp) tmem_alloc(&bars->tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < 8) {
    // ---- epilogue: lane = channel k of a2; the warp pair of a lane quarter splits the point columns ----
    const int k = (warp & 3) * 32 + lane;
    const int half = warp >> 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ph_d = 0;                    // (phase bits in one scalar: an array indexed by li & 1 lives in local memory)
    double acc0 = 0.0, acc1 = 0.0;
    const float u = sU[k], beta = sBeta[k], ig = sIg[k];
    for (int li = 0; li < n_local; ++li) {
      const int it = it_begin + li;
      const int cloud = it / P.npc, pchunk = it - cloud * P.npc;
      const int nvalid = min(P.PC, P.N - pchunk * P.PC);
      const int NT = (nvalid + 15) & ~15;
      const int nh = ((NT >> 1) + 15) & ~15;
      const int pbeg = half ? min(nh, NT) : 0, pend = half ? NT : min(nh, NT);
      const int b = li & 1;
      mbar_wait_relaxed(&bars->d_full[b], (ph_d >> b) & 1u); ph_d ^= 1u << b;
      if (tid == 0) TL(0, li);
      tc_fence_after();
      const uint8_t* col = smem + (size_t)b * P.img_bytes + (k >> 3) * plane + (k & 7) * 2;
      uint8_t* gout = P.dy2_img + (size_t)it * dy2_bytes + (size_t)k * 16;
      // sums for the BN2 backward: s0 = sum dy, s1 = sum dy * xhat with xhat = (a - beta) / gamma, accumulated as
      // sa = sum dy * a (one FMA per element) and finished once per item
      float s0 = 0.f, sa = 0.f;
      // 16 points per step; the TMEM load of the next step is in flight while this one is processed
      auto step = [&](const uint32_t (&r)[16], int g16) {
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float a0 = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(col + (g16 + j) * 16));
          const float a1 = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(col + (g16 + j + 1) * 16));
          const float d0 = a0 > 0.f ? __uint_as_float(r[j]) + u : 0.f;   // (padding rows of the image hold a = 0)
          const float d1 = a1 > 0.f ? __uint_as_float(r[j + 1]) + u : 0.f;
          const uint32_t pk = convfwd::pack_bf16x2(d0, d1);
          const float q0 = __uint_as_float(pk << 16), q1 = __uint_as_float(pk & 0xffff0000u);   // dy as stored
          s0 += q0; s0 += q1;
          sa = fmaf(q0, a0, sa); sa = fmaf(q1, a1, sa);
          o[j >> 1] = pk;
        }
        // 16 consecutive points of this thread's channel = two 16-byte chunks of the transposed image, straight to global
        // memory (a warp writes 512 contiguous bytes per store): no shared-memory write, no bulk store
        uint8_t* dst = gout + (size_t)(g16 >> 3) * kTPlane;
        *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(dst + kTPlane) = make_uint4(o[4], o[5], o[6], o[7]);
      };
      {
        const uint32_t tb = tmem + lane_base + b * 256;
        uint32_t ra[16], rb[16];
        if (pbeg < pend) tmem_ld16(tb + pbeg, ra);
        for (int g16 = pbeg; g16 < pend; g16 += 32) {
          tmem_ld_wait();
          if (g16 + 16 < pend) tmem_ld16(tb + g16 + 16, rb);
          step(ra, g16);
          if (g16 + 16 < pend) {
            tmem_ld_wait();
            if (g16 + 32 < pend) tmem_ld16(tb + g16 + 32, ra);
            step(rb, g16 + 16);
          }
        }
      }
      const float s1 = (sa - beta * s0) * ig;
      acc0 += (double)s0; acc1 += (double)s1;
      if (tid == 0) TL(1, li);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars->d_empty[b]);
        mbar_arrive(&bars->a2_free[b]);      // this warp has read all it needs of the tile
      }
    }
    if (n_local > 0) {
      atomicAdd(P.red2 + 2 * k, acc0);
      atomicAdd(P.red2 + 2 * k + 1, acc1);
    }
  } else if (warp < 12) {
    // ---- scatter teams (next item's arg rows / weights prefetched into registers).  Team `team` owns S buffer
    // `team` and the half-chunks hc with hc % 2 == team, so the two buffers are filled concurrently. ----
    const int team = (warp - 8) >> 1;
    const int t = ((warp - 8) & 1) * 32 + lane;      // channel within the 64-channel half-chunk
    // each warp of a team owns one 32-channel half of the team's S buffer and its own `empty` barrier: the first half is
    // released as soon as the two MMAs that read it are done, so its refill overlaps the rest of the ring step
    uint64_t* my_empty = &bars->sd_empty[2 * team + ((warp - 8) & 1)];
    int prev_off = -1;
    uint32_t ph_e = 1;
    constexpr int kMaxHcT = 8;   // half-chunks per team: C3 <= 1024
    const int nhc_t = nhc >> 1;
    int cur_idx[kMaxHcT], nxt_idx[kMaxHcT];
    float cur_w[kMaxHcT], nxt_w[kMaxHcT];
    float s3r[kMaxHcT];        // BN3 scale of this thread's channels: loaded once, not once per item
#pragma unroll
    for (int h = 0; h < kMaxHcT; ++h) s3r[h] = h < nhc_t ? P.s3[(2 * h + team) * 64 + t] : 0.f;
    // the prefetched values stay RAW until they are used: a multiply at fetch time would wait for the load just issued
    auto fetch = [&](int li, int (&ix)[kMaxHcT], float (&wv)[kMaxHcT]) {
      const int cloud = (it_begin + li) / P.npc;
#pragma unroll
      for (int h = 0; h < kMaxHcT; ++h) {
        if (h < nhc_t) {
          const int c = (2 * h + team) * 64 + t;
          ix[h] = P.gidx[(size_t)cloud * P.C3 + c];
          wv[h] = P.dyext[(size_t)cloud * P.C3 + c];
        }
      }
    };
    uint8_t* sS = sSd(team);
    if (n_local > 0) fetch(0, cur_idx, cur_w);
    for (int li = 0; li < n_local; ++li) {
      const int it = it_begin + li;
      const int cloud = it / P.npc, pchunk = it - cloud * P.npc;
      const int p0 = pchunk * P.PC;
      const int nvalid = min(P.PC, P.N - p0);
      if (li + 1 < n_local) fetch(li + 1, nxt_idx, nxt_w);
#pragma unroll
      for (int h = 0; h < kMaxHcT; ++h) {
        if (h >= nhc_t) break;
        const int row = cur_idx[h] - p0;
        const float w = s3r[h] * cur_w[h];
        mbar_wait(my_empty, ph_e); ph_e ^= 1;
        if (prev_off >= 0) *reinterpret_cast<__nv_bfloat16*>(sS + prev_off) = __float2bfloat16_rn(0.f);
        if (row >= 0 && row < nvalid && w != 0.f) {
          const int off = (t >> 3) * plane + row * 16 + (t & 7) * 2;
          *reinterpret_cast<__nv_bfloat16*>(sS + off) = __float2bfloat16_rn(w);
          prev_off = off;
        } else {
          prev_off = -1;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->sd_full[team]);
      }
#pragma unroll
      for (int h = 0; h < kMaxHcT; ++h) { cur_idx[h] = nxt_idx[h]; cur_w[h] = nxt_w[h]; }
    }
  } else if (warp == kDg3MmaWarp) {
    if (n_local > 0) {
      // phase bits in scalar registers (dynamically indexed arrays would live in local memory), ring position by
      // compare-and-reset, each ring step's four MMAs and commits issued from one elected region: the scalar
      // instruction stream of this warp, not the tensor pipe, used to pace the kernel.
      uint32_t ph_a2 = 0, ph_sd = 0, ph_w = 0, ph_de = 3;
      int st = 0;
      const uint32_t w_base = smem_u32(sW);
      for (int li = 0; li < n_local; ++li) {
        const int it = it_begin + li;
        const int cloud = it / P.npc, pchunk = it - cloud * P.npc;
        const int nvalid = min(P.PC, P.N - pchunk * P.PC);
        const int NT = (nvalid + 15) & ~15;
        const int b = li & 1;
        (void)cloud;
        mbar_wait(&bars->d_empty[b], (ph_de >> b) & 1u); ph_de ^= 1u << b;
        if (lane == 0) TL(4, li);
        tc_fence_after();
        const uint32_t idesc = make_idesc(128, NT, 0, 0);
        const uint32_t d_tmem = tmem + b * 256;
        const uint64_t a2_desc = make_desc(smem_u32(sA2(b)), plane, 128);
        const uint64_t sd_desc0 = make_desc(smem_u32(sSd(0)), plane, 128), sd_desc1 = make_desc(smem_u32(sSd(1)), plane, 128);
        for (int r = 0; r < nring; ++r) {
          mbar_wait(&bars->w_full[st], (ph_w >> st) & 1u); ph_w ^= 1u << st;
          const uint64_t a_desc = make_desc(w_base + (uint32_t)st * kWHalfBytes, kPlaneW, 128);
          uint64_t b_desc;
          const int sb = r & 1;
          // ring order: the nhc scatter steps first, the two Gq steps last -- only those read the A2 tile, so its
          // (HBM-latency) load hides behind the scatter steps instead of heading the item's dependency chain
          const bool scatter_step = r < nhc;
          if (!scatter_step) {
            if (r == nhc) {
              if (lane == 0) TL(5, li);
              mbar_wait(&bars->a2_full[b], (ph_a2 >> b) & 1u); ph_a2 ^= 1u << b;
              if (lane == 0) TL(6, li);
            }
            b_desc = desc_advance(a2_desc, (r - nhc) * 8 * plane);
          } else {
            mbar_wait(&bars->sd_full[sb], (ph_sd >> sb) & 1u); ph_sd ^= 1u << sb;
            b_desc = sb ? sd_desc1 : sd_desc0;
          }
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              mma_bf16_raw(d_tmem, desc_advance(a_desc, ks * 2 * kPlaneW), desc_advance(b_desc, ks * 2 * plane), idesc,
                           (r > 0 || ks > 0) ? 1u : 0u);
              if (scatter_step && ks == 1) mma_commit_raw(&bars->sd_empty[2 * sb]);
            }
            if (scatter_step) mma_commit_raw(&bars->sd_empty[2 * sb + 1]);
            mma_commit_raw(&bars->w_empty[st]);
            if (r == nring - 1) mma_commit_raw(&bars->d_full[b]);
          }
          __syncwarp();
          if (++st == P.wstages) st = 0;
        }
      }
    }
  } else if (warp == kDg3MmaWarp + 1) {
    if (lane == 0 && n_local > 0) {
      uint32_t ph_e = 0xfu;                 // one phase bit per stage
      int wq = 0;
      for (int li = 0; li < n_local; ++li) {
        for (int r = 0; r < nring; ++r, ++wq) {
          const int st = wq % P.wstages;
          mbar_wait_relaxed(&bars->w_empty[st], (ph_e >> st) & 1u); ph_e ^= 1u << st;
          mbar_arrive_expect_tx(&bars->w_full[st], kWHalfBytes);
          const __nv_bfloat16* src = r < nhc ? P.w3n_img + (size_t)r * 8192 : P.gq_img + (size_t)(r - nhc) * 8192;
          bulk_copy_g2s(sW + (size_t)st * kWHalfBytes, src, kWHalfBytes, &bars->w_full[st]);
        }
      }
    }
  } else {
    if (lane == 0) {
      uint32_t ph_f = 3u;
      for (int li = 0; li < n_local; ++li) {
        const int b = li & 1;
        mbar_wait_relaxed(&bars->a2_free[b], (ph_f >> b) & 1u); ph_f ^= 1u << b;
        TL(7, li);
        mbar_arrive_expect_tx(&bars->a2_full[b], P.img_bytes);
        bulk_copy_g2s(sA2(b), P.a2_img + (size_t)(it_begin + li) * P.img_bytes, P.img_bytes, &bars->a2_full[b]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kDg3MmaWarp) tmem_dealloc(tmem, 512);
#ifdef AN3D_TIMELINE
  if (blockIdx.x == 0 && tid == 0) {
    const long long t0 = g_tl[7][0];
    for (int li = 0; li < min(n_local, 64); ++li)
      printf("DG3 C3=%d li=%d load_issue=%lld mma_start=%lld gq_wait=%lld a2_full=%lld d_full=%lld epi_end=%lld\n",
             P.C3, li, g_tl[7][li] - t0, g_tl[4][li] - t0, g_tl[5][li] - t0, g_tl[6][li] - t0, g_tl[0][li] - t0, g_tl[1][li] - t0);
  }
#endif
}

// =============================================================================================
// layer-2 backward: dz2 = s2 (dy2 - m0 - xhat2 m1) ; wgrad2 += a1^T dz2 ; da1 = dz2 W2^T ;
//                   dy1 = da1 * [a1 > 0] (+ BN1 backward sums); dy1 itself is never stored: layer 1 is linear in
//                   the 3 input coordinates, so its whole backward needs only sum_p dy1 * (1, x, y, z) per item and
//                   channel (l1sums; finished by bwd_l1_finish_kernel once the BN1 coefficients are known)
// =============================================================================================
struct L2Params {
  const float* pcs;
  const float* center;
  const float* angle;
  const uint8_t* dy2_img;
  uint32_t img_bytes;
  int B, N, PC, npc, n_items, items_per_cta;
  const float* w1f;  const float* c1f;      // folded layer 1 (forward recompute of a1)
  const float* W1;   const float* b1;       // raw layer-1 weights [3][64], bias
  const float* mean1; const float* inv1; const float* gamma1; const float* beta1;
  const __nv_bfloat16* w2t_img;             // forward image [128 k2][64 k1]
  const __nv_bfloat16* w2p_img;             // [128 k1 (zero padded)][128 k2]
  const float* b2; const float* mean2; const float* inv2; const float* s2;
  const float* coef2;                       // [128][2] m0, m1
  float* gW2;                               // [64][128]
  float* l1sums;                            // [n_items][64][4]: per item and channel sum_p dy1 * (1, x, y, z)
  double* red1;                             // [64][2]
};
// Two worker groups of 8 warps walk alternate items (ping-pong): while one group is in a CUDA-core phase
// (layer-1 recompute, dz2 / dy1 epilogues) the other one's MMAs, TMEM loads and barrier round trips are in
// flight, so the per-item dependency chain  a1 -> D2 -> dz2 -> (wgrad2, da1) -> dy1  no longer idles the SM.
constexpr int kL2Groups = 2;
constexpr int kL2GroupThreads = 256;
constexpr int kL2Threads = kL2Groups * kL2GroupThreads + 64;   // + MMA warp + loader warp
constexpr int kL2MmaWarp = kL2Groups * kL2GroupThreads / 32;
constexpr uint32_t kL2AccStride = 224;                         // TMEM columns per group accumulator (PC <= 208)
constexpr uint32_t kL2AccWG = 448;                             // wgrad2 accumulator [448, 512)

inline size_t l2_smem_bytes(int PC) {
  return kL2Groups * (8 * (size_t)plane_stride(PC) + dy2_img_bytes(PC)) + convfwd::kW2Bytes + 128 * 128 * 2 +
         kL2Groups * 256 * 4 * 4 + kL2Groups * 4 * 64 * 16 + (192 + 64 + 192 + 64 * 5 + 128 * 6) * 4 + 256;
}

struct L2Bars {
  uint64_t w_full, done;
  uint64_t dz_full[kL2Groups], dz_free[kL2Groups], a1_full[kL2Groups], d2_full[kL2Groups], dz_ready[kL2Groups],
      da_full[kL2Groups];
  uint32_t tmem_base;
  float xf[kL2Groups][8];
};

#ifdef AN3D_L2_SLEEPWAIT
#define L2_WAIT(bar, ph) mbar_wait_sleep(bar, ph, AN3D_L2_SLEEPWAIT)
#else
#define L2_WAIT(bar, ph) mbar_wait_relaxed(bar, ph)
#endif
static __global__ void __launch_bounds__(kL2Threads, 1) bwd_l2_kernel(const L2Params P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t plane = plane_stride(P.PC);
  // (functions of the group index, not pointer arrays: a runtime-indexed array would live in local memory)
  const uint32_t dz_bytes = dy2_img_bytes(P.PC);          // dy2 / dz2 tile, TRANSPOSED image (see the file header)
  auto sA1b = [&](int g) { return smem + (size_t)g * 8 * plane; };
  auto sDZb = [&](int g) { return smem + (size_t)16 * plane + (size_t)g * dz_bytes; };
  uint8_t* sW2T = smem + 16 * plane + kL2Groups * dz_bytes;
  uint8_t* sW2P = sW2T + convfwd::kW2Bytes;
  float4* sPtsAll = reinterpret_cast<float4*>(sW2P + 128 * 128 * 2);   // [groups][256] transformed points (x, y, z, -)
  float4* sRedAll = sPtsAll + kL2Groups * 256;                          // [groups][4 parts][64]
  // folded layer-1 weights as one float4 (wx, wy, wz, c) per channel: the recompute reads them with ONE broadcast
  // load per channel instead of four (256 LDS per thread and item before)
  float4* sW1f4 = sRedAll + kL2Groups * 256;                            // 64
  float* sW1 = reinterpret_cast<float*>(sW1f4 + 64);                    // 192
  float* sL1 = sW1 + 192;        // b1, mean1, inv1, gamma1, beta1 : 5 x 64
  float* sL2 = sL1 + 320;        // cx2 (= (b2-mean2)*inv2), inv2, s2, m0, m1, spare : 6 x 128
  L2Bars* bars = reinterpret_cast<L2Bars*>(sL2 + 768);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int it_begin = min(P.n_items, (int)blockIdx.x * P.items_per_cta);
  const int it_end = min(P.n_items, it_begin + P.items_per_cta);
  const int n_local = it_end - it_begin;

  if (tid == 0) {
    mbar_init(&bars->w_full, 1);
    mbar_init(&bars->done, 1);
    for (int i = 0; i < kL2Groups; ++i) {
      mbar_init(&bars->dz_full[i], 1); mbar_init(&bars->dz_free[i], 1);
      mbar_init(&bars->a1_full[i], kL2GroupThreads); mbar_init(&bars->d2_full[i], 1);
      mbar_init(&bars->dz_ready[i], kL2GroupThreads); mbar_init(&bars->da_full[i], 1);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 192; i += kL2Threads) sW1[i] = P.W1[i];
  for (int i = tid; i < 64; i += kL2Threads) {
    sW1f4[i] = make_float4(P.w1f[i], P.w1f[64 + i], P.w1f[128 + i], P.c1f[i]);
    sL1[i] = P.b1[i]; sL1[64 + i] = P.mean1[i]; sL1[128 + i] = P.inv1[i]; sL1[192 + i] = P.gamma1[i]; sL1[256 + i] = P.beta1[i];
  }
  for (int i = tid; i < 128; i += kL2Threads) {
    sL2[i] = (P.b2[i] - P.mean2[i]) * P.inv2[i];
    sL2[128 + i] = P.inv2[i]; sL2[256 + i] = P.s2[i]; sL2[384 + i] = P.coef2[2 * i]; sL2[512 + i] = P.coef2[2 * i + 1];
  }
  if (warp == kL2MmaWarp) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < kL2MmaWarp) {
    const int grp = warp >> 3;             // worker group: items grp, grp + 2, ...
    const int t = tid & (kL2GroupThreads - 1);   // 0..255 within the group
    const int k = t & 127, half = t >> 7;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t kD = grp * kL2AccStride;
    uint8_t* sA1 = sA1b(grp);
    uint8_t* sDZ = sDZb(grp);
    float4* sPts = sPtsAll + grp * 256;
    float4* sRed = sRedAll + grp * 256;
    int prev_it = -1;
    float* xf = bars->xf[grp];
    uint32_t ph = 0;                       // all per-item barriers of the group flip once per item
    double r0 = 0.0, r1 = 0.0;
    // register prefetch of the next item's transform (thread 0) and point (thread t): keeps the global
    // latency out of the barrier at the top of each item
    float pf_c[3] = {0.f, 0.f, 0.f}, pf_ang = 0.f, pf_p[3] = {0.f, 0.f, 0.f};
    auto prefetch = [&](int li) {
      const int it = it_begin + li;
      const int cloud = it / P.npc, pchunk = it - cloud * P.npc;
      const int p0 = pchunk * P.PC;
      const int nvalid = min(P.PC, P.N - p0);
      if (t == 0) {
        pf_c[0] = P.center[cloud * 3]; pf_c[1] = P.center[cloud * 3 + 1]; pf_c[2] = P.center[cloud * 3 + 2];
        pf_ang = P.angle ? P.angle[cloud] : 0.f;
      }
      if (t < nvalid) {
        const float* src = P.pcs + ((int64_t)cloud * P.N + p0 + t) * 3;
        pf_p[0] = src[0]; pf_p[1] = src[1]; pf_p[2] = src[2];
      }
    };
    if (grp < n_local) prefetch(grp);
    for (int li = grp; li < n_local; li += kL2Groups, ph ^= 1) {
      const int it = it_begin + li;
      const int cloud = it / P.npc, pchunk = it - cloud * P.npc;
      const int p0 = pchunk * P.PC;
      const int nvalid = min(P.PC, P.N - p0);
      const int NT = (nvalid + 15) & ~15;
      const int64_t row0 = (int64_t)cloud * P.N + p0;
      // barrier A: every thread of the group is past the previous item's dy1 phase (sRed complete, sPts / xf free)
      if (grp == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
      else asm volatile("bar.sync 2, 256;" ::: "memory");
      if (t < 64 && prev_it >= 0) {
        const float4 a = sRed[t], b = sRed[64 + t], c = sRed[128 + t], d = sRed[192 + t];
        reinterpret_cast<float4*>(P.l1sums)[(size_t)prev_it * 64 + t] =
            make_float4(a.x + b.x + c.x + d.x, a.y + b.y + c.y + d.y, a.z + b.z + c.z + d.z, a.w + b.w + c.w + d.w);
      }
      prev_it = it;
      if (t == 0) {
        float sn = 0.f, cs = 1.f;
        if (P.angle) sincosf(pf_ang, &sn, &cs);
        xf[0] = pf_c[0]; xf[1] = pf_c[1]; xf[2] = pf_c[2]; xf[3] = cs; xf[4] = sn;
      }
      // barrier B: xf visible; sRed has been flushed before this item's dy1 phase rewrites it
      if (grp == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
      else asm volatile("bar.sync 2, 256;" ::: "memory");
      if (t == 0) TL(0, li);
      const float cur_p[3] = {pf_p[0], pf_p[1], pf_p[2]};
      if (li + kL2Groups < n_local) prefetch(li + kL2Groups);
      // ---- recompute a1 (layer 1), one thread per point ----
      if (t < NT) {
        const int p = t;
        if (p < nvalid) {
          const float x0 = cur_p[0] - xf[0], y0 = cur_p[1] - xf[1], z = cur_p[2] - xf[2];
          const float x = x0 * xf[3] - y0 * xf[4], y = x0 * xf[4] + y0 * xf[3];
          sPts[p] = make_float4(x, y, z, 0.f);
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 w = sW1f4[c8 * 8 + j];
              v[j] = fmaf(x, w.x, fmaf(y, w.y, fmaf(z, w.z, w.w)));
            }
            uint4 q;   // (ReLU inside the conversion, as in the forward kernel: same bits)
            q.x = convfwd::pack_bf16x2_relu(v[0], v[1]); q.y = convfwd::pack_bf16x2_relu(v[2], v[3]);
            q.z = convfwd::pack_bf16x2_relu(v[4], v[5]); q.w = convfwd::pack_bf16x2_relu(v[6], v[7]);
            *reinterpret_cast<uint4*>(sA1 + c8 * plane + p * 16) = q;
          }
        } else {
          sPts[p] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) *reinterpret_cast<uint4*>(sA1 + c8 * plane + p * 16) = make_uint4(0, 0, 0, 0);
        }
      }
      if (t == 0) TL(1, li);
      fence_proxy_async_smem();
      mbar_arrive(&bars->a1_full[grp]);
      // ---- dz2 from the raw layer-2 accumulator (xhat2) and dy2 ----
      L2_WAIT(&bars->d2_full[grp], ph);
      if (t == 0) TL(2, li);
      L2_WAIT(&bars->dz_full[grp], ph);
      if (t == 0) TL(3, li);
      tc_fence_after();
      {
        // dz = s2 (dy - m0 - xhat m1), xhat = acc*inv2 + cx   ==   dy*cA + cB + acc*cC
        const float cA = sL2[256 + k];
        const float cB = -sL2[256 + k] * (sL2[384 + k] + sL2[512 + k] * sL2[k]);
        const float cC = -sL2[256 + k] * sL2[512 + k] * sL2[128 + k];
        uint8_t* colT = sDZ + k * 16;            // this channel's chunks: + (pt / 8) * kTPlane
        const int nh = ((NT >> 1) + 15) & ~15;
        const int pbeg = half ? nh : 0, pend = half ? NT : min(nh, NT);
        // (the padding test is compiled out of the common path: as a run-time `if` inside one body it became sixteen
        // predicated selects per step, 9 % of the kernel's instructions)
        auto step = [&](auto pad_tag, const uint32_t (&r)[16], int g16, uint8_t* c0, const uint4 v0, const uint4 v1) {
          constexpr bool kPad = decltype(pad_tag)::value;
          const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          uint32_t o[8];
          const int nreal = nvalid - g16;          // < 16 only in an item's last group (padding rows -> 0)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float z0 = fmaf(__uint_as_float(w[j] << 16), cA, fmaf(__uint_as_float(r[2 * j]), cC, cB));
            float z1 = fmaf(__uint_as_float(w[j] & 0xffff0000u), cA, fmaf(__uint_as_float(r[2 * j + 1]), cC, cB));
            if (kPad) {
              if (2 * j >= nreal) z0 = 0.f;
              if (2 * j + 1 >= nreal) z1 = 0.f;
            }
            o[j] = convfwd::pack_bf16x2(z0, z1);
          }
          *reinterpret_cast<uint4*>(c0) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(c0 + kTPlane) = make_uint4(o[4], o[5], o[6], o[7]);
        };
        const uint32_t tb = tmem + lane_base + kD;
        for (int g16 = pbeg; g16 < pend; g16 += 16) {
          uint32_t r[16];
          tmem_ld16(tb + g16, r);
          uint8_t* c0 = colT + (size_t)(g16 >> 3) * kTPlane;
          const uint4 v0 = *reinterpret_cast<const uint4*>(c0), v1 = *reinterpret_cast<const uint4*>(c0 + kTPlane);
          tmem_ld_wait();
          if (g16 + 16 <= nvalid) step(std::false_type{}, r, g16, c0, v0, v1);      // (warp-uniform)
          else step(std::true_type{}, r, g16, c0, v0, v1);
        }
      }
      if (t == 0) TL(4, li);
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&bars->dz_ready[grp]);
      // ---- dy1 = da1 * [a1 > 0], BN1 backward sums (channels k1 < 64 only) ----
      L2_WAIT(&bars->da_full[grp], ph);
      if (t == 0) TL(5, li);
      tc_fence_after();
      {
        // accumulator rows 64..127 duplicate rows 0..63: channel k1 = k & 63, and the four (lane half, warp
        // group) combinations each take a quarter of the point columns
        const int k1 = k & 63;
        const float wx = sW1[k1], wy = sW1[64 + k1], wz = sW1[128 + k1];
        const float b1 = sL1[k1], mu1 = sL1[64 + k1], inv1 = sL1[128 + k1];
        const int nq = ((NT >> 2) + 15) & ~15;
        const int part = (k >> 6) + 2 * half;
        const int pbeg = min(NT, part * nq), pend = part == 3 ? NT : min(NT, (part + 1) * nq);
        float s0 = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
        // ReLU mask of layer 1 straight from the recomputed a1 tile (still intact: the next item's recompute waits for
        // barrier A): one 2-byte load and a sign test instead of re-evaluating the layer and its BN per element.
        // Padding rows hold a1 = 0, so they drop out without a bounds test.
        const uint8_t* a1col = sA1 + (k1 >> 3) * plane + (k1 & 7) * 2;
        const uint32_t tb = tmem + lane_base + kD;
        for (int g16 = pbeg; g16 < pend; g16 += 16) {
          uint32_t r[16];
          tmem_ld16(tb + g16, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int p = g16 + j;
            const short a1bits = *reinterpret_cast<const short*>(a1col + p * 16);
            const float dy = a1bits > 0 ? __uint_as_float(r[j]) : 0.f;
            const float4 pt = sPts[p];
            s0 += dy;
            sx = fmaf(dy, pt.x, sx);
            sy = fmaf(dy, pt.y, sy);
            sz = fmaf(dy, pt.z, sz);
          }
        }
        sRed[part * 64 + k1] = make_float4(s0, sx, sy, sz);
        // BN1 backward sums: xhat is affine in (x, y, z), so sum dy*xhat follows from the four sums
        r0 += (double)s0;
        r1 += (double)(inv1 * (wx * sx + wy * sy + wz * sz + (b1 - mu1) * s0));
      }
      if (t == 0) TL(6, li);
      tc_fence_before();
    }
    if (n_local > 0) {
      if (grp < n_local) {
        if (grp == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
        else asm volatile("bar.sync 2, 256;" ::: "memory");
        if (t < 64) {
          const float4 a = sRed[t], b = sRed[64 + t], c = sRed[128 + t], d = sRed[192 + t];
          reinterpret_cast<float4*>(P.l1sums)[(size_t)prev_it * 64 + t] =
              make_float4(a.x + b.x + c.x + d.x, a.y + b.y + c.y + d.y, a.z + b.z + c.z + d.z, a.w + b.w + c.w + d.w);
        }
        atomicAdd(P.red1 + 2 * (k & 63), r0);
        atomicAdd(P.red1 + 2 * (k & 63) + 1, r1);
      }
      // ---- wgrad2 accumulator: lane = k2, 64 columns = k1 ----
      if (grp == 0 && half == 0) {
        mbar_wait_relaxed(&bars->done, 0);
        tc_fence_after();
        for (int g16 = 0; g16 < 64; g16 += 16) {
          uint32_t r[16];
          tmem_ld16(tmem + lane_base + kL2AccWG + g16, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(P.gW2 + (size_t)(g16 + j) * 128 + k, __uint_as_float(r[j]));
        }
        tc_fence_before();
      }
    }
  } else if (warp == kL2MmaWarp) {
    if (n_local > 0) {
      mbar_wait(&bars->w_full, 0);
      auto nt_of = [&](int li) {
        const int it = it_begin + li;
        const int cloud = it / P.npc, pchunk = it - cloud * P.npc;
        const int nvalid = min(P.PC, P.N - pchunk * P.PC);
        return (nvalid + 15) & ~15;
      };
      const uint64_t w2t_desc = make_desc(smem_u32(sW2T), kPlaneW, 128), w2p_desc = make_desc(smem_u32(sW2P), kPlaneW, 128);
      const uint64_t a1k_desc[2] = {make_desc(smem_u32(sA1b(0)), plane, 128), make_desc(smem_u32(sA1b(1)), plane, 128)};
      const uint64_t a1m_desc[2] = {make_desc(smem_u32(sA1b(0)), 128, plane), make_desc(smem_u32(sA1b(1)), 128, plane)};
      // dz2 tile (transposed image): K-major A operand of wgrad2 (contraction over points), MN-major B operand of da1
      const uint64_t dzp_desc[2] = {make_desc(smem_u32(sDZb(0)), kTPlane, 128), make_desc(smem_u32(sDZb(1)), kTPlane, 128)};
      const uint64_t dzc_desc[2] = {make_desc(smem_u32(sDZb(0)), 128, kTPlane), make_desc(smem_u32(sDZb(1)), 128, kTPlane)};
      // every group of MMAs and its commits is issued from one elected region (see umma.cuh)
      auto issue_d2 = [&](int li) {
        const int g = li & 1;
        tc_fence_after();
        const uint32_t idesc = make_idesc(128, nt_of(li), 0, 0);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_bf16_raw(tmem + g * kL2AccStride, desc_advance(w2t_desc, ks * 2 * kPlaneW),
                         desc_advance(g ? a1k_desc[1] : a1k_desc[0], ks * 2 * plane), idesc, ks > 0);
          mma_commit_raw(&bars->d2_full[g]);
        }
        __syncwarp();
      };
      bool wg_started = false;
      auto issue_bwd = [&](int li) {
        const int g = li & 1;
        const int NT = nt_of(li);
        tc_fence_after();
        const uint32_t idesc_w = make_idesc(128, 64, 0, 1);
        const uint32_t idesc = make_idesc(128, NT, 0, 1);
        if (elect_one()) {
          const uint64_t dzp = g ? dzp_desc[1] : dzp_desc[0], a1m = g ? a1m_desc[1] : a1m_desc[0];
          const uint64_t dzc = g ? dzc_desc[1] : dzc_desc[0];
          for (int ks = 0; ks < NT / 16; ++ks)
            mma_bf16_raw(tmem + kL2AccWG, desc_advance(dzp, ks * 2 * kTPlane), desc_advance(a1m, ks * 256), idesc_w,
                         (wg_started || ks > 0) ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            mma_bf16_raw(tmem + g * kL2AccStride, desc_advance(w2p_desc, ks * 2 * kPlaneW), desc_advance(dzc, ks * 256),
                         idesc, ks > 0);
          mma_commit_raw(&bars->da_full[g]);
          mma_commit_raw(&bars->dz_free[g]);
        }
        wg_started = true;
        __syncwarp();
      };
      // fixed service order D2(2q), D2(2q+1), BWD(2q), BWD(2q+1): consistent with each group's own sequence
      for (int q = 0; q < n_local; q += 2) {
        mbar_wait(&bars->a1_full[0], (uint32_t)((q >> 1) & 1));
        issue_d2(q);
        if (q + 1 < n_local) { mbar_wait(&bars->a1_full[1], (uint32_t)((q >> 1) & 1)); issue_d2(q + 1); }
        mbar_wait(&bars->dz_ready[0], (uint32_t)((q >> 1) & 1));
        issue_bwd(q);
        if (q + 1 < n_local) { mbar_wait(&bars->dz_ready[1], (uint32_t)((q >> 1) & 1)); issue_bwd(q + 1); }
      }
      mma_commit(&bars->done);
    }
  } else {
    if (lane == 0 && n_local > 0) {
      mbar_arrive_expect_tx(&bars->w_full, convfwd::kW2Bytes + 128 * 128 * 2);
      bulk_copy_g2s(sW2T, P.w2t_img, convfwd::kW2Bytes, &bars->w_full);
      bulk_copy_g2s(sW2P, P.w2p_img, 128 * 128 * 2, &bars->w_full);
      for (int li = 0; li < n_local; ++li) {
        const int g = li & 1;
        mbar_wait_relaxed(&bars->dz_free[g], (uint32_t)(((li >> 1) & 1) ^ 1));
        mbar_arrive_expect_tx(&bars->dz_full[g], dz_bytes);
        bulk_copy_g2s(sDZb(g), P.dy2_img + (size_t)(it_begin + li) * dz_bytes, dz_bytes, &bars->dz_full[g]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kL2MmaWarp) tmem_dealloc(tmem, 512);
#ifdef AN3D_TIMELINE
  if (blockIdx.x == 0 && tid == 0) {
    const long long t0 = g_tl[0][0];
    for (int li = 0; li < min(n_local, 64); ++li)
      printf("L2 li=%d start=%lld a1_done=%lld d2_full=%lld dz_full=%lld dz_done=%lld da_full=%lld dy1_done=%lld\n", li,
             g_tl[0][li] - t0, g_tl[1][li] - t0, g_tl[2][li] - t0, g_tl[3][li] - t0, g_tl[4][li] - t0, g_tl[5][li] - t0,
             g_tl[6][li] - t0);
  }
#endif
}

}  // namespace convbwd
}  // namespace an3d
