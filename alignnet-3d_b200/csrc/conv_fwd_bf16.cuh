// Fused PointNet conv stack (3 -> 64 -> 128 -> C3, BN + ReLU folded, max-pool) for sm_100a.
//
// conv_stack_fwd_kernel: one persistent CTA per SM walks a contiguous range of work items (item = one cloud, or
// one <=256-point chunk of a cloud).  Per item:
//   front-end warps : recentre/rotate the points, layer 1 (K = 3, CUDA-core FFMA) -> bf16 A1 tile in shared
//                     memory; after the layer-2 MMA, TMEM -> BN affine + ReLU -> bf16 A2 tile
//   MMA warp        : tcgen05.mma  D2[pts, 128 ch] = A1 W2        (K = 64; POINT-major: lane = point)
//                     tcgen05.mma  D3[128 ch, pts] = W3^T[chunk] A2^T (K = 128; CHANNEL-major) per 128-channel
//                     chunk, accumulators in TMEM, two point-halves double-buffered against the epilogue
//   back-end warps  : TMEM -> per-channel running max over the cloud's points (one thread owns one channel, so
//                     the max-pool needs no cross-thread reduction)
//   loader thread   : bulk async copies (TMA unit) of the pre-packed W2^T / W3^T images
// The two layers use opposite accumulator orientations on purpose.  Layer 3 is channel-major (weights = MMA "A"
// operand) so that the max over points is a per-thread reduction.  Layer 2 is point-major (activations = MMA "A"
// operand; the very same shared-memory images serve both roles) so that its epilogue thread owns one POINT and
// writes the next layer's operand tile in 16-byte pieces (8 channels of one point are contiguous in the plane
// layout): ~3 instructions per element instead of ~7 with 2-byte stores.  The kernel is bound by instruction
// issue, not by the tensor pipe or TMEM bandwidth (profiles/r2_tmem_microbench.txt: >= 600 B/clk/SM of TMEM
// read-back is available), so instruction count per accumulator element is what the design minimises.
// W3 is sign-folded with sign(gamma) at pack time so that the pooled extreme of the raw accumulator is always a
// max (BN + ReLU are monotone).
//
// conv_stats2_kernel: layer-2 BATCH statistics without running layer 2 (see below); a light kernel, several CTAs
// per SM.
#pragma once
#include <algorithm>

#include "common.cuh"
#include "umma.cuh"

namespace an3d {
namespace convfwd {

using namespace umma;

constexpr int kThreads = 576;          // warps 0-3 / 10-13 back-end, 4-7 / 14-17 front-end, 8 MMA, 9 weight loader
// Back-end warps per TMEM lane quarter.  The max-reduction is ALU-pipe bound (LOP3 and FMNMX3 issue at half rate; the
// microbenchmark drains 112 columns in 621 cycles with 8 warps and in 490 with 16), but a third warp per quarter
// (704 threads, register cap 80 instead of 96, spills in the front end) made the kernel SLOWER on B200:
// 1.114 -> 1.212 ms per c3 step.  Two it is.
constexpr int kBackWarpsPerQuarter = 2;
constexpr int kFrontThreads = 256;
constexpr int kMaxPC = 256;            // points per item
constexpr uint32_t kW2Bytes = 128 * 64 * 2;
constexpr uint32_t kW3ChunkBytes = 128 * 128 * 2;
constexpr uint32_t kPlaneW2 = 128 * 16;   // plane stride of the weight images (rows = 128 channels)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemAcc0 = 0, kTmemAcc1 = 128, kTmemD2 = 256;   // D2: two point-major tiles of 128 columns

enum Mode { MODE_FULL_TRAIN = 1, MODE_FULL_EVAL = 2 };

// -DAN3D_TIMELINE: CTA 0 stamps clock64() at the front end's hand-over points and prints one line per item (see
// conv_bwd_bf16.cuh).  Never defined in the shipped build.
#ifdef AN3D_TIMELINE
static __device__ long long g_ftl[8][64];
static __device__ long long g_mtl[4][256];    // MMA warp, per chunk: loop top, weights ready, accumulator half 0 free, half 1 free
#define FTL(slot, li) do { if (blockIdx.x == 0 && (li) < 64) g_ftl[slot][li] = clock64(); } while (0)
#else
#define FTL(slot, li) do { } while (0)
#endif

struct Params {
  const float* pcs;       // [B, N, 3] raw points of this branch
  const float* center;    // [B, 3] subtracted from the points
  const float* angle;     // [B] rotation about z applied after recentring, or nullptr
  int B, N;
  int PC, npc;            // points per item (multiple of 16, <= 256) and items per cloud
  int item_begin_stride;  // items per CTA
  int n_items;
  const float* w1f;       // [3][64] layer-1 weights with the BN scale folded in
  const float* c1f;       // [64]    folded bias
  const __nv_bfloat16* w2t_img;  // plane image of W2^T [128 ch][64 k]
  const float* s2;        // [128] BN scale of layer 2
  const float* t2f;       // [128] folded shift (includes the conv bias)
  const __nv_bfloat16* w3t_img;  // nchunk plane images of (sign-folded) W3^T [128 ch][128 k]
  int nchunk;             // C3 / 128
  int nstages;            // W3 ring depth (2 or 3)
  uint32_t* zext;         // [B][C3] ordered-uint packed max of the raw layer-3 accumulator
  float* gram1;           // stats2 kernel: [CTA][64][80] = A1^T [A1 | 1 | 0] over the CTA's items (Gram matrix and column
                          // sums of the bf16 layer-1 activations; gives the layer-2 batch statistics in closed form)
  __nv_bfloat16* a2_img;  // optional: per item, the A2 tile exactly as staged in shared memory (16 planes),
                          // saved for the backward kernels (bulk store, TMA unit)
  uint32_t idx_mask;      // low mantissa bits that carry the arg-max point index (training)
  uint32_t not15;         // ~15u, passed through a register on purpose: `(x & reg) | imm` is ONE LOP3, while
                          // `(x & imm) | imm` costs two -- and this expression runs once per accumulator element
};

__host__ __device__ inline uint32_t plane_stride(int rows) { return (uint32_t)rows * 16u + 16u; }

inline size_t smem_bytes(int PC, int nstages) {
  const size_t a2 = 16 * (size_t)plane_stride(PC);
  return 2 * a2 + kW2Bytes + (size_t)nstages * kW3ChunkBytes + 2048 /*w1f,c1f,bn2*/ + 256 /*barriers*/ + 128;
}

__device__ __forceinline__ uint32_t to_ordered(uint32_t bits) {
  return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// bf16x2(max(lo, 0), max(hi, 0)) in ONE conversion: the ReLU rides on the pack (cvt's .relu clamps negative results to
// +0), which takes an FMNMX per element off the ALU pipe -- the pipe this kernel's front end is bound by.
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

struct Barriers {
  uint64_t w2_full;
  uint64_t w3_full[3];
  uint64_t w3_empty[3];
  uint64_t a1_full;
  uint64_t d2_full;
  uint64_t a2_full[2];
  uint64_t a2_empty[2];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  float xf[16];  // per-item transform (double-buffered): cx, cy, cz, cos, sin
};

// (x & mask) | Q as ONE LOP3: the immediate has to sit in the instruction's second source slot and the mask in a
// register (ptxas emits an AND and an OR for the C expression, whichever way it is written: 23 % of the kernel's
// instructions were those two).  Runs once per accumulator element of the training pass.
template <int Q>
__device__ __forceinline__ float tag4(uint32_t x, uint32_t mask) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEC;" : "=r"(d) : "r"(x), "n"(Q), "r"(mask));   // (a & c) | b
  return __uint_as_float(d);
}

// Running max over one 16-column group of an accumulator row.  Training mode carries the arg-max point
// index in the low mantissa bits: the column-in-group (an immediate) goes into the low 4 bits of every
// element, the group's base index is spliced in only if the group wins.  The back-end warps are the kernel's
// critical resource (two per scheduler, ~80 % busy in the ncu samples), so this routine is written for them:
// two independent max chains per group (a single chain of eight dependent 3-input maxima left the warp waiting on
// its own previous instruction), the padding columns of a cloud's last group overwritten with -FLT_MAX by sixteen
// predicated moves instead of a 16-way branchy tail, masks hoisted into registers by the caller.
template <int MODE>
__device__ __forceinline__ void reduce_group(uint32_t* r, int col0, int nvalid, int p0, uint32_t keep_mask,
                                             uint32_t not15, float& m) {
  const int k = nvalid - col0;                     // warp-uniform; < 16 only in the last group of a cloud
  if (k < 16) {
#pragma unroll
    for (int q = 1; q < 16; ++q)
      if (q >= k) r[q] = 0xff7fffffu;              // -FLT_MAX: loses against every real accumulator (k >= 1 always)
  }
  float g0, g1;
  if (MODE == MODE_FULL_TRAIN) {
    g0 = fmax3(tag4<0>(r[0], not15), tag4<1>(r[1], not15), tag4<2>(r[2], not15));
    g1 = fmax3(tag4<8>(r[8], not15), tag4<9>(r[9], not15), tag4<10>(r[10], not15));
    g0 = fmax3(g0, tag4<3>(r[3], not15), tag4<4>(r[4], not15));
    g1 = fmax3(g1, tag4<11>(r[11], not15), tag4<12>(r[12], not15));
    g0 = fmax3(g0, tag4<5>(r[5], not15), tag4<6>(r[6], not15));
    g1 = fmax3(g1, tag4<13>(r[13], not15), tag4<14>(r[14], not15));
    const float gm = fmax3(g0, g1, fmaxf(tag4<7>(r[7], not15), tag4<15>(r[15], not15)));
    if (gm > m) m = __uint_as_float((__float_as_uint(gm) & keep_mask) | (uint32_t)(p0 + col0));
  } else {
    g0 = fmax3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
    g1 = fmax3(__uint_as_float(r[8]), __uint_as_float(r[9]), __uint_as_float(r[10]));
    g0 = fmax3(g0, __uint_as_float(r[3]), __uint_as_float(r[4]));
    g1 = fmax3(g1, __uint_as_float(r[11]), __uint_as_float(r[12]));
    g0 = fmax3(g0, __uint_as_float(r[5]), __uint_as_float(r[6]));
    g1 = fmax3(g1, __uint_as_float(r[13]), __uint_as_float(r[14]));
    m = fmax3(m, fmax3(g0, g1, __uint_as_float(r[7])), __uint_as_float(r[15]));
  }
}

// item geometry shared by all roles
struct Item {
  int cloud, p0, nvalid, NT;
};
__device__ __forceinline__ Item item_of(const Params& P, int it) {
  Item I;
  I.cloud = it / P.npc;
  const int pchunk = it - I.cloud * P.npc;
  I.p0 = pchunk * P.PC;
  I.nvalid = min(P.PC, P.N - I.p0);
  I.NT = (I.nvalid + 15) & ~15;
  return I;
}

// Layer 1 for ONE point and all 64 channels: the thread that prefetched the point keeps it in registers, recentres /
// rotates it once and walks the 8 channel groups (= planes of the A1 tile) with the folded weights read as broadcast
// 16-byte shared-memory loads.  `row` = this point's 16-byte slot in plane 0 of the tile.  (Until the third session of
// round 2 a WARP owned a channel group for all points, weights in registers, three points in flight per lane: every warp
// then repeated the recentre / rotate of every point and the raw points went through a shared-memory staging buffer --
// ~570 warp instructions per front-end warp and cloud against ~310 now, the same arithmetic in the same order, so the
// A1 tile is the same bit for bit.  Same-box A/B: forward kernels 1.061 -> 0.974 ms per c3 step, statistics pass
// 0.249 -> 0.240, step 4.596 -> 4.465 ms; profiles/r2_ab_variants.txt (14).)
__device__ __forceinline__ void layer1_point(float px, float py, float pz, bool real, float cx, float cy, float cz, float cs,
                                             float sn, const float* __restrict__ sW1f, const float* __restrict__ sC1f,
                                             uint8_t* row, uint32_t plane) {
  if (!real) {                       // padding rows of the tile (only the cloud's last warp diverges)
#pragma unroll
    for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(row + (size_t)g * plane) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const float x0 = px - cx, y0 = py - cy, z = pz - cz;
  const float x = x0 * cs - y0 * sn, y = x0 * sn + y0 * cs;
  // (Loading the weights of group g + 1 before group g's tile store -- the compiler keeps loads written after a store it
  // cannot disambiguate behind it -- made both kernels ~1.5 % faster in isolation and the STEP 0.4 % slower: at 94 instead
  // of 75 registers per thread the forward CTAs leave no room for the other branch's small kernels to co-reside.
  // profiles/r2_ab_variants.txt (15): rejected.)
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float4 xa = *reinterpret_cast<const float4*>(sW1f + g * 8), xb = *reinterpret_cast<const float4*>(sW1f + g * 8 + 4);
    const float4 ya = *reinterpret_cast<const float4*>(sW1f + 64 + g * 8), yb = *reinterpret_cast<const float4*>(sW1f + 64 + g * 8 + 4);
    const float4 za = *reinterpret_cast<const float4*>(sW1f + 128 + g * 8), zb = *reinterpret_cast<const float4*>(sW1f + 128 + g * 8 + 4);
    const float4 ca = *reinterpret_cast<const float4*>(sC1f + g * 8), cb = *reinterpret_cast<const float4*>(sC1f + g * 8 + 4);
    const float v0 = fmaf(x, xa.x, fmaf(y, ya.x, fmaf(z, za.x, ca.x))), v1 = fmaf(x, xa.y, fmaf(y, ya.y, fmaf(z, za.y, ca.y)));
    const float v2 = fmaf(x, xa.z, fmaf(y, ya.z, fmaf(z, za.z, ca.z))), v3 = fmaf(x, xa.w, fmaf(y, ya.w, fmaf(z, za.w, ca.w)));
    const float v4 = fmaf(x, xb.x, fmaf(y, yb.x, fmaf(z, zb.x, cb.x))), v5 = fmaf(x, xb.y, fmaf(y, yb.y, fmaf(z, zb.y, cb.y)));
    const float v6 = fmaf(x, xb.z, fmaf(y, yb.z, fmaf(z, zb.z, cb.z))), v7 = fmaf(x, xb.w, fmaf(y, yb.w, fmaf(z, zb.w, cb.w)));
    *reinterpret_cast<uint4*>(row + (size_t)g * plane) =
        make_uint4(pack_bf16x2_relu(v0, v1), pack_bf16x2_relu(v2, v3), pack_bf16x2_relu(v4, v5), pack_bf16x2_relu(v6, v7));
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) conv_stack_fwd_kernel(const Params P) {
  extern __shared__ __align__(128) uint8_t smem[];
#ifdef AN3D_TIMELINE
  if (threadIdx.x == 0) FTL(6, 0);                    // kernel entry
#endif
  const uint32_t plane2 = plane_stride(P.PC);         // A2 planes: 16 of them (K = 128)
  const uint32_t plane1 = plane2;                       // A1 uses the same row pitch, 8 planes (K = 64)
  const uint32_t a2_bytes = 16 * plane2;
  auto a2buf = [&](int b) { return smem + (size_t)b * a2_bytes; };   // (no pointer array: it would live in local memory)
  uint8_t* sW2 = smem + 2 * a2_bytes;
  uint8_t* sW3 = sW2 + kW2Bytes;
  float* sW1f = reinterpret_cast<float*>(sW3 + (size_t)P.nstages * kW3ChunkBytes);  // [3][64]
  float* sC1f = sW1f + 192;
  float4* sBn2 = reinterpret_cast<float4*>(sC1f + 64);       // [64] channel pairs: (scale, shift) of 2c, (scale, shift) of 2c+1
  Barriers* bars = reinterpret_cast<Barriers*>(sBn2 + 64);

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int it_begin = min(P.n_items, (int)blockIdx.x * P.item_begin_stride);
  const int it_end = min(P.n_items, it_begin + P.item_begin_stride);
  const int n_local = it_end - it_begin;

  if (tid == 0) {
    mbar_init(&bars->w2_full, 1);
    for (int i = 0; i < 3; ++i) { mbar_init(&bars->w3_full[i], 1); mbar_init(&bars->w3_empty[i], 1); }
    mbar_init(&bars->a1_full, kFrontThreads);
    mbar_init(&bars->d2_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->a2_full[i], kFrontThreads);
      mbar_init(&bars->a2_empty[i], 1);
      mbar_init(&bars->acc_full[i], 1);
      mbar_init(&bars->acc_empty[i], kBackWarpsPerQuarter * 128);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 192; i += kThreads) sW1f[i] = P.w1f[i];
  for (int i = tid; i < 64; i += kThreads) sC1f[i] = P.c1f[i];
  for (int i = tid; i < 64; i += kThreads) sBn2[i] = make_float4(P.s2[2 * i], P.t2f[2 * i], P.s2[2 * i + 1], P.t2f[2 * i + 1]);
  if (warp == 8) tmem_alloc(&bars->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
#ifdef AN3D_TIMELINE
  if (threadIdx.x == 0) FTL(6, 1);                    // prologue done
#endif

  if ((warp >= 4 && warp < 8) || (warp >= 14 && warp < 18)) {
    // ================================ front-end ================================
    // 8 warps.  Layer 1: thread f owns point f.  Layer-2 epilogue: warp (quarter, half)
    // reads TMEM lanes [32 quarter, +32) = points of a 128-point tile, and 64 of the 128 channel columns.
    const int fgroup = warp >= 14 ? 1 : 0;
    const int quarter = warp & 3;
    const int f = fgroup * 128 + quarter * 32 + lane;   // 0..255: the point this thread prefetches and runs layer 1 for
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    uint32_t ph_d2 = 0, ph_a2e = 0;     // (phase bits in one scalar: an array indexed by li & 1 would live in local memory)
    const bool save_a2 = MODE == MODE_FULL_TRAIN && P.a2_img != nullptr;
    // register prefetch of the next item's transform (thread 0) and point (thread f owns point f):
    // the global-load latency is paid behind the current item's work instead of in front of a barrier
    float pf_c0 = 0.f, pf_c1 = 0.f, pf_c2 = 0.f, pf_ang = 0.f, pf_p0 = 0.f, pf_p1 = 0.f, pf_p2 = 0.f;
    auto prefetch = [&](int li) {
      const Item I = item_of(P, it_begin + li);
      const int64_t row0 = (int64_t)I.cloud * P.N + I.p0;
      if (f == 0) {
        pf_c0 = P.center[I.cloud * 3]; pf_c1 = P.center[I.cloud * 3 + 1]; pf_c2 = P.center[I.cloud * 3 + 2];
        pf_ang = P.angle ? P.angle[I.cloud] : 0.f;
      }
      if (f < I.nvalid) {
        const float* src = P.pcs + (row0 + f) * 3;
        pf_p0 = src[0]; pf_p1 = src[1]; pf_p2 = src[2];
      }
    };
    if (n_local > 0) prefetch(0);
    for (int li = 0; li < n_local; ++li) {
      const int it = it_begin + li;
      const Item I = item_of(P, it);
      const int nvalid = I.nvalid, NT = I.NT;
      const int b = li & 1;
      if (f == 0) FTL(0, li);
      if (li >= 2) { mbar_wait_sleep(&bars->a2_empty[b], (ph_a2e >> b) & 1u); ph_a2e ^= 1u << b; }   // (a whole item of slack)
      if (f == 0) FTL(1, li);
      if (save_a2 && f == 0) bulk_wait_read_but1();  // the bulk store of item li-2 no longer reads this buffer
      // A1 aliases the A2 buffer this item will fill after its layer-2 MMA has consumed A1
      uint8_t* sA1 = a2buf(b);
      if (f == 0) {
        float sn = 0.f, cs = 1.f;
        if (P.angle) sincosf(pf_ang, &sn, &cs);
        float* xf = bars->xf + (li & 1) * 8;
        xf[0] = pf_c0; xf[1] = pf_c1; xf[2] = pf_c2; xf[3] = cs; xf[4] = sn;
      }
      const float my_x = pf_p0, my_y = pf_p1, my_z = pf_p2;     // thread f keeps its own point (prefetch() below reloads pf_*)
      asm volatile("bar.sync 1, 256;" ::: "memory");            // the transform thread 0 just wrote is visible to all
      if (f == 0) FTL(2, li);
      const float* xfr = bars->xf + (li & 1) * 8;
      const float cx = xfr[0], cy = xfr[1], cz = xfr[2], cs = xfr[3], sn = xfr[4];
      if (li + 1 < n_local) prefetch(li + 1);
      // ---- layer 1: y = relu(W1f^T p' + c1f), thread f computes all 64 channels of point f (consecutive lanes write
      // consecutive 16-byte rows of each plane: conflict-free)
      if (f < NT) layer1_point(my_x, my_y, my_z, f < nvalid, cx, cy, cz, cs, sn, sW1f, sC1f, sA1 + f * 16, plane1);
      if (f == 0) FTL(3, li);
      fence_proxy_async_smem();
      mbar_arrive(&bars->a1_full);
      // ---- layer-2 epilogue: this thread owns the points  t*128 + 32*quarter + lane  and 64 channels
      mbar_wait_relaxed(&bars->d2_full, ph_d2); ph_d2 ^= 1;
      if (f == 0) FTL(4, li);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int prow = t * 128 + quarter * 32;     // warp-uniform
        if (prow >= NT) break;
        const int p = prow + lane;
        const bool real = p < nvalid;
        uint8_t* dst = a2buf(b) + (size_t)(fgroup * 8) * plane2 + p * 16;
        const uint32_t tb = tmem + lane_base + kTmemD2 + (uint32_t)t * 128u + (uint32_t)fgroup * 64u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld32(tb + h * 32, r);
          tmem_ld_wait();
          // (padding rows of the tile are zeroed by a branch, not by a select per packed word: only the item's last
          // warp diverges, and then only to eight stores)
          if (real) {
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              uint32_t o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 bn = sBn2[fgroup * 32 + h * 16 + c8 * 4 + j];      // broadcast: channels 2c, 2c+1
                o[j] = pack_bf16x2_relu(fmaf(__uint_as_float(r[c8 * 8 + 2 * j]), bn.x, bn.y),
                                        fmaf(__uint_as_float(r[c8 * 8 + 2 * j + 1]), bn.z, bn.w));
              }
              *reinterpret_cast<uint4*>(dst + (size_t)(h * 4 + c8) * plane2) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          } else if (p < NT) {
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8)
              *reinterpret_cast<uint4*>(dst + (size_t)(h * 4 + c8) * plane2) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
      if (f == 0) FTL(5, li);
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&bars->a2_full[b]);
      if (save_a2) {
        asm volatile("bar.sync 2, 256;" ::: "memory");      // every front-end thread has written + fenced
        if (f == 0) bulk_copy_s2g(reinterpret_cast<uint8_t*>(P.a2_img) + (size_t)it * a2_bytes, a2buf(b), a2_bytes);
      }
    }
    if (save_a2 && f == 0) bulk_wait_read_all();
  } else if (warp < 4 || (warp >= 10 && warp < 14)) {
    // ================================ back-end =================================
    // kBackWarpsPerQuarter warps per TMEM lane quarter (warps with equal w%4): the 16-column groups of every accumulator
    // half are dealt round-robin to them -- the MMA thread can only refill a half once it is drained.
    const int bgroup = warp < 4 ? 0 : 1;
    const int e = (warp & 3) * 32 + lane;          // channel within the 128-channel chunk
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ph_full = 0;
    const int C3 = P.nchunk * 128;
    const uint32_t not15 = P.not15;
    const uint32_t keep_mask = ~(P.idx_mask & ~15u);     // clears the group-index bits of the winning element
    for (int li = 0; li < n_local; ++li) {
      const Item I = item_of(P, it_begin + li);
      const int nvalid = I.nvalid, NT = I.NT, p0 = I.p0;
      int N0 = ((NT >> 1) + 15) & ~15;
      if (N0 > NT) N0 = NT;
#pragma unroll 1     // (unrolled eight times the kernel was 226 KB of SASS: instruction-cache misses on every role switch)
      for (int j = 0; j < P.nchunk; ++j) {
        float m = -INFINITY;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int nh = h ? NT - N0 : N0;
          if (nh == 0) continue;
#ifdef AN3D_FWD_ACCWAIT_SLEEP
          mbar_wait_sleep(&bars->acc_full[h], (ph_full >> h) & 1u, AN3D_FWD_ACCWAIT_SLEEP); ph_full ^= 1u << h;
#else
          mbar_wait_relaxed(&bars->acc_full[h], (ph_full >> h) & 1u); ph_full ^= 1u << h;
#endif
          tc_fence_after();
          const int off = h ? N0 : 0;
          const uint32_t tbase = tmem + lane_base + (h ? kTmemAcc1 : kTmemAcc0);
          // software pipeline: the load of the next group is in flight while this one is reduced
          constexpr int kStride = 16 * kBackWarpsPerQuarter;
          uint32_t ra[16], rb[16];
          int g16 = bgroup * 16;
          if (g16 < nh) tmem_ld16(tbase + g16, ra);
          for (; g16 < nh; g16 += 2 * kStride) {
            tmem_ld_wait();
            const int g2 = g16 + kStride;
            if (g2 < nh) tmem_ld16(tbase + g2, rb);
            reduce_group<MODE>(ra, off + g16, nvalid, p0, keep_mask, not15, m);
            if (g2 < nh) {
              tmem_ld_wait();
              if (g2 + kStride < nh) tmem_ld16(tbase + g2 + kStride, ra);
              reduce_group<MODE>(rb, off + g2, nvalid, p0, keep_mask, not15, m);
            }
          }
          tc_fence_before();
          mbar_arrive(&bars->acc_empty[h]);
        }
        uint32_t* zp = P.zext + (size_t)I.cloud * C3 + j * 128 + e;
        // the warps of a lane quarter hold partial maxima of the same channel: always combine atomically
        atomicMax(zp, to_ordered(__float_as_uint(m)));   // zext pre-zeroed
      }
    }
  } else if (warp == 8) {
    // ================================ MMA issuer (whole warp, one elected lane issues) ===============
    if (n_local > 0) {
      // phase bits live in scalar registers (dynamically indexed arrays would go to local memory); every group
      // of MMAs and its commits is issued from one elected region with descriptors derived by constant increments
      uint32_t ph_a1 = 0, ph_a2f = 0, ph_w3f = 0, ph_acce0 = 1, ph_acce1 = 1;
      int stage = 0;
      mbar_wait(&bars->w2_full, 0);
      const uint64_t w2_desc = make_desc(smem_u32(sW2), kPlaneW2, 128);
      // layer 2, point-major: D2[tile t][pt, ch] = A1[128 pts of tile t][64] * W2^T[128 ch][64]^T.  Tile 1 exists when
      // the item has more than 128 points; its MMA reads 128 rows whatever NT is (rows beyond the item are other
      // planes' bytes: finite or not, they only reach accumulator lanes nobody reads).
      auto issue_l2 = [&](int li) {
        const Item I = item_of(P, it_begin + li);
        mbar_wait(&bars->a1_full, ph_a1); ph_a1 ^= 1;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t idesc = make_idesc(128, 128, 0, 0);
          const uint64_t ad = make_desc(smem_u32(a2buf(li & 1)), plane1, 128);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_bf16_raw(tmem + kTmemD2, desc_advance(ad, ks * 2 * plane1), desc_advance(w2_desc, ks * 2 * kPlaneW2), idesc,
                         ks > 0);
          if (I.NT > 128) {
            const uint64_t ad1 = desc_advance(ad, 128 * 16);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_bf16_raw(tmem + kTmemD2 + 128, desc_advance(ad1, ks * 2 * plane1), desc_advance(w2_desc, ks * 2 * kPlaneW2),
                           idesc, ks > 0);
          }
          mma_commit_raw(&bars->d2_full);
        }
        __syncwarp();
      };
      issue_l2(0);
      for (int li = 0; li < n_local; ++li) {
        const Item I = item_of(P, it_begin + li);
        const int NT = I.NT;
        int N0 = ((NT >> 1) + 15) & ~15;
        if (N0 > NT) N0 = NT;
        const int N1 = NT - N0;
        const int b = li & 1;
        mbar_wait(&bars->a2_full[b], (ph_a2f >> b) & 1u); ph_a2f ^= 1u << b;
        tc_fence_after();
        const uint64_t b0_desc = make_desc(smem_u32(a2buf(b)), plane2, 128);
        const uint64_t b1_desc = desc_advance(b0_desc, N0 * 16);
        const uint32_t idesc0 = make_idesc(128, N0, 0, 0), idesc1 = make_idesc(128, N1 > 0 ? N1 : 16, 0, 0);
        bool l2_pending = li + 1 < n_local;
        for (int j = 0; j < P.nchunk; ++j) {
#ifdef AN3D_TIMELINE
          const int ci = li * P.nchunk + j;
          if (blockIdx.x == 0 && lane == 0 && ci < 256) g_mtl[0][ci] = clock64();
#endif
          mbar_wait(&bars->w3_full[stage], (ph_w3f >> stage) & 1u); ph_w3f ^= 1u << stage;
#ifdef AN3D_TIMELINE
          if (blockIdx.x == 0 && lane == 0 && ci < 256) g_mtl[1][ci] = clock64();
#endif
          const uint64_t a_desc = make_desc(smem_u32(sW3 + (size_t)stage * kW3ChunkBytes), kPlaneW2, 128);
          mbar_wait(&bars->acc_empty[0], ph_acce0); ph_acce0 ^= 1;
#ifdef AN3D_TIMELINE
          if (blockIdx.x == 0 && lane == 0 && ci < 256) g_mtl[2][ci] = clock64();
#endif
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_bf16_raw(tmem + kTmemAcc0, desc_advance(a_desc, ks * 2 * kPlaneW2), desc_advance(b0_desc, ks * 2 * plane2),
                           idesc0, ks > 0);
            mma_commit_raw(&bars->acc_full[0]);
          }
          __syncwarp();
          if (N1 > 0) {
            mbar_wait(&bars->acc_empty[1], ph_acce1); ph_acce1 ^= 1;
#ifdef AN3D_TIMELINE
            if (blockIdx.x == 0 && lane == 0 && ci < 256) g_mtl[3][ci] = clock64();
#endif
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                mma_bf16_raw(tmem + kTmemAcc1, desc_advance(a_desc, ks * 2 * kPlaneW2),
                             desc_advance(b1_desc, ks * 2 * plane2), idesc1, ks > 0);
              mma_commit_raw(&bars->acc_full[1]);
            }
            __syncwarp();
          }
          mma_commit(&bars->w3_empty[stage]);
          if (++stage == P.nstages) stage = 0;
          // layer 2 of the NEXT item as soon as its A1 tile is ready -- probed without blocking after every chunk
          // (the front end starts that tile only when this item's A2 tile is done, so a blocking wait after
          // chunk 0 would stall this item's remaining chunks behind the front end); forced after the last chunk
          if (l2_pending && (j == P.nchunk - 1 || __all_sync(0xffffffffu, mbar_try_wait(&bars->a1_full, ph_a1)))) {
            issue_l2(li + 1);
            l2_pending = false;
          }
        }
        mma_commit(&bars->a2_empty[b]);
      }
    }
  } else {
    // ================================ weight loader ============================
    if (lane == 0 && n_local > 0) {
      mbar_arrive_expect_tx(&bars->w2_full, kW2Bytes);
      bulk_copy_g2s(sW2, P.w2t_img, kW2Bytes, &bars->w2_full);
      uint32_t ph_e = 7u;
      const int total = n_local * P.nchunk;
      for (int q = 0; q < total; ++q) {
        const int stage = q % P.nstages;
        mbar_wait_sleep(&bars->w3_empty[stage], (ph_e >> stage) & 1u, 128u); ph_e ^= 1u << stage;
        mbar_arrive_expect_tx(&bars->w3_full[stage], kW3ChunkBytes);
        bulk_copy_g2s(sW3 + (size_t)stage * kW3ChunkBytes,
                      P.w3t_img + (size_t)(q % P.nchunk) * (kW3ChunkBytes / 2), kW3ChunkBytes,
                      &bars->w3_full[stage]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, kTmemCols);
#ifdef AN3D_TIMELINE
  if (blockIdx.x == 0 && tid == 0) {
    printf("FWDK nchunk=%d n_local=%d entry->prologue_done=%lld prologue_done->first_top=%lld last_l2epi->exit=%lld total=%lld\n", P.nchunk, n_local,
           g_ftl[6][1] - g_ftl[6][0], g_ftl[0][0] - g_ftl[6][1], clock64() - g_ftl[5][min(n_local, 64) - 1], clock64() - g_ftl[6][0]);
    {
      // MMA warp: where each layer-3 chunk's issue time goes (averages over chunks 16.. of the first 256)
      const int nc = min(n_local * P.nchunk, 256);
      long long w3 = 0, e0 = 0, e1 = 0, tot = 0;
      int cnt = 0;
      for (int c = 16; c + 1 < nc; ++c) {
        w3 += g_mtl[1][c] - g_mtl[0][c]; e0 += g_mtl[2][c] - g_mtl[1][c]; e1 += g_mtl[3][c] - g_mtl[2][c];
        tot += g_mtl[0][c + 1] - g_mtl[0][c]; ++cnt;
      }
      if (cnt) printf("FWDM nchunk=%d chunks=%d per chunk: total=%lld wait_w3=%lld wait_acc0=%lld issue0+wait_acc1=%lld rest=%lld\n", P.nchunk, cnt,
                      tot / cnt, w3 / cnt, e0 / cnt, e1 / cnt, (tot - w3 - e0 - e1) / cnt);
    }
    const long long t0 = g_ftl[0][0];
    for (int li = 0; li < min(n_local, 64); ++li)
      printf("FWD nchunk=%d li=%d top=%lld a2_empty=%lld staged=%lld l1_done=%lld d2_full=%lld l2epi_done=%lld\n", P.nchunk, li,
             g_ftl[0][li] - t0, g_ftl[1][li] - t0, g_ftl[2][li] - t0, g_ftl[3][li] - t0, g_ftl[4][li] - t0, g_ftl[5][li] - t0);
  }
#endif
}

// =============================================================================================================
// Layer-2 BATCH statistics without running layer 2.  With z2 = a1 W2 (bias apart),
//   sum_p z2[p,c] = (sum_p a1[p,:]) . w_c        sum_p z2[p,c]^2 = w_c^T (A1^T A1) w_c
// so the pass only computes layer 1 and accumulates the 64 x 64 Gram matrix of its (bf16) activations plus their
// column sums on the tensor cores (contraction over points, MN-major operands straight from the A1 tile, one extra
// 'ones' plane) -- no per-item accumulator read-back at all.  The kernel is latency-bound (load points -> layer 1 ->
// a handful of MMAs per item), so it is built small -- 9 warps, 128 TMEM columns, two 10-plane tiles of shared
// memory -- and several CTAs share an SM.
// =============================================================================================================
constexpr int kStatsThreads = 288;       // warps 0-7 layer 1, warp 8 MMA
constexpr uint32_t kStatsTmemCols = 128;

// The Gram MMA is issued with M = 128 (lane = channel): its A operand spans 16 planes from the tile's start although a
// tile has 10 (channels 80..127 are don't-care accumulator lanes).  The bytes it reads there must merely EXIST inside
// the CTA's allocation -- the second tile plus 6 more planes' worth, which also hold the small arrays.
inline size_t stats2_smem_bytes(int PC) {
  const size_t tail = 1024 + 256;                        // folded layer-1 weights + barriers
  return 20 * (size_t)plane_stride(PC) + std::max(tail, 6 * (size_t)plane_stride(PC));
}

struct StatsBars {
  uint64_t a1_full[2], a1_empty[2], done;
  uint32_t tmem_base;
  float xf[16];
};

static __global__ void __launch_bounds__(kStatsThreads) conv_stats2_kernel(const Params P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t plane1 = plane_stride(P.PC);
  const uint32_t a1_bytes = 10 * plane1;                 // 8 planes of channels + the 'ones' plane + a zero plane
  auto a1buf = [&](int b) { return smem + (size_t)b * a1_bytes; };
  float* sW1f = reinterpret_cast<float*>(smem + 2 * a1_bytes);   // [3][64]
  float* sC1f = sW1f + 192;
  StatsBars* bars = reinterpret_cast<StatsBars*>(sC1f + 64);

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int it_begin = min(P.n_items, (int)blockIdx.x * P.item_begin_stride);
  const int it_end = min(P.n_items, it_begin + P.item_begin_stride);
  const int n_local = it_end - it_begin;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->a1_full[i], 256); mbar_init(&bars->a1_empty[i], 1); }
    mbar_init(&bars->done, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 192; i += kStatsThreads) sW1f[i] = P.w1f[i];
  for (int i = tid; i < 64; i += kStatsThreads) sC1f[i] = P.c1f[i];
  if (warp == 8) tmem_alloc(&bars->tmem_base, kStatsTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < 8) {
    const int f = tid;                              // 0..255: the point this thread prefetches and runs layer 1 for
    float pf_c0 = 0.f, pf_c1 = 0.f, pf_c2 = 0.f, pf_ang = 0.f, pf_p0 = 0.f, pf_p1 = 0.f, pf_p2 = 0.f;
    auto prefetch = [&](int li) {
      const Item I = item_of(P, it_begin + li);
      const int64_t row0 = (int64_t)I.cloud * P.N + I.p0;
      if (f == 0) {
        pf_c0 = P.center[I.cloud * 3]; pf_c1 = P.center[I.cloud * 3 + 1]; pf_c2 = P.center[I.cloud * 3 + 2];
        pf_ang = P.angle ? P.angle[I.cloud] : 0.f;
      }
      if (f < I.nvalid) {
        const float* src = P.pcs + (row0 + f) * 3;
        pf_p0 = src[0]; pf_p1 = src[1]; pf_p2 = src[2];
      }
    };
    uint32_t ph_e = 0;
    if (n_local > 0) prefetch(0);
    for (int li = 0; li < n_local; ++li) {
      const Item I = item_of(P, it_begin + li);
      const int nvalid = I.nvalid, NT = I.NT;
      const int b = li & 1;
      if (li >= 2) { mbar_wait_sleep(&bars->a1_empty[b], (ph_e >> b) & 1u, 128u); ph_e ^= 1u << b; }
      if (f == 0) {
        float sn = 0.f, cs = 1.f;
        if (P.angle) sincosf(pf_ang, &sn, &cs);
        float* xf = bars->xf + b * 8;
        xf[0] = pf_c0; xf[1] = pf_c1; xf[2] = pf_c2; xf[3] = cs; xf[4] = sn;
      }
      const float my_x = pf_p0, my_y = pf_p1, my_z = pf_p2;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float* xfr = bars->xf + b * 8;
      const float cx = xfr[0], cy = xfr[1], cz = xfr[2], cs = xfr[3], sn = xfr[4];
      if (li + 1 < n_local) prefetch(li + 1);
      if (f < NT) {
        // thread f owns point f: 64 channels, then its entries of the 'ones' plane (channel 64 = 1 for real points: the column
        // sums) and of the zero plane (channels 72..79)
        uint8_t* row = a1buf(b) + f * 16;
        layer1_point(my_x, my_y, my_z, f < nvalid, cx, cy, cz, cs, sn, sW1f, sC1f, row, plane1);
        *reinterpret_cast<uint4*>(row + 8 * (size_t)plane1) = make_uint4(f < nvalid ? 0x00003f80u : 0u, 0, 0, 0);
        *reinterpret_cast<uint4*>(row + 9 * (size_t)plane1) = make_uint4(0, 0, 0, 0);
      }
      fence_proxy_async_smem();
      mbar_arrive(&bars->a1_full[b]);
    }
    if (n_local > 0 && warp < 4) {
      // Gram accumulator of the whole item range: lanes = layer-1 channel k (64 of them: warps 0, 1 hold real
      // rows; M = 128 leaves lanes 64..127 as the products of the ones / zero planes), 80 columns
      mbar_wait_relaxed(&bars->done, 0);
      tc_fence_after();
      if (warp < 2) {
        const int k = tid;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int g16 = 0; g16 < 80; g16 += 16) {
          uint32_t r[16];
          tmem_ld16(tmem + lane_base + g16, r);
          tmem_ld_wait();
          // this CTA's partial sums go to its own slot; a second kernel adds the slots in CTA order (bit-reproducible)
          float* dst = P.gram1 + (size_t)blockIdx.x * (64 * 80) + k * 80 + g16;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                              __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      }
      tc_fence_before();
    }
  } else if (n_local > 0) {
    const uint32_t idesc = make_idesc(128, 80, 1, 1);
    uint32_t ph = 0;
    for (int li = 0; li < n_local; ++li) {
      const Item I = item_of(P, it_begin + li);
      const int b = li & 1;
      mbar_wait(&bars->a1_full[b], (ph >> b) & 1u); ph ^= 1u << b;
      tc_fence_after();
      if (elect_one()) {
        // contraction over the item's points: both operands are the A1 tile read MN-major (rows = points)
        const uint64_t d = make_desc(smem_u32(a1buf(b)), 128, plane1);
        for (int ks = 0; ks < I.NT / 16; ++ks)
          mma_bf16_raw(tmem, desc_advance(d, ks * 256), desc_advance(d, ks * 256), idesc, (li > 0 || ks > 0) ? 1u : 0u);
        mma_commit_raw(&bars->a1_empty[b]);
        if (li == n_local - 1) mma_commit_raw(&bars->done);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, kStatsTmemCols);
}

}  // namespace convfwd
}  // namespace an3d
