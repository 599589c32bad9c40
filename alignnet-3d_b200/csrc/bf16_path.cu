// bf16 fast path: host orchestration and the small fold / pack / finalize kernels around the
// fused tcgen05 conv-stack kernel (conv_fwd_bf16.cuh).
#include <algorithm>

#include "bf16_path.cuh"
#include "conv_fwd_bf16.cuh"
#include "conv_bwd_bf16.cuh"

namespace an3d {

namespace {

constexpr int kMaxSmem = 227 * 1024;

// ---------------------------------------------------------------------------------------------
// weight packing: bf16 plane images  (element (row r, k) -> (k/8)*rows*8 + r*8 + k%8)
// ---------------------------------------------------------------------------------------------
// W2 [64][128] (TF [Cin][Cout]) -> W2^T image [128 ch][64 k]
__global__ void pack_w2t_kernel(const float* W2, __nv_bfloat16* img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 64) return;
  const int c = i / 64, k = i % 64;
  img[(k >> 3) * 1024 + c * 8 + (k & 7)] = __float2bfloat16_rn(W2[k * 128 + c]);
}

// W3 [128][C3] -> nchunk images of W3^T [128 ch][128 k], column c multiplied by sign(gamma3[c])
__global__ void pack_w3t_kernel(const float* W3, const float* gamma3, __nv_bfloat16* img, int C3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C3 * 128) return;
  const int c = i / 128, k = i % 128;
  const float sg = gamma3[c] < 0.f ? -1.f : 1.f;
  const int j = c >> 7, r = c & 127;
  img[(size_t)j * 16384 + (k >> 3) * 1024 + r * 8 + (k & 7)] = __float2bfloat16_rn(sg * W3[(size_t)k * C3 + c]);
}

// ---------------------------------------------------------------------------------------------
// layer-1 batch statistics in closed form: first and second moments of the transformed points
// ---------------------------------------------------------------------------------------------
__global__ void moments_kernel(const float* pcs, const float* center, const float* angle, int N, int64_t total,
                               double* mom /*[9]*/) {
  double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / N);
    const float x0 = pcs[i * 3] - center[b * 3], y0 = pcs[i * 3 + 1] - center[b * 3 + 1],
                z = pcs[i * 3 + 2] - center[b * 3 + 2];
    float x = x0, y = y0;
    if (angle) {
      float sn, cs;
      sincosf(angle[b], &sn, &cs);
      x = x0 * cs - y0 * sn;
      y = x0 * sn + y0 * cs;
    }
    s[0] += x; s[1] += y; s[2] += z;
    s[3] += (double)x * x; s[4] += (double)x * y; s[5] += (double)x * z;
    s[6] += (double)y * y; s[7] += (double)y * z; s[8] += (double)z * z;
  }
  __shared__ double sm[8][9];
  for (int q = 0; q < 9; ++q)
    for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0)
    for (int q = 0; q < 9; ++q) sm[w][q] = s[q];
  __syncthreads();
  if (threadIdx.x < 9) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i][threadIdx.x];
    atomicAdd(mom + threadIdx.x, t);
  }
}

// Fixed-shape sums of per-CTA partial results: out[r][c] = sum_p parts[p][r][c].  The cross-CTA reductions of the
// statistics passes are bit-reproducible this way (fp32 atomics would add in arrival order).  Block = 32 elements x 8
// part-lanes: lane l adds the parts p = l, l + 8, ... in ascending order, the eight lane sums are then added in lane
// order -- a fixed tree, whatever the schedule.  `extra` (optional, [rows]) receives column `cols_out` of the parts,
// summed the same way in fp64.
constexpr int kSumLanes = 8;
__global__ void __launch_bounds__(256) sum_parts_kernel(const float* parts, int nparts, int rows, int cols_in, int cols_out,
                                                        float* out, double* extra) {
  __shared__ double sm[kSumLanes][32];
  const int e = blockIdx.x * 32 + (threadIdx.x & 31), l = threadIdx.x >> 5;
  const int n_main = rows * cols_out, n_all = n_main + (extra ? rows : 0);
  double acc = 0.0;
  if (e < n_all) {
    const bool is_extra = e >= n_main;
    const int r = is_extra ? e - n_main : e / cols_out, c = is_extra ? cols_out : e - r * cols_out;
    const float* src = parts + (size_t)r * cols_in + c;
    const size_t stride = (size_t)rows * cols_in;
    if (is_extra) {
#pragma unroll 4
      for (int p = l; p < nparts; p += kSumLanes) acc += (double)src[p * stride];
    } else {
      float a = 0.f;
#pragma unroll 4
      for (int p = l; p < nparts; p += kSumLanes) a += src[p * stride];     // (independent loads: unrolling overlaps them)
      acc = (double)a;
    }
  }
  sm[l][threadIdx.x & 31] = acc;
  __syncthreads();
  if (l == 0 && e < n_all) {
    if (e >= n_main) {
      double t = 0.0;
      for (int i = 0; i < kSumLanes; ++i) t += sm[i][threadIdx.x];
      extra[e - n_main] = t;
    } else {
      float t = 0.f;
      for (int i = 0; i < kSumLanes; ++i) t += (float)sm[i][threadIdx.x];
      out[e] = t;
    }
  }
}

struct BnIo {
  const float *gamma, *beta;
  float *state_mean, *state_var;
  float *scale, *shift, *mean, *inv;
};

__device__ __forceinline__ void bn_fold(const BnIo& io, int c, float mu, float var, int training, float decay,
                                        float* sc_out, float* sh_out) {
  if (training) {
    const float om = 1.f - decay;
    io.state_mean[c] = io.state_mean[c] - om * (io.state_mean[c] - mu);
    io.state_var[c] = io.state_var[c] - om * (io.state_var[c] - var);
  } else {
    mu = io.state_mean[c];
    var = io.state_var[c];
  }
  const float rs = 1.0f / sqrtf(var + kBnEps);
  const float sc = io.gamma[c] * rs;
  io.mean[c] = mu;
  io.inv[c] = rs;
  io.scale[c] = sc;
  io.shift[c] = io.beta[c] - mu * sc;
  *sc_out = sc;
  *sh_out = io.beta[c] - mu * sc;
}

// layer 1: z = W1^T p + b1 ; mean_z = W1^T mean_p + b1 ; var_z = w^T Cov w  (exact, fp64)
__global__ void fold_l1_kernel(const double* mom, double count, const float* W1 /*[3][64]*/, const float* b1, BnIo io,
                               int training, float decay, float* w1f, float* c1f) {
  const int c = threadIdx.x;
  if (c >= 64) return;
  float mu = 0.f, var = 1.f;
  if (training) {
    const double mx = mom[0] / count, my = mom[1] / count, mz = mom[2] / count;
    const double cxx = mom[3] / count - mx * mx, cxy = mom[4] / count - mx * my, cxz = mom[5] / count - mx * mz,
                 cyy = mom[6] / count - my * my, cyz = mom[7] / count - my * mz, czz = mom[8] / count - mz * mz;
    const double wx = W1[c], wy = W1[64 + c], wz = W1[128 + c];
    mu = (float)(wx * mx + wy * my + wz * mz + (double)b1[c]);
    const double v = wx * wx * cxx + wy * wy * cyy + wz * wz * czz + 2.0 * (wx * wy * cxy + wx * wz * cxz + wy * wz * cyz);
    var = (float)fmax(v, 0.0);
  }
  float sc, sh;
  bn_fold(io, c, mu, var, training, decay, &sc, &sh);
  w1f[c] = sc * W1[c];
  w1f[64 + c] = sc * W1[64 + c];
  w1f[128 + c] = sc * W1[128 + c];
  c1f[c] = sc * b1[c] + sh;
}

// layers 2/3: statistics of the raw accumulator (bias excluded; layer 3 sign-folded by sign(gamma))
__global__ void fold_acc_kernel(const double* stats, double count, const float* bias, BnIo io, int C, int training,
                                float decay, int sign_folded, float* shift_folded) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mu = 0.f, var = 1.f;
  if (training) {
    const double sg = (sign_folded && io.gamma[c] < 0.f) ? -1.0 : 1.0;
    const double m = stats[2 * c] / count;
    const double v = stats[2 * c + 1] / count - m * m;
    mu = (float)(sg * m + (double)bias[c]);
    var = (float)fmax(v, 0.0);
  }
  float sc, sh;
  bn_fold(io, c, mu, var, training, decay, &sc, &sh);
  if (shift_folded) shift_folded[c] = sc * bias[c] + sh;
}

// Layer-2 batch statistics from the Gram matrix of the layer-1 activations (MODE_STATS2 of the conv kernel):
//   stats2[c] = ( sum_k sa1[k] w[k,c] ,  sum_{k,k'} w[k,c] G1[k,k'] w[k',c] ),  w = bf16-rounded W2 [64][128].
// Grid 8 x 128 threads = 16 channels x 8 row groups per block.
__global__ void __launch_bounds__(128) stats2_from_gram1_kernel(const float* W2, const float* gram1, double* stats2) {
  __shared__ float sg[64][81];
  __shared__ float sw[64][17];
  __shared__ double red[2][8][16];
  const int c0 = blockIdx.x * 16;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (int i = threadIdx.x; i < 64 * 80; i += 128) sg[i / 80][i % 80] = gram1[i];
  for (int i = threadIdx.x; i < 64 * 16; i += 128) {
    const int k = i >> 4, cc = i & 15;
    sw[k][cc] = __bfloat162float(__float2bfloat16_rn(W2[k * 128 + c0 + cc]));
  }
  __syncthreads();
  double m = 0.0, qd = 0.0;
  for (int k = ty * 8; k < ty * 8 + 8; ++k) {
    float t = 0.f;
#pragma unroll 8
    for (int kp = 0; kp < 64; ++kp) t = fmaf(sg[k][kp], sw[kp][tx], t);
    qd += (double)sw[k][tx] * (double)t;
    m += (double)sg[k][64] * (double)sw[k][tx];
  }
  red[0][ty][tx] = m;
  red[1][ty][tx] = qd;
  __syncthreads();
  if (ty == 0) {
    for (int i = 1; i < 8; ++i) { m += red[0][i][tx]; qd += red[1][i][tx]; }
    stats2[2 * (c0 + tx)] = m;
    stats2[2 * (c0 + tx) + 1] = qd;
  }
}

// Layer-3 batch statistics without touching the [M, C3] accumulator: with r3 = a2 W3b,
//   sum_m r3[m,c]   = sa2 . w_c              sum_m r3[m,c]^2 = w_c^T (A2^T A2) w_c
// (bf16-rounded weights, Gram matrix from the tensor-core pass).  The product GW = (A2^T A2) W3b is kept
// for the backward pass (dense part of wgrad3).  CTA = 32 channels, 512 threads = 32 channels x 16 row
// groups of 8 rows; Gram matrix and weight tile staged in shared memory.
constexpr int kGwThreads = 512;
constexpr size_t kGwSmem = (128 * 128 + 128 * 33) * sizeof(float) + 2 * 16 * 32 * sizeof(double);
__global__ void __launch_bounds__(kGwThreads) gw3_stats_kernel(const float* W3, const float* gram, const double* sa2, int C3,
                                                               float* gw, double* stats3) {
  extern __shared__ __align__(16) uint8_t gw_smem[];
  float* sg = reinterpret_cast<float*>(gw_smem);          // [128][128]
  float* sw = sg + 128 * 128;                              // [128][33]
  double* red = reinterpret_cast<double*>(sw + 128 * 33);  // [2][16][32]
  const int c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 128 * 128 / 4; i += kGwThreads)
    reinterpret_cast<float4*>(sg)[i] = reinterpret_cast<const float4*>(gram)[i];
  for (int i = threadIdx.x; i < 128 * 32; i += kGwThreads) {
    const int kp = i >> 5, cc = i & 31;
    sw[kp * 33 + cc] = __bfloat162float(__float2bfloat16_rn(W3[(size_t)kp * C3 + c0 + cc]));
  }
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.f;
  for (int kp = 0; kp < 128; kp += 4) {
    const float w0 = sw[kp * 33 + tx], w1 = sw[(kp + 1) * 33 + tx], w2 = sw[(kp + 2) * 33 + tx], w3 = sw[(kp + 3) * 33 + tx];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float4 g = *reinterpret_cast<const float4*>(sg + (ty + 16 * r) * 128 + kp);
      acc[r] = fmaf(g.x, w0, fmaf(g.y, w1, fmaf(g.z, w2, fmaf(g.w, w3, acc[r]))));
    }
  }
  double m = 0.0, qd = 0.0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int k = ty + 16 * r;
    const float w = sw[k * 33 + tx];
    if (gw) gw[(size_t)k * C3 + c0 + tx] = acc[r];
    qd += (double)w * (double)acc[r];
    m += sa2[k] * (double)w;
  }
  red[ty * 32 + tx] = m;
  red[512 + ty * 32 + tx] = qd;
  __syncthreads();
  if (ty == 0) {
    for (int i = 1; i < 16; ++i) { m += red[i * 32 + tx]; qd += red[512 + i * 32 + tx]; }
    stats3[2 * (c0 + tx)] = m;
    stats3[2 * (c0 + tx) + 1] = qd;
  }
}

// pooled feature from the packed extreme of the sign-folded raw accumulator:
//   z_ext = sign(gamma) * unpack(key) + b3 ;  g = relu(scale * z_ext + shift)   (BN + ReLU are monotone)
__global__ void pool_finalize_kernel(const uint32_t* zext, int B, int C3, const float* gamma, const float* bias,
                                     const float* scale, const float* shift, uint32_t idx_mask, float* G, int64_t ldg,
                                     int32_t* gidx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * C3) return;
  const int b = (int)(i / C3), c = (int)(i % C3);
  const uint32_t key = zext[i];
  const uint32_t bits = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
  const float v = __uint_as_float(bits & ~idx_mask);
  const float sg = gamma[c] < 0.f ? -1.f : 1.f;
  const float z = sg * v + bias[c];
  G[(int64_t)b * ldg + c] = fmaxf(fmaf(z, scale[c], shift[c]), 0.f);
  if (gidx) gidx[i] = (int32_t)(bits & idx_mask);
}

BnIo bn_io(const Model& m, const PlanF32& p, const float* params, float* state, int br, int bn) {
  BnIo io;
  const int ch = m.bn_branch[bn].ch;
  const int64_t po = m.bn_param_off(false, br, bn), so = m.bn_state_off(false, br, bn), sl = m.bn_slot_off(false, br, bn);
  io.gamma = params + po;
  io.beta = params + po + ch;
  io.state_mean = state + so;
  io.state_var = state + so + ch;
  io.scale = p.bn.scale + sl;
  io.shift = p.bn.shift + sl;
  io.mean = p.bn.mean + sl;
  io.inv = p.bn.inv + sl;
  return io;
}

template <int MODE>
int launch_fused(const convfwd::Params& P, int grid, size_t smem, cudaStream_t st) {
  AN3D_CUDA_CHECK(cudaFuncSetAttribute(convfwd::conv_stack_fwd_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
  prof_mark(PROF_CONV_FULL, true, st);
  convfwd::conv_stack_fwd_kernel<MODE><<<grid, convfwd::kThreads, smem, st>>>(P);
  prof_mark(PROF_CONV_FULL, false, st);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

// layer-2 batch statistics pass: a light, latency-bound kernel -- as many CTAs per SM as fit
int launch_stats2(convfwd::Params P, int sms, float* gram1_out, cudaStream_t st) {
  const size_t smem = convfwd::stats2_smem_bytes(P.PC);
  AN3D_CUDA_CHECK(cudaFuncSetAttribute(convfwd::conv_stats2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // (without the carve-out preference the driver sizes shared memory for ONE block of this kernel per SM: the first
  // version of this launch ran 148 CTAs although two fit)
  AN3D_CUDA_CHECK(cudaFuncSetAttribute(convfwd::conv_stats2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
  // CTAs per SM from the kernel's own footprint (shared memory + 1 KB reserved per block, 288 threads, 72 registers,
  // 128 TMEM columns).  cudaOccupancyMaxActiveBlocksPerMultiprocessor answered 1 for this kernel on B200 although two
  // blocks fit (86.9 KB each at 208 points per item), so the launch does its own arithmetic; if the hardware disagrees
  // the surplus CTAs simply queue.
  const int by_smem = (int)((228 * 1024) / (smem + 1024)), by_regs = 65536 / (convfwd::kStatsThreads * 72);
  const int per_sm = std::max(1, std::min(std::min(by_smem, by_regs), (int)(512 / convfwd::kStatsTmemCols)));
  const int grid = std::max(1, std::min(std::min(P.n_items, sms * per_sm), kMaxParts1));
  P.item_begin_stride = (P.n_items + grid - 1) / grid;
  const int nparts = (P.n_items + P.item_begin_stride - 1) / P.item_begin_stride;   // CTAs that own items (and write a slot)
  prof_mark(PROF_CONV_STATS2, true, st);
  convfwd::conv_stats2_kernel<<<grid, convfwd::kStatsThreads, smem, st>>>(P);
  prof_mark(PROF_CONV_STATS2, false, st);
  AN3D_LAUNCH_CHECK();
  sum_parts_kernel<<<(64 * 80 + 31) / 32, 256, 0, st>>>(P.gram1, nparts, 64, 80, 80, gram1_out, nullptr);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

}  // namespace

int bf16_supported(const Model& m) {
  for (int s = 0; s < 3; ++s) {
    if (m.conv[s].size() != 3 || m.conv[s][0].cout != 64 || m.conv[s][1].cout != 128 || m.conv[s][2].cout % 128 != 0 ||
        m.conv[s][2].cout > 1024) {
      set_error("bf16 path implements conv stacks of the form [64, 128, C] with C a multiple of 128 (<= 1024); stage %d "
                "differs -- use AN3D_PRECISION_FP32 for this architecture", s);
      return AN3D_ERR_UNSUPPORTED;
    }
  }
  return AN3D_OK;
}

void plan_bf16(const Model& m, int B, int N, int flags, Arena& a, PlanBf16* q) {
  const bool training = (flags & AN3D_TRAINING) != 0;
  const int64_t M = (int64_t)B * N;
  // training tiles are capped at 208 points so that the backward kernels (two tile images, two
  // scatter tiles and a weight ring in shared memory) fit; inference uses up to 256
  const int maxpc = training ? 208 : convfwd::kMaxPC;
  q->npc = (N + maxpc - 1) / maxpc;
  q->PC = (((N + q->npc - 1) / q->npc) + 15) & ~15;
  int bits = 0;
  while ((1 << bits) < N) ++bits;
  q->idx_mask = (1u << bits) - 1u;
  q->img_bytes = 16 * (int64_t)convfwd::plane_stride(q->PC);
  for (int s = 0; s < 3; ++s) {
    const int C3 = m.conv[s].back().cout;
    q->w2t[s] = a.take<__nv_bfloat16>(128 * 64);
    for (int br = 0; br < 2; ++br) {
      q->w3t[s][br] = a.take<__nv_bfloat16>((int64_t)C3 * 128);
      q->w1f[s][br] = a.take<float>(192);
      q->c1f[s][br] = a.take<float>(64);
      q->t2f[s][br] = a.take<float>(128);
      q->moments[s][br] = a.take<double>(16);
      q->stats2[s][br] = a.take<double>(256);
      q->gram1[s][br] = a.take<float>(64 * 80);
      q->stats3[s][br] = a.take<double>(2 * (int64_t)C3);
      q->zext[s][br] = a.take<uint32_t>((int64_t)B * C3);
      q->a2img[s][br] = training ? a.take<__nv_bfloat16>((int64_t)B * q->npc * (q->img_bytes / 2)) : nullptr;
      q->sa2[s][br] = a.take<double>(128);
      q->gram[s][br] = a.take<float>(128 * 128);
      q->gw[s][br] = training ? a.take<float>(128 * (int64_t)C3) : nullptr;
    }
  }
  if (training) {
    for (int br = 0; br < 2; ++br) {
      q->gram1_parts[br] = a.take<float>((int64_t)kMaxParts1 * 64 * 80);
      q->gram_parts[br] = a.take<float>((int64_t)kMaxParts * 128 * 132);
    }
    int64_t c3max = 0;
    for (int s = 0; s < 3; ++s) {
      const int C3 = m.conv[s].back().cout;
      c3max = std::max<int64_t>(c3max, C3);
      q->w3n[s] = a.take<__nv_bfloat16>((int64_t)C3 * 128);
      q->w2p[s] = a.take<__nv_bfloat16>(128 * 128);
    }
    for (int br = 0; br < 2; ++br) {
      BwdScratch& w = q->bw[br];
      w.dyext = a.take<float>((int64_t)B * c3max);
      w.red3 = a.take<double>(2 * c3max);
      w.coef3 = a.take<float>(4 * c3max);
      w.gq = a.take<__nv_bfloat16>(128 * 128);
      w.gq_f32 = a.take<float>(128 * 128);
      w.uvec = a.take<float>(128);
      w.dy2img = a.take<__nv_bfloat16>((int64_t)B * q->npc * (q->img_bytes / 2));
      w.red2 = a.take<double>(256);
      w.coef2 = a.take<float>(256);
      w.l1sums = a.take<float>((int64_t)B * q->npc * 256);
      w.red1 = a.take<double>(128);
      w.coef1 = a.take<float>(128);
    }
  }
}

int pack_weights_bf16(const Model& m, const PlanF32& p, const float* params, cudaStream_t st) {
  for (int s = 0; s < 3; ++s) {
    const int C3 = m.conv[s].back().cout;
    pack_w2t_kernel<<<(128 * 64 + 255) / 256, 256, 0, st>>>(params + m.conv[s][1].w, p.bf.w2t[s]);
    AN3D_LAUNCH_CHECK();
    for (int br = 0; br < 2; ++br) {
      const float* gamma3 = params + m.bn_param_off(false, br, m.conv[s][2].bn);
      pack_w3t_kernel<<<(C3 * 128 + 255) / 256, 256, 0, st>>>(params + m.conv[s][2].w, gamma3, p.bf.w3t[s][br], C3);
      AN3D_LAUNCH_CHECK();
    }
  }
  return AN3D_OK;
}

int conv_stack_forward_bf16(const Model& m, const PlanF32& p, int s, int br, const float* pcs, const float* center,
                            const float* angle, const float* params, float* state, bool training, float decay,
                            cudaStream_t st) {
  const PlanBf16& q = p.bf;
  const int B = p.B, N = p.N;
  const int64_t M = p.M;
  const Lin &L1 = m.conv[s][0], &L2 = m.conv[s][1], &L3 = m.conv[s][2];
  const int C3 = L3.cout;
  int dev = 0, sms = 148;
  AN3D_CUDA_CHECK(cudaGetDevice(&dev));
  AN3D_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

  BnIo io1 = bn_io(m, p, params, state, br, L1.bn), io2 = bn_io(m, p, params, state, br, L2.bn),
       io3 = bn_io(m, p, params, state, br, L3.bn);

  convfwd::Params P;
  P.pcs = pcs; P.center = center; P.angle = angle; P.B = B; P.N = N; P.PC = q.PC; P.npc = q.npc;
  P.n_items = B * q.npc;
  const int grid = std::min(P.n_items, sms);
  P.item_begin_stride = (P.n_items + grid - 1) / grid;
  P.w1f = q.w1f[s][br]; P.c1f = q.c1f[s][br]; P.w2t_img = q.w2t[s]; P.s2 = io2.scale; P.t2f = q.t2f[s][br];
  P.w3t_img = q.w3t[s][br]; P.nchunk = C3 / 128;
  P.nstages = convfwd::smem_bytes(q.PC, 3) <= (size_t)kMaxSmem ? 3 : 2;
  P.zext = q.zext[s][br]; P.gram1 = q.gram1_parts[br];
  P.a2_img = training ? q.a2img[s][br] : nullptr;
  P.idx_mask = q.idx_mask; P.not15 = ~15u;
  const size_t smem = convfwd::smem_bytes(q.PC, P.nstages);
  if (smem > (size_t)kMaxSmem) {
    set_error("conv_stack_forward_bf16: tile needs %zu bytes of shared memory", smem);
    return AN3D_ERR_UNSUPPORTED;
  }

  if (training) {
    AN3D_CUDA_CHECK(cudaMemsetAsync(q.moments[s][br], 0, 16 * sizeof(double), st));
    const int mb = (int)std::min<int64_t>((M + 255) / 256, 4 * sms);
    moments_kernel<<<mb, 256, 0, st>>>(pcs, center, angle, N, M, q.moments[s][br]);
    AN3D_LAUNCH_CHECK();
  }
  const bool prepared = p.prepared && !training;   // inference with unchanged parameters: the folds below are still valid
  if (!prepared) {
    fold_l1_kernel<<<1, 64, 0, st>>>(q.moments[s][br], (double)M, params + L1.w, params + L1.b, io1, training ? 1 : 0, decay,
                                     q.w1f[s][br], q.c1f[s][br]);
    AN3D_LAUNCH_CHECK();
  }
  if (training) {
    AN3D_TRY(launch_stats2(P, sms, q.gram1[s][br], st));
    stats2_from_gram1_kernel<<<8, 128, 0, st>>>(params + L2.w, q.gram1[s][br], q.stats2[s][br]);
    AN3D_LAUNCH_CHECK();
  }
  if (!prepared) {
    fold_acc_kernel<<<1, 128, 0, st>>>(q.stats2[s][br], (double)M, params + L2.b, io2, 128, training ? 1 : 0, decay, 0,
                                       q.t2f[s][br]);
    AN3D_LAUNCH_CHECK();
  }
  AN3D_CUDA_CHECK(cudaMemsetAsync(q.zext[s][br], 0, (size_t)B * C3 * sizeof(uint32_t), st));
  if (training) AN3D_TRY(launch_fused<convfwd::MODE_FULL_TRAIN>(P, grid, smem, st));
  else AN3D_TRY(launch_fused<convfwd::MODE_FULL_EVAL>(P, grid, smem, st));
  if (training) {
    // Gram matrix of the saved layer-2 activations on the tensor cores -> layer-3 BN statistics
    convbwd::Gram2Params W;
    W.a2_img = reinterpret_cast<const uint8_t*>(q.a2img[s][br]); W.img_bytes = (uint32_t)q.img_bytes;
    W.N = N; W.PC = q.PC; W.npc = q.npc; W.n_items = P.n_items; W.parts = q.gram_parts[br];
    const int nranges = std::max(1, std::min(std::min(P.n_items, sms), kMaxParts));
    W.items_per_cta = (P.n_items + nranges - 1) / nranges;
    const int nparts = (P.n_items + W.items_per_cta - 1) / W.items_per_cta;
    const size_t gsmem = convbwd::gram2_smem_bytes(q.PC);
    AN3D_CUDA_CHECK(cudaFuncSetAttribute(convbwd::gram2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
    prof_mark(PROF_GRAM2, true, st);
    convbwd::gram2_kernel<<<nranges, convbwd::kGram2Threads, gsmem, st>>>(W);
    prof_mark(PROF_GRAM2, false, st);
    AN3D_LAUNCH_CHECK();
    sum_parts_kernel<<<(128 * 128 + 128 + 31) / 32, 256, 0, st>>>(q.gram_parts[br], nparts, 128, convbwd::kGram2PartCols, 128,
                                                                    q.gram[s][br], q.sa2[s][br]);
    AN3D_LAUNCH_CHECK();
    AN3D_CUDA_CHECK(cudaFuncSetAttribute(gw3_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGwSmem));
    gw3_stats_kernel<<<C3 / 32, kGwThreads, kGwSmem, st>>>(params + L3.w, q.gram[s][br], q.sa2[s][br], C3, q.gw[s][br],
                                                           q.stats3[s][br]);
    AN3D_LAUNCH_CHECK();
  }
  if (!prepared) {
    fold_acc_kernel<<<(C3 + 127) / 128, 128, 0, st>>>(q.stats3[s][br], (double)M, params + L3.b, io3, C3, training ? 1 : 0,
                                                      decay, 0, nullptr);
    AN3D_LAUNCH_CHECK();
  }
  const int64_t ldg = s == EMB ? 2 * C3 : C3;
  const int64_t tot = (int64_t)B * C3;
  pool_finalize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(q.zext[s][br], B, C3, io3.gamma, params + L3.b,
                                                                      io3.scale, io3.shift, training ? q.idx_mask : 0u,
                                                                      p.g[s][br], ldg, training ? p.gidx[s][br] : nullptr);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

}  // namespace an3d

