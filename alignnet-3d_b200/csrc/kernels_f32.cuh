// fp32 CUDA-core kernels of the parity path (AN3D_PRECISION_FP32) and the small per-cloud /
// per-sample kernels shared with the bf16 path.  Header-only (included by forward.cu /
// backward.cu) so every kernel is compiled once per translation unit that launches it.
#pragma once
#include "common.cuh"

namespace an3d {

// --------------------------------------------------------------------------------------------
// Generic tiled SGEMM:  C[M,N] (+)= pro(A) * B (+ bias)
//   A(m,k) = TA ? A[k*lda + m] : A[m*lda + k]      B(k,n) = TB ? B[n*ldb + k] : B[k*ldb + n]
//   pro(): optional per-channel affine + ReLU (the BN+ReLU of the producing layer, applied on
//   load so activations are never materialised) and optional dropout mask; the channel index is
//   k (TA = false) or m (TA = true).
// --------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A = nullptr;
  int64_t lda = 0;
  const float* B = nullptr;
  int64_t ldb = 0;
  float* C = nullptr;
  int64_t ldc = 0;
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;
  const float* pro_scale = nullptr;
  const float* pro_shift = nullptr;
  const float* pro_mask = nullptr;
  float pro_mask_scale = 1.f;
  int accumulate = 0;  // atomicAdd into C
  int ksplit = 1;
};

template <bool TA, bool TB>
static __global__ void __launch_bounds__(256) gemm_f32_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][68];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  int kchunk = (g.K + g.ksplit - 1) / g.ksplit;
  kchunk = (kchunk + 15) / 16 * 16;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(g.K, kbeg + kchunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int mm, kk;
      if (TA) { kk = e >> 6; mm = e & 63; } else { mm = e >> 4; kk = e & 15; }
      const int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < g.M && k < kend) {
        const int64_t idx = TA ? (int64_t)k * g.lda + m : (int64_t)m * g.lda + k;
        v = g.A[idx];
        if (g.pro_scale) {
          const int ch = TA ? m : k;
          v = fmaxf(fmaf(v, g.pro_scale[ch], g.pro_shift[ch]), 0.f);
        }
        if (g.pro_mask) v *= g.pro_mask[idx] * g.pro_mask_scale;
      }
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int nn, kk;
      if (TB) { nn = e >> 4; kk = e & 15; } else { kk = e >> 6; nn = e & 63; }
      const int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < g.N && k < kend) v = g.B[TB ? (int64_t)n * g.ldb + k : (int64_t)k * g.ldb + n];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias && blockIdx.z == 0) v += g.bias[n];
      float* dst = g.C + (int64_t)m * g.ldc + n;
      if (g.accumulate) atomicAdd(dst, v); else *dst = v;
    }
  }
}

inline int launch_gemm(const GemmArgs& g, bool ta, bool tb, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return AN3D_OK;
  dim3 grid((g.M + 63) / 64, (g.N + 63) / 64, g.ksplit);
  if (ta && tb) gemm_f32_kernel<true, true><<<grid, 256, 0, st>>>(g);
  else if (ta) gemm_f32_kernel<true, false><<<grid, 256, 0, st>>>(g);
  else if (tb) gemm_f32_kernel<false, true><<<grid, 256, 0, st>>>(g);
  else gemm_f32_kernel<false, false><<<grid, 256, 0, st>>>(g);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

// --------------------------------------------------------------------------------------------
// Column reductions over Z[R,C] (double accumulation, atomics across row chunks)
// --------------------------------------------------------------------------------------------
enum ColMode { COL_SUM = 0, COL_SQDIFF = 1, COL_DY = 2, COL_SUMSQ = 3 };   // SUMSQ: sum z and sum z^2 in one pass (double)

struct ColArgs {
  const float* Z = nullptr;   // [R, C] leading dim ldz
  int64_t ldz = 0;
  int R = 0, C = 0;
  const float* mean = nullptr;   // SQDIFF, DY
  const float* inv = nullptr;    // DY
  const float* scale = nullptr;  // DY (relu mask uses z*scale+shift > 0)
  const float* shift = nullptr;
  const float* dA = nullptr;     // DY: gradient wrt post-activation [R, C] ld = ldd
  int64_t ldd = 0;
  const float* mask = nullptr;   // DY: dropout mask on the activation, same layout as dA
  float mask_scale = 1.f;
  double* acc0 = nullptr;
  double* acc1 = nullptr;
};

template <int MODE>
static __global__ void __launch_bounds__(256) col_reduce_kernel(ColArgs a, int rows_per_block) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(a.R, r0 + rows_per_block);
  double s0 = 0.0, s1 = 0.0;
  if (c < a.C) {
    float mean = 0.f, inv = 0.f, sc = 0.f, sh = 0.f;
    if (MODE == COL_SQDIFF || MODE == COL_DY) mean = a.mean[c];
    if (MODE == COL_DY) { inv = a.inv[c]; sc = a.scale[c]; sh = a.shift[c]; }
    // four rows per iteration: the loads of a batch are independent and in flight together
    for (int rb = r0 + threadIdx.y; rb < r1; rb += 32) {
      float z[4], dy[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = rb + 8 * u;
        z[u] = 0.f; dy[u] = 0.f;
        if (r < r1) {
          z[u] = a.Z[(int64_t)r * a.ldz + c];
          if (MODE == COL_DY) {
            dy[u] = a.dA[(int64_t)r * a.ldd + c];
            if (a.mask) dy[u] *= a.mask[(int64_t)r * a.ldd + c] * a.mask_scale;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (rb + 8 * u >= r1) continue;
        if (MODE == COL_SUM) {
          s0 += (double)z[u];
        } else if (MODE == COL_SUMSQ) {
          s0 += (double)z[u];
          s1 += (double)z[u] * (double)z[u];
        } else if (MODE == COL_SQDIFF) {
          const float d = z[u] - mean;
          s1 += (double)d * (double)d;
        } else {
          const float g = fmaf(z[u], sc, sh) > 0.f ? dy[u] : 0.f;
          s0 += (double)g;
          s1 += (double)g * (double)((z[u] - mean) * inv);
        }
      }
    }
  }
  __shared__ double sm0[8][33], sm1[8][33];
  sm0[threadIdx.y][threadIdx.x] = s0;
  sm1[threadIdx.y][threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.y == 0 && c < a.C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) { s0 += sm0[i][threadIdx.x]; s1 += sm1[i][threadIdx.x]; }
    if (MODE != COL_SQDIFF) atomicAdd(a.acc0 + c, s0);
    if (MODE != COL_SUM) atomicAdd(a.acc1 + c, s1);
  }
}

inline int launch_col_reduce(const ColArgs& a, int mode, cudaStream_t st) {
  if (a.R <= 0 || a.C <= 0) return AN3D_OK;
  const int rows_per_block = 128;
  dim3 grid((a.C + 31) / 32, (a.R + rows_per_block - 1) / rows_per_block);
  dim3 block(32, 8);
  if (mode == COL_SUM) col_reduce_kernel<COL_SUM><<<grid, block, 0, st>>>(a, rows_per_block);
  else if (mode == COL_SQDIFF) col_reduce_kernel<COL_SQDIFF><<<grid, block, 0, st>>>(a, rows_per_block);
  else if (mode == COL_SUMSQ) col_reduce_kernel<COL_SUMSQ><<<grid, block, 0, st>>>(a, rows_per_block);
  else col_reduce_kernel<COL_DY><<<grid, block, 0, st>>>(a, rows_per_block);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

// mean = acc0 / R
static __global__ void bn_mean_kernel(const double* acc0, float* mean, int C, double inv_rows) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) mean[c] = (float)(acc0[c] * inv_rows);
}

// Training: var = acc1/R; scale/shift/inv; EMA shadow update (utils/tf_util.py:475-480).
// Eval (acc1 == nullptr): statistics come from the shadows.
static __global__ void bn_finalize_kernel(const double* acc1, double inv_rows, const float* gamma, const float* beta,
                                   float* state_mean, float* state_var, float* mean, float* inv, float* scale,
                                   float* shift, int C, int training, float decay) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mu, var;
  if (training) {
    mu = mean[c];
    var = (float)(acc1[c] * inv_rows);
    const float om = 1.f - decay;
    state_mean[c] = state_mean[c] - om * (state_mean[c] - mu);
    state_var[c] = state_var[c] - om * (state_var[c] - var);
  } else {
    mu = state_mean[c];
    var = state_var[c];
    mean[c] = mu;
  }
  const float rs = 1.0f / sqrtf(var + kBnEps);
  const float sc = gamma[c] * rs;
  inv[c] = rs;
  scale[c] = sc;
  shift[c] = beta[c] - mu * sc;
}

// --------------------------------------------------------------------------------------------
// Max-pool over the N points of each cloud of relu(z*scale+shift); first max wins.
// --------------------------------------------------------------------------------------------
static __global__ void pool_kernel(const float* Z, int64_t ldz, int N, int C, const float* scale, const float* shift,
                            float* G, int64_t ldg, int32_t* idx) {
  const int b = blockIdx.x;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = scale[c], sh = shift[c];
  float best = -INFINITY;
  int bi = 0;
  const float* z = Z + (int64_t)b * N * ldz + c;
  for (int n = 0; n < N; ++n) {
    const float y = fmaxf(fmaf(z[(int64_t)n * ldz], sc, sh), 0.f);
    if (y > best) { best = y; bi = n; }
  }
  G[(int64_t)b * ldg + c] = best;
  if (idx) idx[(int64_t)b * C + c] = bi;
}

// --------------------------------------------------------------------------------------------
// Point-stage prologues
// --------------------------------------------------------------------------------------------
// centroid of each cloud (models/tp8.py:104): one warp per cloud
static __global__ void centroid_kernel(const float* pcs, int N, float* mu, int B) {
  const int b = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  double s[3] = {0, 0, 0};
  const float* p = pcs + (int64_t)b * N * 3;
  for (int n = lane; n < N; n += 32) { s[0] += p[n * 3]; s[1] += p[n * 3 + 1]; s[2] += p[n * 3 + 2]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    for (int d = 0; d < 3; ++d) s[d] += __shfl_xor_sync(0xffffffffu, s[d], o);
  if (lane == 0)
    for (int d = 0; d < 3; ++d) mu[b * 3 + d] = (float)(s[d] / N);
}

// out = Rz(+a)(p - c)  (models/tp8.py:106,113,122-127); angle == nullptr -> no rotation
static __global__ void stage_input_kernel(const float* pcs, const float* center, const float* angle, float* out, int N,
                                   int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / N);
  const float x = pcs[i * 3] - center[b * 3], y = pcs[i * 3 + 1] - center[b * 3 + 1], z = pcs[i * 3 + 2] - center[b * 3 + 2];
  if (angle) {
    float s, c;
    sincosf(angle[b], &s, &c);
    out[i * 3] = x * c - y * s;
    out[i * 3 + 1] = x * s + y * c;
  } else {
    out[i * 3] = x;
    out[i * 3 + 1] = y;
  }
  out[i * 3 + 2] = z;
}

// --------------------------------------------------------------------------------------------
// Per-sample epilogues of the three MLPs
// --------------------------------------------------------------------------------------------
static __global__ void post_s1_kernel(const float* d1, const float* mu, float* c1, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * 3) c1[i] = d1[i] + mu[i];  // tp8.py:109
}

__device__ __forceinline__ float floor_mod(float x, float y) {  // tf.mod / np.mod for floats [TF-sem]
  float r = fmodf(x, y);
  if (r != 0.f && ((y < 0.f) != (r < 0.f))) r += y;
  return r;
}

// tf_get_angles (tp8.py:294-301 + 202-212) for one row of logits
__device__ __forceinline__ float decode_angle_scaled(const float* lg, int nb, int* kout) {
  int k = 0;
  float best = lg[0];
  for (int j = 1; j < nb; ++j)
    if (lg[j] > best) { best = lg[j]; k = j; }
  const float pi = 3.14159265358979323846f;
  const float res = lg[nb + k] * (pi / (float)nb);
  const float a = (float)k * (2.0f * pi / (float)nb) + res;
  if (kout) *kout = k;
  return floor_mod(a + pi, 2.0f * pi) - pi;
}

// o2 [B,3+2nb] -> c2 = o2[:, :3] + c1 (tp8.py:117), logits copy (:118), decoded yaw (:123).  One warp per sample:
// coalesced copy, arg-max by shuffles with tf.argmax's first-maximum tie rule.
static __global__ void post_s2_kernel(const float* o2, const float* c1, float* c2, float* logits, float* ang, int32_t* angk,
                               int B, int nb) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* o = o2 + (int64_t)b * (3 + 2 * nb);
  float* lg = logits + (int64_t)b * 2 * nb;
  float best = -INFINITY;
  int k = 0x7fffffff;
  for (int j = lane; j < 2 * nb; j += 32) {
    const float v = o[3 + j];
    lg[j] = v;
    if (j < nb && (v > best || k == 0x7fffffff)) { best = v; k = j; }
  }
  for (int off = 16; off > 0; off >>= 1) {
    const float vo = __shfl_xor_sync(0xffffffffu, best, off);
    const int ko = __shfl_xor_sync(0xffffffffu, k, off);
    if (ko != 0x7fffffff && (k == 0x7fffffff || vo > best || (vo == best && ko < k))) { best = vo; k = ko; }
  }
  if (lane < 3) c2[b * 3 + lane] = o[lane] + c1[b * 3 + lane];
  if (lane == 0) {
    const float pi = 3.14159265358979323846f;
    const float a = (float)k * (2.0f * pi / (float)nb) + o[3 + nb + k] * (pi / (float)nb);
    ang[b] = floor_mod(a + pi, 2.0f * pi) - pi;
    angk[b] = k;
  }
}

// head out [B,3+2nb] -> pred_translations = o[:, :3] + (c2_2 - c2_1) (tp8.py:155), remaining logits (:156)
static __global__ void post_head_kernel(const float* o, const float* c2a, const float* c2b, float* pred_t, float* rem, int B,
                                 int nb) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* r = o + (int64_t)b * (3 + 2 * nb);
  if (lane < 3) pred_t[b * 3 + lane] = r[lane] + (c2b[b * 3 + lane] - c2a[b * 3 + lane]);
  for (int j = lane; j < 2 * nb; j += 32) rem[(int64_t)b * 2 * nb + j] = r[3 + j];
}

// Counter-based keep mask (splitmix64 hash of (seed, site, element)); value 1 with prob keep.
static __global__ void dropout_mask_kernel(float* mask, int64_t count, uint64_t seed, uint32_t site, float keep,
                                           const uint64_t* seed_dev) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  if (seed_dev) seed = *seed_dev;
  uint64_t x = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1) + ((uint64_t)site << 56);
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x = x ^ (x >> 31);
  const float u = (float)(x >> 40) * (1.0f / 16777216.0f);
  mask[i] = u < keep ? 1.f : 0.f;
}

// --------------------------------------------------------------------------------------------
// Backward helpers
// --------------------------------------------------------------------------------------------
// dZ = scale * (dy - sum_dy/R - xhat * sum_dyxhat/R), dy recomputed as in COL_DY; in place on dA.
static __global__ void bn_bwd_apply_kernel(ColArgs a, float* dZ, int64_t ldo, double inv_rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)a.R * a.C;
  if (i >= total) return;
  const int r = (int)(i / a.C), c = (int)(i % a.C);
  const float z = a.Z[(int64_t)r * a.ldz + c];
  float dy = a.dA[(int64_t)r * a.ldd + c];
  if (a.mask) dy *= a.mask[(int64_t)r * a.ldd + c] * a.mask_scale;
  if (!(fmaf(z, a.scale[c], a.shift[c]) > 0.f)) dy = 0.f;
  const float xhat = (z - a.mean[c]) * a.inv[c];
  const float m0 = (float)(a.acc0[c] * inv_rows), m1 = (float)(a.acc1[c] * inv_rows);
  dZ[(int64_t)r * ldo + c] = a.scale[c] * (dy - m0 - xhat * m1);
}

// Eval-mode BN backward is not needed (training only).  d gamma = sum dy*xhat, d beta = sum dy.
static __global__ void bn_bwd_params_kernel(const double* acc0, const double* acc1, float* dgamma, float* dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { dgamma[c] = (float)acc1[c]; dbeta[c] = (float)acc0[c]; }
}

static __global__ void add_double_to_float_kernel(const double* src, float* dst, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) dst[c] += (float)src[c];
}

// scatter pooled gradient to the arg rows: dA[(b*N + idx[b,c]), c] = dG[b,c]   (dA pre-zeroed)
static __global__ void pool_bwd_scatter_kernel(const float* dG, int64_t ldg, const int32_t* idx, float* dA, int N, int C,
                                        int B) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * C) return;
  const int b = (int)(i / C), c = (int)(i % C);
  dA[((int64_t)b * N + idx[i]) * C + c] = dG[(int64_t)b * ldg + c];
}

// Per-cloud reduction of the stage-input gradient dq [B,N,3]:
//   dcenter -= R^T sum_n dq ;  dangle = sum_n (-dq_x*q_y + dq_y*q_x)   (q = rotated input, saved)
// One warp per cloud.  angle == nullptr: plain recentring (stage 2).
static __global__ void input_bwd_kernel(const float* dq, const float* q, const float* angle, float* dcenter, float* dangle,
                                 int N, int B) {
  const int b = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  double s[4] = {0, 0, 0, 0};
  const float* g = dq + (int64_t)b * N * 3;
  const float* p = q + (int64_t)b * N * 3;
  for (int n = lane; n < N; n += 32) {
    const float gx = g[n * 3], gy = g[n * 3 + 1], gz = g[n * 3 + 2];
    s[0] += gx; s[1] += gy; s[2] += gz;
    if (angle) s[3] += (double)(-gx * p[n * 3 + 1] + gy * p[n * 3]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    for (int d = 0; d < 4; ++d) s[d] += __shfl_xor_sync(0xffffffffu, s[d], o);
  if (lane == 0) {
    float gx = (float)s[0], gy = (float)s[1];
    if (angle) {
      float sn, cs;
      sincosf(angle[b], &sn, &cs);
      const float tx = gx * cs + gy * sn, ty = -gx * sn + gy * cs;
      gx = tx; gy = ty;
      dangle[b] = (float)s[3];
    }
    dcenter[b * 3] -= gx;
    dcenter[b * 3 + 1] -= gy;
    dcenter[b * 3 + 2] -= (float)s[2];
  }
}

}  // namespace an3d
