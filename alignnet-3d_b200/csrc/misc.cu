// Flat-buffer TF-Adam, host-decode of angle logits, batched z-axis rigid transforms.
#include <cmath>

#include <algorithm>

#include "common.cuh"

namespace an3d {
namespace {

constexpr float kPi = 3.14159265358979323846f;

// tf.train.AdamOptimizer [TF-sem]: m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ;
// p <- p - lr_t m / (sqrt(v) + eps),  lr_t = lr sqrt(1-b2^t)/(1-b1^t)  (train.py:212-217)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr_t, float grad_scale, float b1, float b2,
                            float eps, const int64_t* __restrict__ step_dev, float lr) {
  if (step_dev) {   // graph-replay form: bias correction from the device-resident step count
    const double t = (double)*step_dev;
    lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
  }
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    const float4 gg = *reinterpret_cast<const float4*>(g + i4);
    float4 mm = *reinterpret_cast<float4*>(m + i4), vv = *reinterpret_cast<float4*>(v + i4),
           pp = *reinterpret_cast<float4*>(p + i4);
    const float ga[4] = {gg.x * grad_scale, gg.y * grad_scale, gg.z * grad_scale, gg.w * grad_scale};
    float* ma = &mm.x; float* va = &vv.x; float* pa = &pp.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ma[k] = b1 * ma[k] + (1.f - b1) * ga[k];
      va[k] = b2 * va[k] + (1.f - b2) * ga[k] * ga[k];
      pa[k] = pa[k] - lr_t * ma[k] / (sqrtf(va[k]) + eps);
    }
    *reinterpret_cast<float4*>(m + i4) = mm;
    *reinterpret_cast<float4*>(v + i4) = vv;
    *reinterpret_cast<float4*>(p + i4) = pp;
  } else {
    for (int64_t i = i4; i < n; ++i) {
      const float gi = g[i] * grad_scale;
      const float mi = b1 * m[i] + (1.f - b1) * gi;
      const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    }
  }
}

// tf.train.MomentumOptimizer(lr, momentum) [TF-sem, ApplyMomentum with use_nesterov = false]: accum <- momentum accum + g ;
// p <- p - lr accum  (train.py:211-212)
__global__ void momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ acc, int64_t n,
                                float lr, float momentum, float grad_scale) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    const float4 gg = *reinterpret_cast<const float4*>(g + i4);
    float4 aa = *reinterpret_cast<float4*>(acc + i4), pp = *reinterpret_cast<float4*>(p + i4);
    aa.x = momentum * aa.x + gg.x * grad_scale; pp.x = pp.x - lr * aa.x;
    aa.y = momentum * aa.y + gg.y * grad_scale; pp.y = pp.y - lr * aa.y;
    aa.z = momentum * aa.z + gg.z * grad_scale; pp.z = pp.z - lr * aa.z;
    aa.w = momentum * aa.w + gg.w * grad_scale; pp.w = pp.w - lr * aa.w;
    *reinterpret_cast<float4*>(acc + i4) = aa;
    *reinterpret_cast<float4*>(p + i4) = pp;
  } else {
    for (int64_t i = i4; i < n; ++i) {
      const float ai = momentum * acc[i] + g[i] * grad_scale;
      acc[i] = ai;
      p[i] = p[i] - lr * ai;
    }
  }
}

__device__ __forceinline__ float floor_modf(float x, float y) {
  float r = fmodf(x, y);
  if (r != 0.f && ((y < 0.f) != (r < 0.f))) r += y;
  return r;
}

__global__ void decode_angles_kernel(const float* logits, float* angles, int B, int nb, int scaled) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* lg = logits + (int64_t)b * 2 * nb;
  int k = 0;
  float best = lg[0];
  for (int j = 1; j < nb; ++j)
    if (lg[j] > best) { best = lg[j]; k = j; }
  const float apc = 2.0f * kPi / (float)nb;
  if (scaled == 1) {  // tf_get_angles, models/tp8.py:294-301
    const float a = (float)k * apc + lg[nb + k] * (kPi / (float)nb);
    angles[b] = floor_modf(a + kPi, 2.0f * kPi) - kPi;
  } else if (scaled == 2) {  // tf_classLogits2angle -> tf_class2angle2, models/tp8.py:213-226,248-251
    const float a = (float)k * apc + lg[nb + k];
    angles[b] = floor_modf(a + kPi, 2.0f * kPi) - kPi;
  } else {       // classLogits2angle, models/tp8.py:229-244 (quirk Q1: unscaled residual)
    float a = (float)k * apc + lg[nb + k];
    if (a > kPi) a -= 2.0f * kPi;
    angles[b] = a;
  }
}

// p' = Rz(theta)(p - c) + c + t  (tp_utils/pointcloud.py:279-298)
__global__ void rigid_apply_kernel(const float* __restrict__ pts, const float* __restrict__ t,
                                   const float* __restrict__ ang, const float* __restrict__ ctr,
                                   float* __restrict__ out, int N, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / N);
  float cx = 0.f, cy = 0.f, cz = 0.f, tx = 0.f, ty = 0.f, tz = 0.f, s = 0.f, c = 1.f;
  if (ctr) { cx = ctr[b * 3]; cy = ctr[b * 3 + 1]; cz = ctr[b * 3 + 2]; }
  if (t) { tx = t[b * 3]; ty = t[b * 3 + 1]; tz = t[b * 3 + 2]; }
  if (ang) sincosf(ang[b], &s, &c);
  const float x = pts[i * 3] - cx, y = pts[i * 3 + 1] - cy, z = pts[i * 3 + 2] - cz;
  out[i * 3] = (c * x - s * y) + cx + tx;
  out[i * 3 + 1] = (s * x + c * y) + cy + ty;
  out[i * 3 + 2] = z + cz + tz;
}

// a21: tf_transform_pcs (models/tp8.py:361-371) as coded.  tf_translate_pcs (:357-358) returns the TILED TRANSLATION (quirk
// Q6), so each translate step replaces the point; the rotation is a row vector times tf_get_rotation_matrix_z (:26-27).
__global__ void transform_pcs_q6_kernel(const float* __restrict__ pcs, const float* __restrict__ t,
                                        const float* __restrict__ ang, const float* __restrict__ ctr,
                                        float* __restrict__ out, int N, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / N);
  float x = pcs[i * 3], y = pcs[i * 3 + 1], z = pcs[i * 3 + 2];
  if (ctr) { x = -ctr[b * 3]; y = -ctr[b * 3 + 1]; z = -ctr[b * 3 + 2]; }
  if (ang) {
    float s, c;
    sincosf(ang[b], &s, &c);
    const float nx = x * c + y * s, ny = -x * s + y * c;     // [x y z] [[c,-s,0],[s,c,0],[0,0,1]]
    x = nx; y = ny;
  }
  if (t) { x = -t[b * 3]; y = -t[b * 3 + 1]; z = -t[b * 3 + 2]; }
  if (ctr) { x = ctr[b * 3]; y = ctr[b * 3 + 1]; z = ctr[b * 3 + 2]; }
  out[i * 3] = x; out[i * 3 + 1] = y; out[i * 3 + 2] = z;
}

// a21: point_distances = tf.norm(a - g, axis=1) -> [B,3] (the norm runs over the POINT axis, tp8.py:386);
// acc += sum_d point_distances[b,d]^2.  One block per cloud.
__global__ void __launch_bounds__(128) p2p_norm_kernel(const float* __restrict__ a, const float* __restrict__ g, int N,
                                                       double* acc) {
  __shared__ float sm[3][4];
  const int b = blockIdx.x;
  float s[3] = {0.f, 0.f, 0.f};
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const int64_t i = ((int64_t)b * N + n) * 3;
#pragma unroll
    for (int d = 0; d < 3; ++d) { const float e = a[i + d] - g[i + d]; s[d] = fmaf(e, e, s[d]); }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    for (int o = 16; o > 0; o >>= 1) s[d] += __shfl_xor_sync(0xffffffffu, s[d], o);
    if ((threadIdx.x & 31) == 0) sm[d][threadIdx.x >> 5] = s[d];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int d = 0; d < 3; ++d) {
      const float nrm = sqrtf(sm[d][0] + sm[d][1] + sm[d][2] + sm[d][3]);
      tot += (double)(nrm * nrm);
    }
    atomicAdd(acc, tot);
  }
}
__global__ void p2p_final_kernel(const double* acc, int B, float* loss_out) {
  const float loss = (float)(acc[0] / (3.0 * B));       // reduce_mean over [B,3]; min(loss, loss_180) = loss (:388-393)
  loss_out[0] = loss / (float)B;
  loss_out[1] = loss;
}

// t' = -d + Rz(theta) d + t, d = c_new - c_old  (tp_utils/pointcloud.py:309-318)
__global__ void recenter_kernel(const float* t, const float* ang, const float* c_old, const float* c_new, float* out,
                                int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float dx = c_new[i * 3] - c_old[i * 3], dy = c_new[i * 3 + 1] - c_old[i * 3 + 1],
              dz = c_new[i * 3 + 2] - c_old[i * 3 + 2];
  float s, c;
  sincosf(ang[i], &s, &c);
  out[i * 3] = -dx + (c * dx - s * dy) + t[i * 3];
  out[i * 3 + 1] = -dy + (s * dx + c * dy) + t[i * 3 + 1];
  out[i * 3 + 2] = -dz + dz + t[i * 3 + 2];
}

}  // namespace
}  // namespace an3d

using namespace an3d;

extern "C" {

int an3d_adam_step(float* params, const float* grads, float* m, float* v, int64_t count, float lr, int64_t step,
                   float grad_scale, float beta1, float beta2, float eps, void* stream) {
  if (!params || !grads || !m || !v || count < 0 || step < 1) {
    set_error("an3d_adam_step: bad argument (step must be >= 1)");
    return AN3D_ERR_INVALID;
  }
  if (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)m | (uintptr_t)v) & 15) {
    set_error("an3d_adam_step: buffers must be 16-byte aligned");
    return AN3D_ERR_ALIGN;
  }
  AN3D_TRY(check_device());
  const double lr_t = (double)lr * std::sqrt(1.0 - std::pow((double)beta2, (double)step)) /
                      (1.0 - std::pow((double)beta1, (double)step));
  const int64_t nthreads = (count + 3) / 4;
  if (nthreads == 0) return AN3D_OK;
  adam_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, grads, m, v, count,
                                                                                   (float)lr_t, grad_scale, beta1,
                                                                                   beta2, eps, nullptr, lr);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

// ---- yaw-constrained point-to-point ICP (icp.py:69-78) -------------------------------------------------------------
constexpr int kIcpThreads = 256;
constexpr int kIcpTile = 1024;        // target points staged per shared-memory tile

static __device__ __forceinline__ double block_sum_d(double v, double* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < kIcpThreads / 32; ++i) r += sm[i];
  return r;                            // every thread gets the total
}

static __global__ void __launch_bounds__(kIcpThreads) icp_yaw_kernel(const float* src, const int64_t* src_off,
                                                                     const int32_t* src_n, const float* tgt,
                                                                     const int64_t* tgt_off, const int32_t* tgt_n,
                                                                     const float* init, float radius, int its, float* out,
                                                                     float* stats) {
  __shared__ float st[kIcpTile * 3];
  __shared__ double red[kIcpThreads / 32];
  const int pair = blockIdx.x;
  const float* S = src + src_off[pair] * 3;
  const float* Q = tgt + tgt_off[pair] * 3;
  const int ns = src_n[pair], nt = tgt_n[pair];
  double T[12];                        // rows of [R | t], replicated in every thread
  for (int i = 0; i < 12; ++i) T[i] = (double)init[pair * 16 + i];
  double fitness = 0.0, rmse = 0.0;
  int it_done = 0;
  const float r2 = radius * radius;
  // correspondences of the current transform: per source point the nearest target (index kept as coordinates)
  // sums over inliers: count, sum d2, sum p, sum q, sum px*qx + py*qy, sum px*qy - py*qx, (centred later)
  auto pass = [&](double (&acc)[11]) {
    for (int i = 0; i < 11; ++i) acc[i] = 0.0;
    for (int base = 0; base < ns; base += kIcpThreads) {
      const int i = base + threadIdx.x;
      float px = 0.f, py = 0.f, pz = 0.f;
      if (i < ns) {
        const float x = S[3 * i], y = S[3 * i + 1], z = S[3 * i + 2];
        px = (float)(T[0] * x + T[1] * y + T[2] * z + T[3]);
        py = (float)(T[4] * x + T[5] * y + T[6] * z + T[7]);
        pz = (float)(T[8] * x + T[9] * y + T[10] * z + T[11]);
      }
      float best = 3.0e38f, bx = 0.f, by = 0.f, bz = 0.f;
      for (int t0 = 0; t0 < nt; t0 += kIcpTile) {
        const int cnt = min(kIcpTile, nt - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt * 3; k += kIcpThreads) st[k] = Q[(size_t)t0 * 3 + k];
        __syncthreads();
        if (i < ns) {
          for (int k = 0; k < cnt; ++k) {
            const float dx = st[3 * k] - px, dy = st[3 * k + 1] - py, dz = st[3 * k + 2] - pz;
            const float d = dx * dx + dy * dy + dz * dz;
            if (d < best) { best = d; bx = st[3 * k]; by = st[3 * k + 1]; bz = st[3 * k + 2]; }
          }
        }
      }
      if (i < ns && best <= r2) {
        acc[0] += 1.0; acc[1] += (double)best;
        acc[2] += px; acc[3] += py; acc[4] += pz; acc[5] += bx; acc[6] += by; acc[7] += bz;
        acc[8] += (double)px * bx + (double)py * by;
        acc[9] += (double)px * by - (double)py * bx;
      }
    }
    for (int i = 0; i < 10; ++i) acc[i] = block_sum_d(acc[i], red);
  };
  double acc[11];
  if (ns > 0 && nt > 0) {
    pass(acc);
    fitness = acc[0] / ns;
    rmse = acc[0] > 0.0 ? sqrt(acc[1] / acc[0]) : 0.0;
    for (int it = 0; it < its; ++it) {
      const double n = acc[0];
      if (n <= 0.0) break;
      const double pbx = acc[2] / n, pby = acc[3] / n, pbz = acc[4] / n, qbx = acc[5] / n, qby = acc[6] / n, qbz = acc[7] / n;
      // centred sums: sum (p - pb).(q - qb) = sum p.q - n pb.qb  (xy only), likewise the cross term
      const double dot = acc[8] - n * (pbx * qbx + pby * qby), crs = acc[9] - n * (pbx * qby - pby * qbx);
      const double th = atan2(crs, dot), c = cos(th), s = sin(th);
      const double ux = qbx - (c * pbx - s * pby), uy = qby - (s * pbx + c * pby), uz = qbz - pbz;
      double N[12];                    // U * T with U = [Rz(th) | u]
      for (int j = 0; j < 4; ++j) {
        N[j] = c * T[j] - s * T[4 + j];
        N[4 + j] = s * T[j] + c * T[4 + j];
        N[8 + j] = T[8 + j];
      }
      N[3] += ux; N[7] += uy; N[11] += uz;
      for (int j = 0; j < 12; ++j) T[j] = N[j];
      it_done = it + 1;
      const double pf = fitness, pr = rmse;
      pass(acc);
      fitness = acc[0] / ns;
      rmse = acc[0] > 0.0 ? sqrt(acc[1] / acc[0]) : 0.0;
      if (fabs(pf - fitness) < 1e-6 && fabs(pr - rmse) < 1e-6) break;
    }
  }
  if (threadIdx.x < 12) out[pair * 16 + threadIdx.x] = (float)T[threadIdx.x];
  if (threadIdx.x >= 12 && threadIdx.x < 16) out[pair * 16 + threadIdx.x] = threadIdx.x == 15 ? 1.f : 0.f;
  if (threadIdx.x == 0) { stats[pair * 3] = (float)fitness; stats[pair * 3 + 1] = (float)rmse; stats[pair * 3 + 2] = (float)it_done; }
}

int an3d_icp_yaw(const float* src, const int64_t* src_off, const int32_t* src_n, const float* tgt, const int64_t* tgt_off,
                 const int32_t* tgt_n, const float* init, int32_t pairs, float radius, int32_t its, float* out,
                 float* stats, void* stream) {
  if (!src_off || !src_n || !tgt_off || !tgt_n || !init || !out || !stats || pairs < 0 || its < 0 || !(radius > 0.f)) {
    set_error("an3d_icp_yaw: bad argument");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_device());
  if (pairs == 0) return AN3D_OK;
  icp_yaw_kernel<<<pairs, kIcpThreads, 0, (cudaStream_t)stream>>>(src, src_off, src_n, tgt, tgt_off, tgt_n, init, radius, its,
                                                                out, stats);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

// ---- batch assembly (provider.py:60-71,97-98,125-126) -----------------------------------------------------------
static __global__ void resample_gather_kernel(const float* __restrict__ pts, const int64_t* __restrict__ off,
                                              const int32_t* __restrict__ idx, int64_t total, int N, int stride,
                                              const float* __restrict__ jitter, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / N);
  const int r = idx[i];
  float x = 0.f, y = 0.f, z = 0.f;
  if (r >= 0) {
    const float* src = pts + (off[b] + r) * stride;
    x = src[0]; y = src[1]; z = src[2];
  }
  if (jitter) { x += jitter[3 * i]; y += jitter[3 * i + 1]; z += jitter[3 * i + 2]; }
  out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z;
}

int an3d_resample_gather(const float* points, const int64_t* cloud_offset, const int32_t* sample_idx, int32_t batch,
                         int32_t num_points, int32_t stride, const float* jitter, float* out, void* stream) {
  if (!cloud_offset || !sample_idx || !out || batch < 0 || num_points < 0 || stride < 3) {
    set_error("an3d_resample_gather: bad argument");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_device());
  const int64_t total = (int64_t)batch * num_points;
  if (total == 0) return AN3D_OK;
  resample_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(points, cloud_offset, sample_idx,
                                                                                           total, num_points, stride, jitter, out);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

// ---- evaluation metrics (evaluation.py:16-46,128-211) -----------------------------------------------------------
static __device__ __forceinline__ double floor_mod_d(double x, double y) {
  double r = fmod(x, y);
  if (r != 0.0 && ((y < 0.0) != (r < 0.0))) r += y;
  return r;
}
static __device__ __forceinline__ double angle_diff_d(double a, double b) {       // evaluation.py:26-28
  const double pi = 3.14159265358979323846;
  return floor_mod_d(b - a + pi, 2.0 * pi) - pi;
}

static __global__ void __launch_bounds__(256) evaluate_kernel(const double* pt, const double* pa, const double* pc,
                                                              const double* gt, const double* ga, const double* gc,
                                                              const uint8_t* is_test, int n, int inverted, double* acc) {
  __shared__ double sacc[3 * 5 * 14];
  for (int i = threadIdx.x; i < 210; i += blockDim.x) sacc[i] = 0.0;
  __syncthreads();
  const double pi = 3.14159265358979323846;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    // translate_transform_to_new_center_of_rotation (pointcloud.py:309-318): t' = -d + Rz(a) d + t, d = c_gt - c_pred
    const double a = pa[i];
    const double dx = gc[3 * i] - pc[3 * i], dy = gc[3 * i + 1] - pc[3 * i + 1];
    const double cs = cos(a), sn = sin(a);
    const double tx = -dx + (cs * dx - sn * dy) + pt[3 * i], ty = -dy + (sn * dx + cs * dy) + pt[3 * i + 1];
    const double ex = tx - gt[3 * i], ey = ty - gt[3 * i + 1];
    const double dist_t = sqrt(ex * ex + ey * ey);
    double dist_a = fabs(angle_diff_d(a, ga[i])) / pi * 180.0;
    if (inverted) dist_a = fmin(dist_a, fabs(angle_diff_d(a + pi, ga[i])) / pi * 180.0);
    if (dist_t > 10000.0) continue;                                               // evaluation.py:168
    const double lt[3] = {dist_t < 0.02 ? 1.0 : 0.0, dist_t < 0.1 ? 1.0 : 0.0, dist_t < 0.2 ? 1.0 : 0.0};
    const double la[3] = {dist_a < 1.0 ? 1.0 : 0.0, dist_a < 5.0 ? 1.0 : 0.0, dist_a < 10.0 ? 1.0 : 0.0};
    const double row[14] = {1.0, lt[0], lt[1], lt[2], dist_t, dist_t * dist_t, la[0], la[1], la[2], dist_a, dist_a * dist_a,
                            fmin(lt[0], la[0]), fmin(lt[1], la[1]), fmin(lt[2], la[2])};
    const double cd = sqrt(gc[3 * i] * gc[3 * i] + gc[3 * i + 1] * gc[3 * i + 1] + gc[3 * i + 2] * gc[3 * i + 2]);
    const bool test = is_test ? is_test[i] != 0 : false;
    const double lim[5] = {1e300, 5.0, 10.0, 15.0, 20.0};
    for (int s = 0; s < 3; ++s) {
      if ((s == 1 && test) || (s == 2 && !test)) continue;
      for (int r = 0; r < 5; ++r) {
        if (cd > lim[r]) continue;
        for (int q = 0; q < 14; ++q)
          if (row[q] != 0.0) atomicAdd(&sacc[(s * 5 + r) * 14 + q], row[q]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 210; i += blockDim.x)
    if (sacc[i] != 0.0) atomicAdd(acc + i, sacc[i]);
}

int an3d_evaluate(const double* pred_translations, const double* pred_angles, const double* pred_centers,
                  const double* gt_translations, const double* gt_angles, const double* gt_pc1centers,
                  const uint8_t* is_test, int32_t n, int32_t accept_inverted_angle, double* acc, void* stream) {
  if (!pred_translations || !pred_angles || !pred_centers || !gt_translations || !gt_angles || !gt_pc1centers || !acc || n < 0) {
    set_error("an3d_evaluate: bad argument");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_device());
  AN3D_CUDA_CHECK(cudaMemsetAsync(acc, 0, 210 * sizeof(double), (cudaStream_t)stream));
  if (n == 0) return AN3D_OK;
  const int blocks = std::min(296, (n + 255) / 256);
  evaluate_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred_translations, pred_angles, pred_centers, gt_translations,
                                                          gt_angles, gt_pc1centers, is_test, n, accept_inverted_angle, acc);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

static __global__ void step_advance_kernel(int64_t* step, uint64_t* seed, uint64_t seed_base) {
  const int64_t t = *step + 1;
  *step = t;
  if (seed) *seed = seed_base + (uint64_t)t;
}

int an3d_step_advance(int64_t* step_dev, uint64_t* seed_dev, uint64_t seed_base, void* stream) {
  if (!step_dev) {
    set_error("an3d_step_advance: step_dev is null");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_device());
  step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev, seed_dev, seed_base);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

int an3d_adam_step_dev(float* params, const float* grads, float* m, float* v, int64_t count, float lr,
                       const int64_t* step_dev, float grad_scale, float beta1, float beta2, float eps, void* stream) {
  if (!params || !grads || !m || !v || count < 0 || !step_dev) {
    set_error("an3d_adam_step_dev: bad argument");
    return AN3D_ERR_INVALID;
  }
  if (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)m | (uintptr_t)v) & 15) {
    set_error("an3d_adam_step_dev: buffers must be 16-byte aligned");
    return AN3D_ERR_ALIGN;
  }
  AN3D_TRY(check_device());
  const int64_t nthreads = (count + 3) / 4;
  if (nthreads == 0) return AN3D_OK;
  adam_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, grads, m, v, count, 0.f,
                                                                                   grad_scale, beta1, beta2, eps, step_dev,
                                                                                   lr);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

int an3d_momentum_step(float* params, const float* grads, float* accum, int64_t count, float lr, float momentum,
                       float grad_scale, void* stream) {
  if (!params || !grads || !accum || count < 0) {
    set_error("an3d_momentum_step: bad argument");
    return AN3D_ERR_INVALID;
  }
  if (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)accum) & 15) {
    set_error("an3d_momentum_step: buffers must be 16-byte aligned");
    return AN3D_ERR_ALIGN;
  }
  AN3D_TRY(check_device());
  const int64_t nthreads = (count + 3) / 4;
  if (nthreads == 0) return AN3D_OK;
  momentum_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, grads, accum, count, lr,
                                                                                       momentum, grad_scale);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

int an3d_decode_angles(const float* logits, float* angles, int32_t batch, int32_t num_bins, int32_t scaled,
                       void* stream) {
  if (!logits || !angles || batch < 0 || num_bins < 1) {
    set_error("an3d_decode_angles: bad argument");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_device());
  if (batch == 0) return AN3D_OK;
  decode_angles_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(logits, angles, batch, num_bins, scaled);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

int an3d_rigid_apply(const float* pts, const float* translation, const float* angle, const float* center, float* out,
                     int32_t batch, int32_t num_points, void* stream) {
  if (!pts || !out || batch < 0 || num_points < 0) {
    set_error("an3d_rigid_apply: bad argument");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_device());
  const int64_t total = (int64_t)batch * num_points;
  if (total == 0) return AN3D_OK;
  rigid_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pts, translation, angle, center,
                                                                                       out, num_points, total);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

int an3d_transform_pcs(const float* pcs, const float* translations, const float* angles, const float* rotation_centers,
                       float* out, int32_t batch, int32_t num_points, void* stream) {
  if (!pcs || !out || batch < 0 || num_points < 0) {
    set_error("an3d_transform_pcs: bad argument");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_device());
  const int64_t total = (int64_t)batch * num_points;
  if (total == 0) return AN3D_OK;
  transform_pcs_q6_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      pcs, translations, angles, rotation_centers, out, num_points, total);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

int an3d_loss_p2p(const float* pcs1, const float* pred_translations, const float* pred_angles,
                  const float* pred_s2_pc1centers, const float* translations, const float* rel_angles,
                  const float* pc1_centers, int32_t batch, int32_t num_points, float* loss_out, void* workspace,
                  int64_t workspace_bytes, void* stream) {
  if (!pcs1 || !pred_translations || !pred_angles || !pred_s2_pc1centers || !translations || !rel_angles || !pc1_centers ||
      !loss_out || !workspace || batch < 1 || num_points < 1) {
    set_error("an3d_loss_p2p: bad argument");
    return AN3D_ERR_INVALID;
  }
  const int64_t cloud = (int64_t)batch * num_points * 3;
  if (workspace_bytes < (int64_t)(2 * cloud * sizeof(float) + 16)) {
    set_error("an3d_loss_p2p: workspace too small: need %lld bytes", (long long)(2 * cloud * sizeof(float) + 16));
    return AN3D_ERR_WORKSPACE;
  }
  AN3D_TRY(check_device());
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = static_cast<double*>(workspace);
  float* a = reinterpret_cast<float*>(acc + 2);
  float* g = a + cloud;
  AN3D_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(double), st));
  AN3D_TRY(an3d_transform_pcs(pcs1, pred_translations, pred_angles, pred_s2_pc1centers, a, batch, num_points, stream));
  AN3D_TRY(an3d_transform_pcs(pcs1, translations, rel_angles, pc1_centers, g, batch, num_points, stream));
  p2p_norm_kernel<<<batch, 128, 0, st>>>(a, g, num_points, acc);
  AN3D_LAUNCH_CHECK();
  p2p_final_kernel<<<1, 1, 0, st>>>(acc, batch, loss_out);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

int an3d_recenter_translations(const float* translations, const float* angles, const float* old_centers,
                               const float* new_centers, float* out, int32_t count, void* stream) {
  if (!translations || !angles || !old_centers || !new_centers || !out || count < 0) {
    set_error("an3d_recenter_translations: bad argument");
    return AN3D_ERR_INVALID;
  }
  AN3D_TRY(check_device());
  if (count == 0) return AN3D_OK;
  recenter_kernel<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(translations, angles, old_centers, new_centers,
                                                                         out, count);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

}  // extern "C"
