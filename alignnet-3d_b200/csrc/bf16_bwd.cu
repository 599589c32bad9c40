// bf16 fast path, backward: small coefficient / packing kernels and the host orchestration around
// the tcgen05 backward kernels of conv_bwd_bf16.cuh.  FC layers, loss and the inter-stage glue are
// shared with the fp32 path (backward_f32.cu).
#include <algorithm>

#include "bf16_path.cuh"
#include "conv_bwd_bf16.cuh"

namespace an3d {

namespace {

constexpr int kMaxSmem = 227 * 1024;

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// W3 [128][C3] -> C3/64 half-chunk images [128 rows k][64 c] (K-major in c), unfolded
__global__ void pack_w3n_kernel(const float* W3, __nv_bfloat16* img, int C3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C3 * 128) return;
  const int k = i / C3, c = i % C3;
  const int hc = c >> 6, cc = c & 63;
  img[(size_t)hc * 8192 + (cc >> 3) * 1024 + k * 8 + (cc & 7)] = __float2bfloat16_rn(W3[i]);
}

// W2 [64][128] -> image [128 rows][128 k2]; rows 64..127 repeat rows 0..63 so that all four TMEM lane
// quarters of the da1 accumulator carry real data (each warp pair then handles a quarter of the points)
__global__ void pack_w2p_kernel(const float* W2, __nv_bfloat16* img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 128) return;
  const int k1 = i / 128, k2 = i % 128;
  img[(k2 >> 3) * 1024 + k1 * 8 + (k2 & 7)] = __float2bfloat16_rn(W2[(k1 & 63) * 128 + k2]);
}

// dyext = dG * [g > 0]; per-channel sums of dyext and dyext * xhat_ext (BN3 backward)
__global__ void pool_bwd_prep_kernel(const float* dG, int64_t lddg, const float* G, int64_t ldg, const uint32_t* zext,
                                     int B, int C3, const float* gamma, const float* bias, const float* mean,
                                     const float* inv, uint32_t idx_mask, float* dyext, double* red3, int bchunk) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C3) return;
  const int b0 = blockIdx.y * bchunk, b1 = min(B, b0 + bchunk);
  const float sg = gamma[c] < 0.f ? -1.f : 1.f;
  const float bi = bias[c], mu = mean[c], iv = inv[c];
  double s0 = 0.0, s1 = 0.0;
  // four samples per iteration: their twelve loads are independent and in flight together
  for (int bb = b0; bb < b1; bb += 4) {
    float g[4], dg[4];
    uint32_t key[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int b = bb + u;
      g[u] = 0.f; dg[u] = 0.f; key[u] = 0u;
      if (b < b1) {
        g[u] = G[(int64_t)b * ldg + c];
        dg[u] = dG[(int64_t)b * lddg + c];
        key[u] = zext[(int64_t)b * C3 + c];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int b = bb + u;
      if (b >= b1) continue;
      const float d = g[u] > 0.f ? dg[u] : 0.f;
      dyext[(int64_t)b * C3 + c] = d;
      const uint32_t bits = (key[u] & 0x80000000u) ? (key[u] & 0x7fffffffu) : ~key[u];
      const float z = sg * __uint_as_float(bits & ~idx_mask) + bi;
      s0 += (double)d;
      s1 += (double)d * (double)((z - mu) * iv);
    }
  }
  atomicAdd(red3 + c, s0);
  atomicAdd(red3 + C3 + c, s1);
}

// BN3 backward coefficients: dgamma, dbeta; q = -s3 inv3 dgamma / M ; p' = -s3 dbeta / M - q mu3 + q b3
__global__ void bwd3_coeff_kernel(const double* red3, int C3, double count, const float* scale, const float* inv,
                                  const float* mean, const float* bias, float* dgamma, float* dbeta, float* coef3) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C3) return;
  const double db = red3[c], dg = red3[C3 + c];
  dbeta[c] = (float)db;
  dgamma[c] = (float)dg;
  const double s3 = scale[c];
  const double q = -s3 * (double)inv[c] * dg / count;
  const double pp = -s3 * db / count - q * (double)mean[c] + q * (double)bias[c];
  coef3[c] = (float)q;
  coef3[C3 + c] = (float)pp;
}

// Gq[k', k] = sum_c W3b[k',c] q[c] W3b[k,c] (bf16-rounded weights) and u[k] = sum_c p'[c] W3b[k,c].
// Block = 32 x 32 tile of Gq over a 64-channel slice (grid 4 x 4 x C3/64), both weight tiles staged in shared
// memory once; partial sums accumulated in fp32 with atomics ...
__global__ void __launch_bounds__(256) gq_partial_kernel(const float* W3, const float* coef3, int C3, float* gq_f32,
                                                         float* uvec) {
  __shared__ float sa[32][65];   // rows k' (scaled by q)
  __shared__ float sb[32][65];   // rows k
  const int kp0 = blockIdx.x * 32, k0 = blockIdx.y * 32, c0 = blockIdx.z * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 64; i += 256) {
    const int r = i >> 6, cc = i & 63;
    sa[r][cc] = bf16r(W3[(size_t)(kp0 + r) * C3 + c0 + cc]) * coef3[c0 + cc];
    sb[r][cc] = bf16r(W3[(size_t)(k0 + r) * C3 + c0 + cc]);
  }
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
  for (int cc = 0; cc < 64; ++cc) {
    const float wk = sb[tx][cc];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = fmaf(sa[ty + 8 * r][cc], wk, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) atomicAdd(gq_f32 + (size_t)(kp0 + ty + 8 * r) * 128 + k0 + tx, acc[r]);
  if (blockIdx.x == 0 && ty == 0) {
    float u = 0.f;
    for (int cc = 0; cc < 64; ++cc) u = fmaf(coef3[C3 + c0 + cc], sb[tx][cc], u);
    atomicAdd(uvec + k0 + tx, u);
  }
}

// ... then packed as the two K-half bf16 images of the A operand (rows k, contraction k')
__global__ void gq_pack_kernel(const float* gq_f32, __nv_bfloat16* gq_img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 128) return;
  const int kp = i >> 7, k = i & 127;
  const int h = kp >> 6, kk = kp & 63;
  gq_img[(size_t)h * 8192 + (kk >> 3) * 1024 + k * 8 + (kk & 7)] = __float2bfloat16_rn(gq_f32[i]);
}

// grads.W3[k,c] += sa2[k] p'[c] + q[c] * GW[k,c]   (GW = G2 W3b from the forward pass; the sparse part T1 is
// reduced straight into grads.W3 by t1_sparse_kernel)
__global__ void wgrad3_dense_kernel(const float* gw, const double* sa2, const float* coef3, int C3, float* gW3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * C3) return;
  const int k = i / C3, c = i - k * C3;
  // (a reduction, not `+=`: the other branch's kernels and the sparse part T1 add into the same weights concurrently)
  atomicAdd(gW3 + i, (float)sa2[k] * coef3[C3 + c] + coef3[c] * gw[i]);
}

// BN backward coefficients of layers 2 / 1 from (sum dy, sum dy*xhat)
__global__ void bn_bwd_coeff_kernel(const double* red, int C, double count, float* dgamma, float* dbeta, float* coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dbeta[c] = (float)red[2 * c];
  dgamma[c] = (float)red[2 * c + 1];
  coef[2 * c] = (float)(red[2 * c] / count);
  coef[2 * c + 1] = (float)(red[2 * c + 1] / count);
}

// layer-1 backward, finishing step.  bwd_l2_kernel left, per item and channel, S = sum_p dy1 * (1, x, y, z).  With
// xhat1 = alpha . (x, y, z) + alpha0 (affine in the input) and the item's point moments P,
//   A   = sum_p dz1     = s1 (S0 - n m0 - m1 sum_p xhat)
//   B_d = sum_p dz1 q_d = s1 (S_d - m0 P_d - m1 sum_p xhat q_d),   d in {x, y, z}
// give wgrad1 (sum of B over items) and the gradient of the stage input reduced per cloud to d(center) and
// d(angle).  One WARP per item (a block's eight warps walk its items side by side; the one-block-per-item version spent its
// 22 us per launch in three block-wide barriers per item): point moments by warp shuffles, then each lane finishes two of
// the 64 channels.
__global__ void __launch_bounds__(256) bwd_l1_finish_kernel(const float* pcs, const float* center, const float* angle,
                                                            const float* l1sums, int N, int PC, int npc, int n_items,
                                                            int items_per_block, const float* W1, const float* b1,
                                                            const float* mean1, const float* inv1, const float* scale1,
                                                            const float* coef1, float* gW1, float* dcenter, float* dangle,
                                                            int want_input_grad) {
  __shared__ float gsum[8][64][3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float wx[2], wy[2], wz[2], s1[2], m0[2], m1[2], ax[2], ay[2], az[2], a0[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = lane + 32 * h;
    wx[h] = W1[c]; wy[h] = W1[64 + c]; wz[h] = W1[128 + c];
    const float iv = inv1[c];
    s1[h] = scale1[c]; m0[h] = coef1[2 * c]; m1[h] = coef1[2 * c + 1];
    ax[h] = iv * wx[h]; ay[h] = iv * wy[h]; az[h] = iv * wz[h]; a0[h] = iv * (b1[c] - mean1[c]);
  }
  float gwx[2] = {0.f, 0.f}, gwy[2] = {0.f, 0.f}, gwz[2] = {0.f, 0.f};   // wgrad1 partial sums over this warp's items
  const int it_end = min(n_items, (int)(blockIdx.x + 1) * items_per_block);
  for (int it = blockIdx.x * items_per_block + warp; it < it_end; it += 8) {
    const int cloud = it / npc, pchunk = it - cloud * npc;
    const int p0 = pchunk * PC;
    const int nvalid = min(PC, N - p0);
    const int64_t row0 = (int64_t)cloud * N + p0;
    float sn = 0.f, cs = 1.f;
    if (angle) sincosf(angle[cloud], &sn, &cs);
    const float cx = center[cloud * 3], cy = center[cloud * 3 + 1], cz = center[cloud * 3 + 2];
    float m[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // Px Py Pz Pxx Pxy Pxz Pyy Pyz Pzz
    for (int p = lane; p < nvalid; p += 32) {
      const float* src = pcs + (row0 + p) * 3;
      const float x0 = src[0] - cx, y0 = src[1] - cy;
      const float x = x0 * cs - y0 * sn, y = x0 * sn + y0 * cs, z = src[2] - cz;
      m[0] += x; m[1] += y; m[2] += z;
      m[3] = fmaf(x, x, m[3]); m[4] = fmaf(x, y, m[4]); m[5] = fmaf(x, z, m[5]);
      m[6] = fmaf(y, y, m[6]); m[7] = fmaf(y, z, m[7]); m[8] = fmaf(z, z, m[8]);
    }
#pragma unroll
    for (int q = 0; q < 9; ++q)
      for (int o = 16; o > 0; o >>= 1) m[q] += __shfl_xor_sync(0xffffffffu, m[q], o);
    const float n = (float)nvalid;
    float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 S = reinterpret_cast<const float4*>(l1sums)[(size_t)it * 64 + lane + 32 * h];
      const float sxh = ax[h] * m[0] + ay[h] * m[1] + az[h] * m[2] + n * a0[h];
      const float sxh_x = ax[h] * m[3] + ay[h] * m[4] + az[h] * m[5] + a0[h] * m[0];
      const float sxh_y = ax[h] * m[4] + ay[h] * m[6] + az[h] * m[7] + a0[h] * m[1];
      const float sxh_z = ax[h] * m[5] + ay[h] * m[7] + az[h] * m[8] + a0[h] * m[2];
      const float A = s1[h] * (S.x - n * m0[h] - m1[h] * sxh);
      const float Bx = s1[h] * (S.y - m0[h] * m[0] - m1[h] * sxh_x);
      const float By = s1[h] * (S.z - m0[h] * m[1] - m1[h] * sxh_y);
      const float Bz = s1[h] * (S.w - m0[h] * m[2] - m1[h] * sxh_z);
      gwx[h] += Bx; gwy[h] += By; gwz[h] += Bz;
      // d(sum_p dq_d) = sum_c W1[d,c] A_c ; d(angle) = sum_c (-W1[x,c] By_c + W1[y,c] Bx_c)
      f0 = fmaf(wx[h], A, f0); f1 = fmaf(wy[h], A, f1); f2 = fmaf(wz[h], A, f2); f3 += -wx[h] * By + wy[h] * Bx;
    }
    if (want_input_grad) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        f0 += __shfl_xor_sync(0xffffffffu, f0, o); f1 += __shfl_xor_sync(0xffffffffu, f1, o);
        f2 += __shfl_xor_sync(0xffffffffu, f2, o); f3 += __shfl_xor_sync(0xffffffffu, f3, o);
      }
      if (lane == 0) {
        // one writer per (cloud, component) when a cloud is one item; atomics keep multi-item clouds correct
        if (dangle) atomicAdd(dangle + cloud, f3);
        atomicAdd(dcenter + cloud * 3, -(f0 * cs + f1 * sn));
        atomicAdd(dcenter + cloud * 3 + 1, -(-f0 * sn + f1 * cs));
        atomicAdd(dcenter + cloud * 3 + 2, -f2);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    gsum[warp][lane + 32 * h][0] = gwx[h]; gsum[warp][lane + 32 * h][1] = gwy[h]; gsum[warp][lane + 32 * h][2] = gwz[h];
  }
  __syncthreads();
  if (threadIdx.x < 192) {
    const int d = threadIdx.x >> 6, c = threadIdx.x & 63;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += gsum[w][c][d];
    atomicAdd(gW1 + d * 64 + c, t);
  }
}

}  // namespace

int pack_weights_bf16_bwd(const Model& m, const PlanF32& p, const float* params, cudaStream_t st) {
  for (int s = 0; s < 3; ++s) {
    const int C3 = m.conv[s].back().cout;
    pack_w3n_kernel<<<(C3 * 128 + 255) / 256, 256, 0, st>>>(params + m.conv[s][2].w, p.bf.w3n[s], C3);
    AN3D_LAUNCH_CHECK();
    pack_w2p_kernel<<<(128 * 128 + 255) / 256, 256, 0, st>>>(params + m.conv[s][1].w, p.bf.w2p[s]);
    AN3D_LAUNCH_CHECK();
  }
  return AN3D_OK;
}

// dG: gradient w.r.t. the pooled feature [B, C3] (leading dim lddg).  Accumulates parameter gradients
// into `grads`; with want_input_grad adds -R^T sum dq into dcenter [B,3] and sum(...) into dangle [B]
// (dangle must be zero on entry).
int conv_stack_backward_bf16(const Model& m, const PlanF32& p, int s, int br, const float* pcs, const float* center,
                             const float* angle, const float* dG, int64_t lddg, const float* params, float* grads,
                             bool want_input_grad, float* dcenter, float* dangle, cudaStream_t st) {
  const PlanBf16& q = p.bf;
  const BwdScratch& w = q.bw[br];
  const int B = p.B, N = p.N;
  const int64_t M = p.M;
  const Lin &L1 = m.conv[s][0], &L2 = m.conv[s][1], &L3 = m.conv[s][2];
  const int C3 = L3.cout;
  int dev = 0, sms = 148;
  AN3D_CUDA_CHECK(cudaGetDevice(&dev));
  AN3D_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  auto slot = [&](int bn) { return m.bn_slot_off(false, br, bn); };
  auto poff = [&](int bn) { return m.bn_param_off(false, br, bn); };
  const float *sc1 = p.bn.scale + slot(L1.bn), *mean1 = p.bn.mean + slot(L1.bn), *inv1 = p.bn.inv + slot(L1.bn);
  const float *sc2 = p.bn.scale + slot(L2.bn), *mean2 = p.bn.mean + slot(L2.bn), *inv2 = p.bn.inv + slot(L2.bn);
  const float *sc3 = p.bn.scale + slot(L3.bn), *mean3 = p.bn.mean + slot(L3.bn), *inv3 = p.bn.inv + slot(L3.bn);
  const float *gamma1 = params + poff(L1.bn), *beta1 = gamma1 + 64;
  const float *gamma2 = params + poff(L2.bn), *beta2 = gamma2 + 128;
  const float* gamma3 = params + poff(L3.bn);
  const int n_items = B * q.npc;
  const int64_t ldg = s == EMB ? 2 * C3 : C3;

  // ---- layer 3: pooled gradient -> BN3 coefficients ----
  AN3D_CUDA_CHECK(cudaMemsetAsync(w.red3, 0, 2 * (size_t)C3 * sizeof(double), st));
  AN3D_CUDA_CHECK(cudaMemsetAsync(w.red2, 0, 256 * sizeof(double), st));
  AN3D_CUDA_CHECK(cudaMemsetAsync(w.red1, 0, 128 * sizeof(double), st));
  {
    const int bchunk = 16;
    dim3 grid((C3 + 127) / 128, (B + bchunk - 1) / bchunk);
    pool_bwd_prep_kernel<<<grid, 128, 0, st>>>(dG, lddg, p.g[s][br], ldg, q.zext[s][br], B, C3, gamma3, params + L3.b, mean3,
                                               inv3, q.idx_mask, w.dyext, w.red3, bchunk);
    AN3D_LAUNCH_CHECK();
    bwd3_coeff_kernel<<<(C3 + 127) / 128, 128, 0, st>>>(w.red3, C3, (double)M, sc3, inv3, mean3, params + L3.b,
                                                        grads + poff(L3.bn), grads + poff(L3.bn) + C3, w.coef3);
    AN3D_LAUNCH_CHECK();
    AN3D_CUDA_CHECK(cudaMemsetAsync(w.gq_f32, 0, 128 * 128 * sizeof(float), st));
    AN3D_CUDA_CHECK(cudaMemsetAsync(w.uvec, 0, 128 * sizeof(float), st));
    gq_partial_kernel<<<dim3(4, 4, C3 / 64), 256, 0, st>>>(params + L3.w, w.coef3, C3, w.gq_f32, w.uvec);
    AN3D_LAUNCH_CHECK();
    gq_pack_kernel<<<64, 256, 0, st>>>(w.gq_f32, w.gq);
    AN3D_LAUNCH_CHECK();
  }
  // ---- wgrad3: sparse part + Gram on the tensor cores, dense correction on CUDA cores ----
  {
    // (the Gram matrix A2^T A2 was computed on the tensor cores by the forward pass)
    prof_mark(PROF_BWD_T1, true, st);
    // sparse part T1 = A2^T S: gather-scale-accumulate on CUDA cores (1/N of the dense FLOPs)
    convbwd::T1Params T;
    T.a2_img = reinterpret_cast<const uint8_t*>(q.a2img[s][br]); T.img_bytes = (uint32_t)q.img_bytes;
    T.gidx = p.gidx[s][br]; T.dyext = w.dyext; T.s3 = sc3; T.B = B; T.N = N;
    T.PC = q.PC; T.npc = q.npc; T.C3 = C3; T.n_items = n_items; T.t1 = grads + L3.w;
    // 1024 resident threads per SM: one CTA of 1024 channels, or two of <= 512
    const int tr = std::max(1, std::min(n_items, (sms / 4) * (C3 <= 512 ? 2 : 1)));
    T.items_per_cta = (n_items + tr - 1) / tr;
    const size_t tsmem = convbwd::t1_smem_bytes(q.PC);
    AN3D_CUDA_CHECK(cudaFuncSetAttribute(convbwd::t1_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
    convbwd::t1_sparse_kernel<<<dim3(tr, 4), convbwd::t1_threads(C3), tsmem, st>>>(T);
    prof_mark(PROF_BWD_T1, false, st);
    AN3D_LAUNCH_CHECK();
    wgrad3_dense_kernel<<<(128 * C3 + 255) / 256, 256, 0, st>>>(q.gw[s][br], q.sa2[s][br], w.coef3, C3, grads + L3.w);
    AN3D_LAUNCH_CHECK();
  }
  // ---- dgrad3 -> dy2 images + BN2 backward sums ----
  {
    convbwd::Dg3Params D;
    D.a2_img = reinterpret_cast<const uint8_t*>(q.a2img[s][br]); D.dy2_img = reinterpret_cast<uint8_t*>(w.dy2img);
    D.img_bytes = (uint32_t)q.img_bytes; D.gidx = p.gidx[s][br]; D.dyext = w.dyext; D.s3 = sc3; D.gq_img = w.gq;
    D.w3n_img = q.w3n[s]; D.uvec = w.uvec; D.gamma2 = gamma2; D.beta2 = beta2; D.B = B; D.N = N; D.PC = q.PC; D.npc = q.npc;
    D.C3 = C3; D.n_items = n_items; D.red2 = w.red2; D.wstages = convbwd::dg3_wstages(q.PC);
    const int grid = std::min(n_items, sms);
    D.items_per_cta = (n_items + grid - 1) / grid;
    const size_t smem = convbwd::dg3_smem_bytes(q.PC);
    if (smem > (size_t)kMaxSmem) { set_error("dgrad3 tile too large"); return AN3D_ERR_UNSUPPORTED; }
    AN3D_CUDA_CHECK(cudaFuncSetAttribute(convbwd::dgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_mark(PROF_BWD_DGRAD3, true, st);
    convbwd::dgrad3_kernel<<<grid, convbwd::kDg3Threads, smem, st>>>(D);
    prof_mark(PROF_BWD_DGRAD3, false, st);
    AN3D_LAUNCH_CHECK();
    bn_bwd_coeff_kernel<<<1, 128, 0, st>>>(w.red2, 128, (double)M, grads + poff(L2.bn), grads + poff(L2.bn) + 128, w.coef2);
    AN3D_LAUNCH_CHECK();
  }
  // ---- layer 2 backward -> wgrad2, dy1 + BN1 backward sums ----
  {
    convbwd::L2Params P2;
    P2.pcs = pcs; P2.center = center; P2.angle = angle; P2.dy2_img = reinterpret_cast<const uint8_t*>(w.dy2img);
    P2.img_bytes = (uint32_t)q.img_bytes; P2.B = B; P2.N = N; P2.PC = q.PC; P2.npc = q.npc; P2.n_items = n_items;
    const int grid = std::min(n_items, sms);
    P2.items_per_cta = (n_items + grid - 1) / grid;
    P2.w1f = q.w1f[s][br]; P2.c1f = q.c1f[s][br]; P2.W1 = params + L1.w; P2.b1 = params + L1.b; P2.mean1 = mean1;
    P2.inv1 = inv1; P2.gamma1 = gamma1; P2.beta1 = beta1; P2.w2t_img = q.w2t[s]; P2.w2p_img = q.w2p[s];
    P2.b2 = params + L2.b; P2.mean2 = mean2; P2.inv2 = inv2; P2.s2 = sc2; P2.coef2 = w.coef2; P2.gW2 = grads + L2.w;
    P2.l1sums = w.l1sums; P2.red1 = w.red1;
    const size_t smem = convbwd::l2_smem_bytes(q.PC);
    if (smem > (size_t)kMaxSmem) { set_error("bwd_l2 tile too large"); return AN3D_ERR_UNSUPPORTED; }
    AN3D_CUDA_CHECK(cudaFuncSetAttribute(convbwd::bwd_l2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_mark(PROF_BWD_L2, true, st);
    convbwd::bwd_l2_kernel<<<grid, convbwd::kL2Threads, smem, st>>>(P2);
    prof_mark(PROF_BWD_L2, false, st);
    AN3D_LAUNCH_CHECK();
    bn_bwd_coeff_kernel<<<1, 64, 0, st>>>(w.red1, 64, (double)M, grads + poff(L1.bn), grads + poff(L1.bn) + 64, w.coef1);
    AN3D_LAUNCH_CHECK();
  }
  // ---- layer 1 backward (CUDA cores) ----
  {
    const int ipb = std::max(1, (n_items + 4 * sms - 1) / (4 * sms));     // ~4 blocks per SM, wgrad1 atomics once per block
    bwd_l1_finish_kernel<<<(n_items + ipb - 1) / ipb, 256, 0, st>>>(pcs, center, angle, w.l1sums, N, q.PC, q.npc, n_items, ipb,
                                                                  params + L1.w, params + L1.b, mean1, inv1, sc1, w.coef1,
                                                                  grads + L1.w, dcenter, dangle, want_input_grad ? 1 : 0);
    AN3D_LAUNCH_CHECK();
  }
  // biases of conv layers feed a batch-statistics BN: their gradient is identically zero (left at 0).
  return AN3D_OK;
}

}  // namespace an3d
