// _get_loss_separate (models/tp8.py:304-354) forward and its gradient w.r.t. the 8 end_points,
// including the reference's [B,B] broadcasts and the keep-the-larger
// inverted-angle selection (Q5).
#include "common.cuh"
#include "loss.cuh"

namespace an3d {

namespace {

constexpr float kPi = 3.14159265358979323846f;

__device__ __forceinline__ float floor_modf(float x, float y) {  // tf.mod on floats [TF-sem]
  float r = fmodf(x, y);
  if (r != 0.f && ((y < 0.f) != (r < 0.f))) r += y;
  return r;
}

// tf_angle2class (tp8.py:181-199), element-wise
__device__ __forceinline__ void angle2class(float t, int nb, int* cls, float* residual) {
  const float twopi = 2.0f * kPi;
  const float angle = floor_modf(t, twopi);
  const float apc = twopi / (float)nb;
  const float shifted = floor_modf(angle + apc / 2.0f, twopi);
  const int c = (int)(shifted / apc);
  *cls = c;
  *residual = shifted - ((float)c * apc + apc / 2.0f);
}

__device__ __forceinline__ float huber(float e, float delta) {  // tp8.py:173-178
  const float a = fabsf(e);
  const float q = fminf(a, delta);
  return 0.5f * q * q + delta * (a - q);
}

__device__ __forceinline__ float decode_scaled(const float* lg, int nb, int* kout) {  // tp8.py:294-301
  int k = 0;
  float best = lg[0];
  for (int j = 1; j < nb; ++j)
    if (lg[j] > best) { best = lg[j]; k = j; }
  const float res = lg[nb + k] * (kPi / (float)nb);
  const float a = (float)k * (2.0f * kPi / (float)nb) + res;
  *kout = k;
  return floor_modf(a + kPi, 2.0f * kPi) - kPi;
}

__global__ void loss_angles_kernel(LossScratch s, const float* lg1, const float* lg2, const float* a1gt,
                                   const float* a2gt, int B, int nb) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= B) return;
  int k1, k2;
  const float a1 = decode_scaled(lg1 + (int64_t)j * 2 * nb, nb, &k1);
  const float a2 = decode_scaled(lg2 + (int64_t)j * 2 * nb, nb, &k2);
  s.k1[j] = k1;
  s.k2[j] = k2;
  s.pd[j] = a2 - a1;
  s.gt3[j] = a2gt[j] - a1gt[j];
}

__device__ __forceinline__ double block_sum(double v, double* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sm[i];
  return r;  // valid on thread 0
}

struct SampleArgs {
  const float* pred[5];   // s1c1, s1c2, s2c1, s2c2, pred_t
  const float* gt[5];     // c1, c2, c1, c2, translations
  const float* logits[3]; // lg1, lg2, rem
  const float* ang_gt[2]; // pc1_angles, pc2_angles
  float* dend;            // may be null (loss only)
  int B, nb;
  float w_center, w_t3;   // gradient weights of the huber terms
};

__global__ void loss_sample_kernel(LossScratch s, SampleArgs a) {
  __shared__ double sm[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = i < a.B;
  const int B = a.B, nb = a.nb;
  // huber terms
  for (int t = 0; t < 5; ++t) {
    double h = 0.0;
    if (on) {
      const float delta = t == 4 ? 2.0f : 1.0f;
      const float w = t == 4 ? a.w_t3 : a.w_center;
      for (int d = 0; d < 3; ++d) {
        const float e = a.pred[t][i * 3 + d] - a.gt[t][i * 3 + d];
        h += (double)huber(e, delta);
        if (a.dend) a.dend[(int64_t)t * B * 3 + i * 3 + d] = w * fminf(fmaxf(e, -delta), delta);
      }
    }
    const double tot = block_sum(h, sm);
    if (threadIdx.x == 0) atomicAdd(s.sums + t, tot);
  }
}

// angle terms: class target, cross-entropy, selected residual prediction; both variants (target, target + pi).
// One WARP per sample (the thread-per-sample form walked 100 logits with a 400-byte stride between lanes: 40 us at
// B = 4096 on 32 SMs): lanes stride over the classes, so the logit rows are read coalesced; log-sum-exp once per
// instance, kept in `lse` / `mxs` for the gradient kernel.
__global__ void __launch_bounds__(256) loss_ce_kernel(LossScratch s, SampleArgs a) {
  __shared__ double sm[8];
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool on = i < a.B;
  const int B = a.B, nb = a.nb;
  float* lse_store = reinterpret_cast<float*>(s.S);          // [3][B] max, [3][B] 1 / sum exp  (S is otherwise unused)
  for (int inst = 0; inst < 3; ++inst) {
    double ce[2] = {0.0, 0.0};
    if (on) {
      const float* lg = a.logits[inst] + (int64_t)i * 2 * nb;
      float mx = -INFINITY;
      for (int c = lane; c < nb; c += 32) mx = fmaxf(mx, lg[c]);
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float se = 0.f;
      for (int c = lane; c < nb; c += 32) se += expf(lg[c] - mx);
      for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
      if (lane == 0) {
        lse_store[inst * B + i] = mx;
        lse_store[(3 + inst) * B + i] = 1.0f / se;
        for (int v = 0; v < 2; ++v) {
          float target = inst < 2 ? a.ang_gt[inst][i] : (s.gt3[i] - s.pd[0]);  // class from column 0 (Q4)
          if (v) target = target + kPi;
          int cls;
          float res;
          angle2class(target, nb, &cls, &res);
          cls = min(max(cls, 0), nb - 1);
          ce[v] = (double)(logf(se) + mx - lg[cls]);
          const int slot = (inst * 2 + v) * B + i;
          s.cls[slot] = cls;
          s.pred[slot] = lg[nb + cls];
          if (inst < 2) s.lab[slot] = res / (kPi / (float)nb);
        }
      }
    }
    for (int v = 0; v < 2; ++v) {
      // lanes other than 0 contribute zero: block_sum adds the per-warp values
      const double tot = block_sum(ce[v], sm);
      if (threadIdx.x == 0) atomicAdd(s.sums + 5 + inst * 2 + v, tot);
    }
  }
}

// tf.mod for a positive modulus without fmodf's exact (and slow, iterative) remainder: one multiply by the reciprocal, one
// fused multiply-add, two fix-ups.  Within an ulp of x of the exact value; used only inside the O(B^2) pair loop below, where
// every term carries a weight of 1/B^2 (the O(B) class targets keep the exact form).
__device__ __forceinline__ float floor_mod_fast(float x, float y, float inv_y) {
  float r = fmaf(-floorf(x * inv_y), y, x);
  if (r < 0.f) r += y;
  if (r >= y) r -= y;
  return r;
}

// pairwise residual loss: S_j = sum_i huber(pred_j - label_ij), G_j = sum_i huber'(.)
// Block = 128 columns j x one chunk of `ichunk` rows i (staged in shared memory: every thread walks the same rows).
__global__ void __launch_bounds__(128) loss_pair_kernel(LossScratch s, int B, int nb, int ichunk) {
  __shared__ double sm[8];
  __shared__ float srow[256];
  // the two stage-3 variants cost ~3x the others per pair: they go FIRST (blocks are dispatched in ascending z), so the
  // tail of the grid is made of the light blocks
#ifdef AN3D_LOSS_OLD_ORDER
  const int iv = blockIdx.z, inst = iv >> 1, v = iv & 1;
#else
  const int iv = 5 - (int)blockIdx.z, inst = iv >> 1, v = iv & 1;
#endif
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i0 = blockIdx.y * ichunk, i1 = min(B, i0 + ichunk);
  const float* rows = inst < 2 ? s.lab + iv * B : s.gt3;
  for (int i = threadIdx.x; i < i1 - i0; i += blockDim.x) srow[i] = rows[i0 + i];
  __syncthreads();
  float S = 0.f, G = 0.f;
  if (j < B) {
    const float pred = s.pred[iv * B + j];
    const int n = i1 - i0;
    if (inst < 2) {
#pragma unroll 4
      for (int i = 0; i < n; ++i) {
        const float d = pred - srow[i];
        S += huber(d, 1.0f);
        G += fminf(fmaxf(d, -1.0f), 1.0f);
      }
    } else {
      // stage-3 label of pair (i, j): residual of tf_angle2class(gt_i - pd_j [+ pi]) (quirk Q4), normalised by pi / nb
      const float twopi = 2.0f * kPi, inv_twopi = 1.0f / twopi;
      const float apc = twopi / (float)nb, inv_apc = (float)nb / twopi, half = apc / 2.0f;
      const float inv_scale = (float)nb / kPi;
      const float off = (v ? kPi : 0.f) - s.pd[j];
#pragma unroll 4
      for (int i = 0; i < n; ++i) {
        const float angle = floor_mod_fast(srow[i] + off, twopi, inv_twopi);
        const float shifted = floor_mod_fast(angle + half, twopi, inv_twopi);
        const float c = floorf(shifted * inv_apc);
        const float res = shifted - fmaf(c, apc, half);
        const float d = fmaf(-res, inv_scale, pred);
        S += huber(d, 1.0f);
        G += fminf(fmaxf(d, -1.0f), 1.0f);
      }
    }
    atomicAdd(s.G + iv * B + j, (double)G);
  }
  // the loss value only needs sum_j S_j: one reduction per block instead of a [B] vector summed later
  const double tot = block_sum((double)S, sm);
  if (threadIdx.x == 0) atomicAdd(s.sums + 16 + iv, tot);
}

__global__ void loss_final_kernel(LossScratch s, float* loss_out, int B, float esf, float af, int accept_inverted) {
  if (threadIdx.x != 0) return;
  const double* res_sum = s.sums + 16;
  const double invB = 1.0 / B, inv3B = 1.0 / (3.0 * B);
  float total[3], clsl[3], resl[3];
  for (int inst = 0; inst < 3; ++inst) {
    float tv[2], cv[2], rv[2];
    for (int v = 0; v < 2; ++v) {
      cv[v] = (float)(s.sums[5 + inst * 2 + v] * invB);
      rv[v] = (float)(res_sum[inst * 2 + v] * invB * invB);
      tv[v] = cv[v] + 20.0f * rv[v];
    }
    // tf.cond(L > L180, L, L180): keep the larger (Q5)
    const int sel = accept_inverted ? (tv[0] > tv[1] ? 0 : 1) : 0;
    s.sel[inst] = sel;
    total[inst] = tv[sel];
    clsl[inst] = cv[sel];
    resl[inst] = rv[sel];
  }
  const float h0 = (float)(s.sums[0] * inv3B), h1 = (float)(s.sums[1] * inv3B), h2 = (float)(s.sums[2] * inv3B),
              h3 = (float)(s.sums[3] * inv3B), h4 = (float)(s.sums[4] * inv3B);
  const float stage1_t = (h0 + h1) / 2.0f, stage2_t = (h2 + h3) / 2.0f;
  const float stage2_a = (total[0] + total[1]) / 2.0f;
  const float loss_t = esf * (stage1_t + stage2_t) + h4;
  const float loss_a = esf * stage2_a + total[2];
  const float loss = loss_t + af * loss_a;
  loss_out[0] = loss / (float)B;
  loss_out[1] = loss_t;
  loss_out[2] = loss_a;
  loss_out[3] = h0; loss_out[4] = h1; loss_out[5] = h2; loss_out[6] = h3; loss_out[7] = h4;
  for (int inst = 0; inst < 3; ++inst) {
    loss_out[8 + inst * 3] = total[inst];
    loss_out[9 + inst * 3] = clsl[inst];
    loss_out[10 + inst * 3] = resl[inst];
  }
  for (int i = 17; i < 20; ++i) loss_out[i] = 0.f;
}

// d loss / d logits, one WARP per sample (coalesced rows; max and 1 / sum-exp come from loss_ce_kernel)
__global__ void __launch_bounds__(256) loss_grad_kernel(LossScratch s, const float* lg1, const float* lg2, const float* rem,
                                                        float* dend, int B, int nb, float esf, float af) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= B) return;
  const float* logits[3] = {lg1, lg2, rem};
  float* dl[3] = {dend + (int64_t)5 * B * 3, dend + (int64_t)5 * B * 3 + (int64_t)B * 2 * nb,
                  dend + (int64_t)5 * B * 3 + (int64_t)2 * B * 2 * nb};
  const float* lse_store = reinterpret_cast<const float*>(s.S);
  const float invB = 1.0f / (float)B;
  // stage-3 residual label depends on the decoded stage-2 yaws (Q4): d/d pd_j = w3*20*G_j/(B^2 * pi/nb),
  // d a_r / d logit_r[nb + k_r] = pi/nb  ->  the pi/nb cancels.
  const float gpd = af * invB * 20.0f * invB * invB * (float)s.G[(4 + s.sel[2]) * B + j];
  const int k1 = s.k1[j], k2 = s.k2[j];
  for (int inst = 0; inst < 3; ++inst) {
    const int v = s.sel[inst];
    const int slot = (inst * 2 + v) * B + j;
    const float w = (inst < 2 ? af * esf * 0.5f : af) * invB;  // includes per_transform 1/B
    const float* lg = logits[inst] + (int64_t)j * 2 * nb;
    float* d = dl[inst] + (int64_t)j * 2 * nb;
    const int cls = s.cls[slot];
    const float mx = lse_store[inst * B + j], inv_se = lse_store[(3 + inst) * B + j];
    const float gres = w * 20.0f * invB * invB * (float)s.G[slot];
    for (int c = lane; c < nb; c += 32) {
      d[c] = w * invB * (expf(lg[c] - mx) * inv_se - (c == cls ? 1.f : 0.f));
      float r = c == cls ? gres : 0.f;
      if (inst == 1 && c == k2) r += gpd;
      if (inst == 0 && c == k1) r -= gpd;
      d[nb + c] = r;
    }
  }
}

}  // namespace

LossScratch carve_loss_scratch(float* base, int B) {
  LossScratch s;
  char* p = reinterpret_cast<char*>(base);
  s.sums = reinterpret_cast<double*>(p); p += 24 * sizeof(double);
  s.S = reinterpret_cast<double*>(p); p += 6 * (int64_t)B * sizeof(double);
  s.G = reinterpret_cast<double*>(p); p += 6 * (int64_t)B * sizeof(double);
  s.pd = reinterpret_cast<float*>(p); p += (int64_t)B * sizeof(float);
  s.gt3 = reinterpret_cast<float*>(p); p += (int64_t)B * sizeof(float);
  s.pred = reinterpret_cast<float*>(p); p += 6 * (int64_t)B * sizeof(float);
  s.lab = reinterpret_cast<float*>(p); p += 4 * (int64_t)B * sizeof(float);
  s.cls = reinterpret_cast<int*>(p); p += 6 * (int64_t)B * sizeof(int);
  s.k1 = reinterpret_cast<int*>(p); p += (int64_t)B * sizeof(int);
  s.k2 = reinterpret_cast<int*>(p); p += (int64_t)B * sizeof(int);
  s.sel = reinterpret_cast<int*>(p); p += 4 * sizeof(int);
  s.bytes = p - reinterpret_cast<char*>(base);
  return s;
}

int64_t loss_scratch_floats(int B) { return 64 + 48 * (int64_t)B; }

int run_loss(const Model& m, const an3d_labels* lb, const an3d_outputs* out, int B, float* loss_out, float* scratch,
             float* dend, cudaStream_t st) {
  const int nb = m.nb;
  LossScratch s = carve_loss_scratch(scratch, B);
  AN3D_CUDA_CHECK(cudaMemsetAsync(scratch, 0, (24 + 12 * (int64_t)B) * sizeof(double), st));
  const int tb = 128, nblk = (B + tb - 1) / tb;
  loss_angles_kernel<<<nblk, tb, 0, st>>>(s, out->pred_pc1angle_logits, out->pred_pc2angle_logits, lb->pc1_angles,
                                          lb->pc2_angles, B, nb);
  AN3D_LAUNCH_CHECK();
  SampleArgs a;
  a.pred[0] = out->pred_s1_pc1centers; a.pred[1] = out->pred_s1_pc2centers; a.pred[2] = out->pred_s2_pc1centers;
  a.pred[3] = out->pred_s2_pc2centers; a.pred[4] = out->pred_translations;
  a.gt[0] = lb->pc1_centers; a.gt[1] = lb->pc2_centers; a.gt[2] = lb->pc1_centers; a.gt[3] = lb->pc2_centers;
  a.gt[4] = lb->translations;
  a.logits[0] = out->pred_pc1angle_logits; a.logits[1] = out->pred_pc2angle_logits;
  a.logits[2] = out->pred_remaining_angle_logits;
  a.ang_gt[0] = lb->pc1_angles; a.ang_gt[1] = lb->pc2_angles;
  a.dend = dend; a.B = B; a.nb = nb;
  const float invB = 1.0f / (float)B;
  a.w_center = m.arch.early_stage_factor * 0.5f * invB / 3.0f * invB;
  a.w_t3 = invB / 3.0f * invB;
  loss_sample_kernel<<<nblk, tb, 0, st>>>(s, a);
  AN3D_LAUNCH_CHECK();
  loss_ce_kernel<<<(B + 7) / 8, 256, 0, st>>>(s, a);
  AN3D_LAUNCH_CHECK();
  const int ichunk = 256;
  dim3 grid(nblk, (B + ichunk - 1) / ichunk, 6);
  loss_pair_kernel<<<grid, tb, 0, st>>>(s, B, nb, ichunk);
  AN3D_LAUNCH_CHECK();
  loss_final_kernel<<<1, 32, 0, st>>>(s, loss_out, B, m.arch.early_stage_factor, m.arch.angle_factor,
                                       m.arch.accept_inverted_angle);
  AN3D_LAUNCH_CHECK();
  if (dend) {
    loss_grad_kernel<<<(B + 7) / 8, 256, 0, st>>>(s, out->pred_pc1angle_logits, out->pred_pc2angle_logits,
                                          out->pred_remaining_angle_logits, dend, B, nb, m.arch.early_stage_factor,
                                          m.arch.angle_factor);
    AN3D_LAUNCH_CHECK();
  }
  return AN3D_OK;
}

}  // namespace an3d
