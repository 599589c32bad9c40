// fp32 parity path: forward pass of get_model (models/tp8.py:135-158) with CUDA-core kernels.
#include <algorithm>
#include <cstdlib>

#include "kernels_f32.cuh"
#include "bf16_path.cuh"
#include "fc2_gemm.cuh"
#include "gemm_tc.cuh"

namespace an3d {

struct BnView {
  const float *gamma, *beta;
  float *state_mean, *state_var;
  float *scale, *shift, *mean, *inv;
  double *acc0, *acc1;
  int ch;
};

static BnView bn_view(const Model& m, const PlanF32& p, const float* params, float* state, bool head, int br, int bn) {
  BnView v;
  const int ch = head ? m.bn_head[bn].ch : m.bn_branch[bn].ch;
  const int64_t po = m.bn_param_off(head, br, bn), so = m.bn_state_off(head, br, bn), sl = m.bn_slot_off(head, br, bn);
  v.gamma = params + po;
  v.beta = params + po + ch;
  v.state_mean = state ? state + so : nullptr;
  v.state_var = state ? state + so + ch : nullptr;
  v.scale = p.bn.scale + sl;
  v.shift = p.bn.shift + sl;
  v.mean = p.bn.mean + sl;
  v.inv = p.bn.inv + sl;
  v.acc0 = p.bn.acc0 + sl;
  v.acc1 = p.bn.acc1 + sl;
  v.ch = ch;
  return v;
}

// ---- the two siamese branches of a conv stage on two streams (on by default; AN3D_TWO_STREAMS=0 turns it off) ----
// Measured on B200, c3 training step: 7.21 -> 6.97 ms (profiles/r2_ab_switches.txt).
// The branches of a stage are independent and touch disjoint scratch ([stage][branch] buffers), so the ~10 small
// launches around one branch's persistent kernels (moments, folds, statistics, pool finalize; ~75 us per stage and
// branch) can run under the other branch's persistent kernels, which leave threads and registers free on every SM.
// Fork / join with events, so the pattern is also legal inside a stream capture.  The side stream and events are
// created on first use: run one eager step before capturing a graph (Engine._capture does).
SideStream* side_stream() {
  static SideStream per_device[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStream& x = per_device[dev];
  if (!x.ok) {
    if (cudaStreamCreateWithFlags(&x.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    x.ok = true;
  }
  return &x;
}
// Per-kernel profiling (an3d_profile_begin) serialises the branches: an event pair around a kernel that shares the
// device with the other branch's kernels would time the overlap, not the kernel.
bool two_streams_enabled() {
  static const bool on = [] {
    const char* e = getenv("AN3D_TWO_STREAMS");
    return !(e != nullptr && e[0] == '0');
  }();
  return on && !prof_active();
}
static int conv_stage_two_streams(SideStream* ss, const Model& m, const PlanF32& p, int s, const float* const pcs[2],
                                  const float* const center[2], const float* const angle[2], const float* params,
                                  float* state, bool training, float decay, cudaStream_t st) {
  AN3D_CUDA_CHECK(cudaEventRecord(ss->fork, st));
  AN3D_CUDA_CHECK(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
  const int r0 = conv_stack_forward_bf16(m, p, s, 0, pcs[0], center[0], angle ? angle[0] : nullptr, params, state, training,
                                         decay, st);
  const int r1 = conv_stack_forward_bf16(m, p, s, 1, pcs[1], center[1], angle ? angle[1] : nullptr, params, state, training,
                                         decay, ss->stream);
  AN3D_CUDA_CHECK(cudaEventRecord(ss->join, ss->stream));      // always join: a capture must not end with a dangling fork
  AN3D_CUDA_CHECK(cudaStreamWaitEvent(st, ss->join, 0));
  return r0 != AN3D_OK ? r0 : r1;
}

// bf16 mode: the producing GEMM's epilogue already accumulated sum z (acc0) and sum z^2 (acc1) per column
static __global__ void bn_finalize_sums_kernel(const double* acc0, const double* acc1, double inv_rows, const float* gamma,
                                               const float* beta, float* state_mean, float* state_var, float* mean,
                                               float* inv, float* scale, float* shift, int C, float decay) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = acc0[c] * inv_rows;
  const float mu = (float)m;
  const float var = (float)fmax(acc1[c] * inv_rows - m * m, 0.0);
  const float om = 1.f - decay;
  state_mean[c] = state_mean[c] - om * (state_mean[c] - mu);
  state_var[c] = state_var[c] - om * (state_var[c] - var);
  const float rs = 1.0f / sqrtf(var + kBnEps);
  const float sc = gamma[c] * rs;
  mean[c] = mu;
  inv[c] = rs;
  scale[c] = sc;
  shift[c] = beta[c] - mu * sc;
}

// batch statistics of Z[R,C] (two-pass, tf.nn.moments) or shadows -> scale/shift; EMA update
// one_pass (tensor-core modes of the materialised path): sum z and sum z^2 in ONE read of Z, both in double -- the
// variance sum z^2 / R - mean^2 then carries ~1e-16 E[z^2] / var of cancellation error, far below fp32 resolution; the
// two-pass form is what fp32 arithmetic needs (and what the fp32 mode keeps, as tf.nn.moments does).
// have_sums: the GEMM that produced Z already left the two sums in acc0 / acc1 (gemm_tc.cuh, A-stationary kernel).
static int bn_forward(const BnView& v, const float* Z, int R, bool training, float decay, cudaStream_t st,
                      bool one_pass = false, bool have_sums = false) {
  const int C = v.ch;
  const int tb = 128, nb = (C + tb - 1) / tb;
  if (training && one_pass) {
    if (!have_sums) {
      ColArgs a;
      a.Z = Z; a.ldz = C; a.R = R; a.C = C; a.acc0 = v.acc0; a.acc1 = v.acc1;
      AN3D_TRY(launch_col_reduce(a, COL_SUMSQ, st));
    }
    bn_finalize_sums_kernel<<<nb, tb, 0, st>>>(v.acc0, v.acc1, 1.0 / R, v.gamma, v.beta, v.state_mean, v.state_var, v.mean,
                                               v.inv, v.scale, v.shift, C, decay);
    AN3D_LAUNCH_CHECK();
    return AN3D_OK;
  }
  if (training) {
    ColArgs a;
    a.Z = Z; a.ldz = C; a.R = R; a.C = C; a.acc0 = v.acc0; a.acc1 = v.acc1; a.mean = v.mean;
    AN3D_TRY(launch_col_reduce(a, COL_SUM, st));
    bn_mean_kernel<<<nb, tb, 0, st>>>(v.acc0, v.mean, C, 1.0 / R);
    AN3D_LAUNCH_CHECK();
    AN3D_TRY(launch_col_reduce(a, COL_SQDIFF, st));
  }
  bn_finalize_kernel<<<nb, tb, 0, st>>>(v.acc1, 1.0 / R, v.gamma, v.beta, v.state_mean, v.state_var, v.mean, v.inv,
                                          v.scale, v.shift, C, training ? 1 : 0, decay);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

static int conv_stack_forward(const Model& m, const PlanF32& p, int s, int br, const float* params, float* state,
                              bool training, float decay, cudaStream_t st) {
  const int64_t M = p.M;
  const float* x = p.pin[s][br];
  const float *psc = nullptr, *psh = nullptr;
  for (size_t l = 0; l < m.conv[s].size(); ++l) {
    const Lin& L = m.conv[s][l];
    GemmArgs g;
    g.A = x; g.lda = L.cin; g.B = params + L.w; g.ldb = L.cout; g.C = p.z[s][l][br]; g.ldc = L.cout;
    g.M = (int)M; g.N = L.cout; g.K = L.cin; g.bias = params + L.b; g.pro_scale = psc; g.pro_shift = psh;
    BnView v = bn_view(m, p, params, state, false, br, L.bn);
    bool have_sums = false;
    const bool want_sums = training && p.tc_split > 0;
    AN3D_TRY(gemm_mat(p, g, false, false, st, want_sums ? v.acc0 : nullptr, want_sums ? v.acc1 : nullptr, &have_sums,
                      training ? p.tcx[s][l][br] : nullptr));
    AN3D_TRY(bn_forward(v, p.z[s][l][br], (int)M, training, decay, st, p.tc_split > 0, have_sums));
    x = p.z[s][l][br];
    psc = v.scale;
    psh = v.shift;
  }
  const Lin& L = m.conv[s].back();
  const int64_t ldg = s == EMB ? 2 * L.cout : L.cout;
  dim3 grid(p.B, (L.cout + 127) / 128);
  pool_kernel<<<grid, 128, 0, st>>>(x, L.cout, p.N, L.cout, psc, psh, p.g[s][br], ldg, p.gidx[s][br]);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

// bf16 mode: the weight images of every FC layer, once per forward (fc2_gemm.cuh); the backward reuses them
static int pack_fc_weights(const Model& m, const PlanF32& p, const float* params, cudaStream_t st) {
  for (int s = 0; s < 3; ++s)
    for (size_t l = 0; l < m.fc[s].size(); ++l) {
      const Lin& L = m.fc[s][l];
      fc2::PackArgs a;
      a.src = params + L.w; a.ld = L.cout; a.rows = L.cin; a.cols = L.cout; a.dst = p.fcw[s][l];
      AN3D_TRY(fc2::pack(a, st));
    }
  return AN3D_OK;
}

// get_mlp (models/tp8.py:75-82) for one branch (nbr = 1: the head, or the fp32 mode) or for BOTH siamese branches of
// stage s at once (bf16 mode: every layer is one launch of the GEMM with the two branches as a batch of two problems --
// same weights; their own input, BN prologue, statistics, mask, output).  x: [B, cin] with leading dim ldx, already
// activated.  In bf16 mode each layer is: pack the input (BN affine + ReLU + dropout of the producing layer applied on
// the way) into its bf16 image -> tcgen05 GEMM fed by bulk copies, BN column statistics fused into the epilogue.
static int mlp_forward_n(const Model& m, const PlanF32& p, int s, int nbr, int br0, const float* const x_in[2], int64_t ldx_in,
                         const float* params, float* state, bool training, float decay, const float* const mask[2],
                         cudaStream_t st) {
  const bool head = s == HEAD;
  const float* x[2] = {x_in[0], x_in[1]};
  int64_t ldx = ldx_in;
  const float *psc[2] = {nullptr, nullptr}, *psh[2] = {nullptr, nullptr};
  const size_t nl = m.fc[s].size();
  for (size_t l = 0; l < nl; ++l) {
    const Lin& L = m.fc[s][l];
    const bool fused_stats = p.bf16 && L.bn >= 0 && training;
    if (p.bf16) {
      fc2::PackArgs pa[2];
      fc2::Params f[2];
      for (int i = 0; i < nbr; ++i) {
        const int br = br0 + i;
        pa[i].src = x[i]; pa[i].ld = ldx; pa[i].rows = p.B; pa[i].cols = L.cin; pa[i].scale = psc[i]; pa[i].shift = psh[i];
        if (l == nl - 1 && training && mask[i]) {
          pa[i].mask = mask[i];
          pa[i].mask_scale = 1.0f / m.arch.keep_prob[s];
        }
        pa[i].dst = p.fcx[s][l][br];
        fc2::Params& q = f[i];
        q.A.g = p.fcx[s][l][br]; q.A.rows = p.B; q.A.cols = L.cin; q.a_mn = 0;
        q.B.g = p.fcw[s][l]; q.B.rows = L.cin; q.B.cols = L.cout; q.b_mn = 1;
        q.C = p.fz[s][l][br]; q.ldc = L.cout; q.M = p.B; q.N = L.cout; q.K = L.cin; q.bias = params + L.b;
        if (fused_stats) {   // column sums of z and z^2 come out of the GEMM epilogue
          BnView v = bn_view(m, p, params, state, head, br, L.bn);
          q.stat_sum = v.acc0;
          q.stat_sq = v.acc1;
        }
      }
      AN3D_TRY(fc2::pack(pa[0], st, nbr == 2 ? &pa[1] : nullptr));
      // few output tiles (inference batches): split K so that the launch fills the SMs.  Never in training mode and
      // not under AN3D_DETERMINISTIC: the K slices meet in fp32 reductions, whose order is not reproducible.
      if (!fused_stats && !training && !p.deterministic) {
        const int tiles = nbr * ((p.B + 127) / 128) * ((L.cout + 127) / 128);
        const int ks = std::min(L.cin / 128, 148 / tiles);
        if (ks > 1) {
          for (int i = 0; i < nbr; ++i) {
            f[i].ksplit = ks;
            AN3D_CUDA_CHECK(cudaMemsetAsync(f[i].C, 0, sizeof(float) * (size_t)p.B * L.cout, st));
          }
        }
      }
      AN3D_TRY(fc2::launch(f[0], st, nbr == 2 ? &f[1] : nullptr));
    } else {
      GemmArgs g;
      g.A = x[0]; g.lda = ldx; g.B = params + L.w; g.ldb = L.cout; g.C = p.fz[s][l][br0]; g.ldc = L.cout;
      g.M = p.B; g.N = L.cout; g.K = L.cin; g.bias = params + L.b; g.pro_scale = psc[0]; g.pro_shift = psh[0];
      if (l == nl - 1 && training && mask[0]) {
        g.pro_mask = mask[0];
        g.pro_mask_scale = 1.0f / m.arch.keep_prob[s];
      }
      AN3D_TRY(gemm_mat(p, g, false, false, st));
    }
    for (int i = 0; i < nbr; ++i) {
      const int br = br0 + i;
      if (L.bn >= 0) {
        BnView v = bn_view(m, p, params, state, head, br, L.bn);
        if (fused_stats) {
          bn_finalize_sums_kernel<<<(v.ch + 127) / 128, 128, 0, st>>>(v.acc0, v.acc1, 1.0 / p.B, v.gamma, v.beta, v.state_mean,
                                                                      v.state_var, v.mean, v.inv, v.scale, v.shift, v.ch, decay);
          AN3D_LAUNCH_CHECK();
        } else if (!p.prepared) {
          AN3D_TRY(bn_forward(v, p.fz[s][l][br], p.B, training, decay, st));
        }
        psc[i] = v.scale;
        psh[i] = v.shift;
      }
      x[i] = p.fz[s][l][br];
    }
    ldx = L.cout;
  }
  return AN3D_OK;
}

static int mlp_forward(const Model& m, const PlanF32& p, int s, int br, const float* x, int64_t ldx, const float* params,
                       float* state, bool training, float decay, const float* mask, cudaStream_t st) {
  const float* xs[2] = {x, nullptr};
  const float* ms[2] = {mask, nullptr};
  return mlp_forward_n(m, p, s, 1, br, xs, ldx, params, state, training, decay, ms, st);
}

static int mlp_forward_pair(const Model& m, const PlanF32& p, int s, const float* const x_in[2], int64_t ldx_in,
                            const float* params, float* state, bool training, float decay, const float* const mask[2],
                            cudaStream_t st) {
  return mlp_forward_n(m, p, s, 2, 0, x_in, ldx_in, params, state, training, decay, mask, st);
}

int forward_impl(const Model& m, const float* params, float* state, const float* pcs1, const float* pcs2, int B, int N,
                int flags, float bn_decay, const an3d_dropout* dropout, const an3d_outputs* out, void* workspace,
                int64_t workspace_bytes, cudaStream_t st) {
  PlanF32 p;
  AN3D_TRY(plan_f32(m, B, N, flags, workspace, &p));
  if (p.bytes > workspace_bytes) {
    set_error("workspace too small: need %lld bytes, got %lld", (long long)p.bytes, (long long)workspace_bytes);
    return AN3D_ERR_WORKSPACE;
  }
  const bool training = (flags & AN3D_TRAINING) != 0;
  const bool bf16 = p.bf16;   // the fused kernels; every other mode walks the materialised path below
  p.prepared = bf16 && !training && (flags & AN3D_WEIGHTS_PREPARED) != 0;
  p.deterministic = (flags & AN3D_DETERMINISTIC) != 0;
  if (bf16 && !p.prepared) {
    AN3D_TRY(pack_weights_bf16(m, p, params, st));
    AN3D_TRY(pack_fc_weights(m, p, params, st));
  }
  const int nb = m.nb;
  const int64_t M = p.M;
  if (!p.prepared) {   // (the reduction scratch only feeds batch statistics and the eval-mode finalize kernels)
    AN3D_CUDA_CHECK(cudaMemsetAsync(p.bn.acc0, 0, sizeof(double) * m.bn_total_ch(), st));
    AN3D_CUDA_CHECK(cudaMemsetAsync(p.bn.acc1, 0, sizeof(double) * m.bn_total_ch(), st));
  }
  const float* masks[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  if (training) {
    for (int i = 0; i < 5; ++i) {
      const int s = i < 2 ? S1 : (i < 4 ? S2 : HEAD);
      if (m.arch.keep_prob[s] >= 1.f) continue;
      const int width = m.fc[s][m.fc[s].size() - 2].cout;
      const int64_t cnt = (int64_t)B * width;
      if (dropout && dropout->masks[i]) {
        AN3D_CUDA_CHECK(cudaMemcpyAsync(p.mask[i], dropout->masks[i], cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
      } else {
        const uint64_t seed = dropout ? dropout->seed : 0;
        dropout_mask_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(p.mask[i], cnt, seed, (uint32_t)i,
                                                                          m.arch.keep_prob[s], dropout ? dropout->seed_dev : nullptr);
        AN3D_LAUNCH_CHECK();
      }
      masks[i] = p.mask[i];
    }
  }
  const float* pcs[2] = {pcs1, pcs2};
  float* c1[2] = {out->pred_s1_pc1centers, out->pred_s1_pc2centers};
  float* c2[2] = {out->pred_s2_pc1centers, out->pred_s2_pc2centers};
  float* lg[2] = {out->pred_pc1angle_logits, out->pred_pc2angle_logits};
  const unsigned pt_blocks = (unsigned)((M + 255) / 256);
  const unsigned w_blocks = (unsigned)((B + 3) / 4);   // one warp per sample
  if (bf16) {
    // stage-major order: the two branches of a stage are independent, so their FC layers share launches
    // (mlp_forward_pair) and every under-filled GEMM gets twice the CTAs
    for (int br = 0; br < 2; ++br) {
      centroid_kernel<<<(B + 3) / 4, 128, 0, st>>>(pcs[br], N, p.mu[br], B);
      AN3D_LAUNCH_CHECK();
    }
    const float* mk1[2] = {masks[0], masks[1]};
    const float* mk2[2] = {masks[2], masks[3]};
    SideStream* ss = two_streams_enabled() ? side_stream() : nullptr;
    const float* const ctr_mu[2] = {p.mu[0], p.mu[1]};
    const float* const ctr_c1[2] = {c1[0], c1[1]};
    const float* const ctr_c2[2] = {c2[0], c2[1]};
    const float* const ang2[2] = {p.ang[0], p.ang[1]};
    // stage 1 (tp8.py:106-109)
    if (ss) AN3D_TRY(conv_stage_two_streams(ss, m, p, S1, pcs, ctr_mu, nullptr, params, state, training, bn_decay, st));
    else
    for (int br = 0; br < 2; ++br)
      AN3D_TRY(conv_stack_forward_bf16(m, p, S1, br, pcs[br], p.mu[br], nullptr, params, state, training, bn_decay, st));
    const float* g1[2] = {p.g[S1][0], p.g[S1][1]};
    AN3D_TRY(mlp_forward_pair(m, p, S1, g1, m.conv[S1].back().cout, params, state, training, bn_decay, mk1, st));
    for (int br = 0; br < 2; ++br) {
      post_s1_kernel<<<(B * 3 + 127) / 128, 128, 0, st>>>(p.fz[S1][m.fc[S1].size() - 1][br], p.mu[br], c1[br], B);
      AN3D_LAUNCH_CHECK();
    }
    // stage 2 (tp8.py:113-118)
    if (ss) AN3D_TRY(conv_stage_two_streams(ss, m, p, S2, pcs, ctr_c1, nullptr, params, state, training, bn_decay, st));
    else
    for (int br = 0; br < 2; ++br)
      AN3D_TRY(conv_stack_forward_bf16(m, p, S2, br, pcs[br], c1[br], nullptr, params, state, training, bn_decay, st));
    const float* g2[2] = {p.g[S2][0], p.g[S2][1]};
    AN3D_TRY(mlp_forward_pair(m, p, S2, g2, m.conv[S2].back().cout, params, state, training, bn_decay, mk2, st));
    for (int br = 0; br < 2; ++br) {
      post_s2_kernel<<<w_blocks, 128, 0, st>>>(p.fz[S2][m.fc[S2].size() - 1][br], c1[br], c2[br], lg[br], p.ang[br],
                                               p.angk[br], B, nb);
      AN3D_LAUNCH_CHECK();
    }
    // canonicalise + final embedding (tp8.py:122-130)
    if (ss) AN3D_TRY(conv_stage_two_streams(ss, m, p, EMB, pcs, ctr_c2, ang2, params, state, training, bn_decay, st));
    else
    for (int br = 0; br < 2; ++br)
      AN3D_TRY(conv_stack_forward_bf16(m, p, EMB, br, pcs[br], c2[br], p.ang[br], params, state, training, bn_decay, st));
  } else {
  for (int br = 0; br < 2; ++br) {
    centroid_kernel<<<(B + 3) / 4, 128, 0, st>>>(pcs[br], N, p.mu[br], B);
    AN3D_LAUNCH_CHECK();
    // stage 1 (tp8.py:106-109)
    stage_input_kernel<<<pt_blocks, 256, 0, st>>>(pcs[br], p.mu[br], nullptr, p.pin[S1][br], N, M);
    AN3D_LAUNCH_CHECK();
    AN3D_TRY(conv_stack_forward(m, p, S1, br, params, state, training, bn_decay, st));
    AN3D_TRY(mlp_forward(m, p, S1, br, p.g[S1][br], m.conv[S1].back().cout, params, state, training, bn_decay,
                         masks[br], st));
    post_s1_kernel<<<(B * 3 + 127) / 128, 128, 0, st>>>(p.fz[S1][m.fc[S1].size() - 1][br], p.mu[br], c1[br], B);
    AN3D_LAUNCH_CHECK();
    // stage 2 (tp8.py:113-118)
    stage_input_kernel<<<pt_blocks, 256, 0, st>>>(pcs[br], c1[br], nullptr, p.pin[S2][br], N, M);
    AN3D_LAUNCH_CHECK();
    AN3D_TRY(conv_stack_forward(m, p, S2, br, params, state, training, bn_decay, st));
    AN3D_TRY(mlp_forward(m, p, S2, br, p.g[S2][br], m.conv[S2].back().cout, params, state, training, bn_decay,
                         masks[2 + br], st));
    post_s2_kernel<<<w_blocks, 128, 0, st>>>(p.fz[S2][m.fc[S2].size() - 1][br], c1[br], c2[br], lg[br], p.ang[br],
                                             p.angk[br], B, nb);
    AN3D_LAUNCH_CHECK();
    // canonicalise + final embedding (tp8.py:122-130)
    stage_input_kernel<<<pt_blocks, 256, 0, st>>>(pcs[br], c2[br], p.ang[br], p.pin[EMB][br], N, M);
    AN3D_LAUNCH_CHECK();
    AN3D_TRY(conv_stack_forward(m, p, EMB, br, params, state, training, bn_decay, st));
  }
  }
  // head (tp8.py:144-156)
  AN3D_TRY(mlp_forward(m, p, HEAD, 0, p.feat, 2 * m.conv[EMB].back().cout, params, state, training, bn_decay, masks[4],
                       st));
  post_head_kernel<<<w_blocks, 128, 0, st>>>(p.fz[HEAD][m.fc[HEAD].size() - 1][0], c2[0], c2[1], out->pred_translations,
                                             out->pred_remaining_angle_logits, B, nb);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

}  // namespace an3d
