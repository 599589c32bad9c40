// fp32 parity path: backward of get_model + get_loss (what optimizer.minimize differentiates,
// train.py:217).  Every activation needed was materialised by forward_f32 in the workspace.
#include <algorithm>

#include "kernels_f32.cuh"
#include "loss.cuh"
#include "bf16_path.cuh"
#include "fc2_gemm.cuh"
#include "gemm_tc.cuh"

namespace an3d {

namespace {

struct BnRef {
  const float *scale, *shift, *mean, *inv;
  double *acc0, *acc1;
  float *dgamma, *dbeta;
  int ch;
};

BnRef bn_ref(const Model& m, const PlanF32& p, float* grads, bool head, int br, int bn) {
  BnRef v;
  const int ch = head ? m.bn_head[bn].ch : m.bn_branch[bn].ch;
  const int64_t po = m.bn_param_off(head, br, bn), sl = m.bn_slot_off(head, br, bn);
  v.scale = p.bn.scale + sl;
  v.shift = p.bn.shift + sl;
  v.mean = p.bn.mean + sl;
  v.inv = p.bn.inv + sl;
  v.acc0 = p.bn.acc0 + sl;
  v.acc1 = p.bn.acc1 + sl;
  v.dgamma = grads + po;
  v.dbeta = grads + po + ch;
  v.ch = ch;
  return v;
}

// bn_bwd_apply_kernel for the bf16 FC path: dZ goes straight into the bf16 block image the wgrad / dgrad GEMMs read
// (fc2_gemm.cuh) -- one thread per (row, 8-column chunk), padding written as zeros -- instead of to an fp32 matrix that
// a pack kernel would then re-read.
static __global__ void __launch_bounds__(256) bn_bwd_apply_img_kernel(ColArgs a, __nv_bfloat16* img, double inv_rows) {
  const int rows_pad = (a.R + 127) & ~127, chunks = ((a.C + 127) & ~127) >> 3;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows_pad * chunks) return;
  const int c8 = (int)(i / rows_pad), r = (int)(i - (int64_t)c8 * rows_pad);
  const int c0 = c8 * 8;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  if (r < a.R && c0 < a.C) {
    const int n = min(8, a.C - c0);
    float z[8], dy[8], mk[8];
    const float* zp = a.Z + (int64_t)r * a.ldz + c0;
    const float* dp = a.dA + (int64_t)r * a.ldd + c0;
    const float* mp = a.mask ? a.mask + (int64_t)r * a.ldd + c0 : nullptr;
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    // lanes walk consecutive ROWS (the image's fast axis), so each lane owns one 32-byte sector of its row: read it with
    // two 16-byte loads -- eight scalar loads per array made this kernel LSU-bound (22 us per launch)
    if (n == 8 && al(zp) && al(dp) && (!mp || al(mp))) {
      const float4 z0 = *reinterpret_cast<const float4*>(zp), z1 = *reinterpret_cast<const float4*>(zp + 4);
      const float4 d0 = *reinterpret_cast<const float4*>(dp), d1 = *reinterpret_cast<const float4*>(dp + 4);
      z[0] = z0.x; z[1] = z0.y; z[2] = z0.z; z[3] = z0.w; z[4] = z1.x; z[5] = z1.y; z[6] = z1.z; z[7] = z1.w;
      dy[0] = d0.x; dy[1] = d0.y; dy[2] = d0.z; dy[3] = d0.w; dy[4] = d1.x; dy[5] = d1.y; dy[6] = d1.z; dy[7] = d1.w;
      if (mp) {
        const float4 m0 = *reinterpret_cast<const float4*>(mp), m1 = *reinterpret_cast<const float4*>(mp + 4);
        mk[0] = m0.x; mk[1] = m0.y; mk[2] = m0.z; mk[3] = m0.w; mk[4] = m1.x; mk[5] = m1.y; mk[6] = m1.z; mk[7] = m1.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        z[e] = e < n ? zp[e] : 0.f;
        dy[e] = e < n ? dp[e] : 0.f;
        mk[e] = (mp && e < n) ? mp[e] : 0.f;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (e < n) {
        const int c = c0 + e;
        float g = dy[e];
        if (mp) g *= mk[e] * a.mask_scale;
        const float sc = a.scale[c];
        if (!(fmaf(z[e], sc, a.shift[c]) > 0.f)) g = 0.f;
        const float xhat = (z[e] - a.mean[c]) * a.inv[c];
        const float m0 = (float)(a.acc0[c] * inv_rows), m1 = (float)(a.acc1[c] * inv_rows);
        v[e] = sc * (g - m0 - xhat * m1);
      }
    }
  }
  __nv_bfloat162 b0 = __floats2bfloat162_rn(v[0], v[1]), b1 = __floats2bfloat162_rn(v[2], v[3]),
                 b2 = __floats2bfloat162_rn(v[4], v[5]), b3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
  o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
  const int64_t blk = (int64_t)(r >> 7) * (chunks >> 4) + (c8 >> 4);
  *reinterpret_cast<uint4*>(img + blk * 16384 + ((c8 & 15) * 128 + (r & 127)) * 8) = o;
}

// dA (grad wrt post-activation, post-dropout) -> dZ (grad wrt pre-BN), written to dZ (may alias dA) -- or, with `img`,
// to the bf16 block image of dZ only.
int bn_relu_backward(const BnRef& v, const float* Z, int R, const float* dA, int64_t ldd, const float* mask,
                     float mask_scale, float* dZ, int64_t ldo, cudaStream_t st, __nv_bfloat16* img = nullptr) {
  ColArgs a;
  a.Z = Z; a.ldz = v.ch; a.R = R; a.C = v.ch; a.mean = v.mean; a.inv = v.inv; a.scale = v.scale; a.shift = v.shift;
  a.dA = dA; a.ldd = ldd; a.mask = mask; a.mask_scale = mask_scale; a.acc0 = v.acc0; a.acc1 = v.acc1;
  AN3D_TRY(launch_col_reduce(a, COL_DY, st));
  if (img) {
    const int64_t total = (int64_t)((R + 127) & ~127) * (((v.ch + 127) & ~127) >> 3);
    bn_bwd_apply_img_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, img, 1.0 / R);
  } else {
    const int64_t total = (int64_t)R * v.ch;
    bn_bwd_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, dZ, ldo, 1.0 / R);
  }
  AN3D_LAUNCH_CHECK();
  bn_bwd_params_kernel<<<(v.ch + 127) / 128, 128, 0, st>>>(v.acc0, v.acc1, v.dgamma, v.dbeta, v.ch);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

// ---- BN + ReLU backward of a conv layer in the tensor-core modes of the materialised path (gemm_tc.cuh) ----
// dZ = scale * (dy - sum_dy / R - xhat * sum_dy_xhat / R) goes straight into the split bf16 images the wgrad / dgrad
// GEMMs read (no fp32 dZ matrix, no pack pass), or -- for a layer whose GEMMs stay on the CUDA cores -- to fp32.
// dy = ReLU mask * dA with dA either dense [R, C], or (POOLED: the last conv layer, whose output feeds the max-pool)
// dG[b, c] at row b*N + idx[b, c] and zero elsewhere: the dense [R, C] gradient of the pooled activation -- a memset,
// a scatter and two reads of R*C floats -- never exists.
// Thread mapping of the element-wise kernels: a warp covers 8 rows x 4 chunks of 8 columns, so every global access is a
// full 128-byte line on both sides (row-major fp32: 4 chunks = 128 B per row; image plane: 8 rows x 16 B = 128 B).
struct SplitBwdArgs {
  ColArgs c;                    // Z, mean, inv, scale, shift, acc0, acc1 (+ dA / ldd when dense)
  const float* dG = nullptr;    // POOLED: [B, C] gradient of the pooled features, leading dim ldg
  int64_t ldg = 0;
  const int32_t* idx = nullptr; // POOLED: [B, C] arg rows of the pool
  int N = 0;                    // POOLED: points per cloud
  const float* cb = nullptr;    // [C]  -scale * sum_dy / R
  const float* cc = nullptr;    // [C]  -scale * inv * sum_dy_xhat / R
  __nv_bfloat16* img[3] = {nullptr, nullptr, nullptr};
  int nsplit = 0;               // 0: fp32 output to dZ
  float* dZ = nullptr;
  int64_t ldo = 0;
  double inv_rows = 0.0;
};

// sum_dy / sum_dy*xhat of the pooled layer: one thread per (cloud, channel), a block reduces 8 clouds x 32 channels
static __global__ void __launch_bounds__(256) col_dy_pooled_kernel(SplitBwdArgs a, int B) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s0 = 0.0, s1 = 0.0;
  if (c < a.c.C) {
    const float mean = a.c.mean[c], inv = a.c.inv[c], sc = a.c.scale[c], sh = a.c.shift[c];
    for (int b = blockIdx.y * 64 + threadIdx.y; b < min(B, (int)blockIdx.y * 64 + 64); b += 8) {
      const int64_t r = (int64_t)b * a.N + a.idx[(int64_t)b * a.c.C + c];
      const float z = a.c.Z[r * a.c.ldz + c];
      const float g = fmaf(z, sc, sh) > 0.f ? a.dG[(int64_t)b * a.ldg + c] : 0.f;
      s0 += (double)g;
      s1 += (double)g * (double)((z - mean) * inv);
    }
  }
  __shared__ double sm0[8][33], sm1[8][33];
  sm0[threadIdx.y][threadIdx.x] = s0;
  sm1[threadIdx.y][threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.y == 0 && c < a.c.C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) { s0 += sm0[i][threadIdx.x]; s1 += sm1[i][threadIdx.x]; }
    atomicAdd(a.c.acc0 + c, s0);
    atomicAdd(a.c.acc1 + c, s1);
  }
}

// per-channel constants of dZ = scale * dy + cb + cc * (z - mean)
static __global__ void bn_bwd_coef_kernel(const double* acc0, const double* acc1, double inv_rows, const float* scale,
                                          const float* inv, float* cb, float* cc, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float m0 = (float)(acc0[c] * inv_rows), m1 = (float)(acc1[c] * inv_rows);
  cb[c] = -scale[c] * m0;
  cc[c] = -scale[c] * inv[c] * m1;
}

__device__ __forceinline__ void load8(const float* p, int n, float (&v)[8]) {
  if (n == 8 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = e < n ? p[e] : 0.f;
  }
}

template <bool POOLED>
static __global__ void __launch_bounds__(256) bn_bwd_apply_split_kernel(SplitBwdArgs a) {
  // a warp covers 8 U rows x 4 chunks of 8 columns; a lane owns U rows (r, r + 8) of one chunk.  Dense form, U = 2: both
  // rows' loads are issued before either is consumed and the per-channel constants are fetched once (266 vs 299 us on
  // [819200, 128]: 5.5 TB/s).  Pooled form, U = 1: a second row doubles the per-cloud arg / gradient gathers and was
  // slower (2.37 vs 2.22 ms on [819200, 1024]).
  constexpr int U = POOLED ? 1 : 2;
  const int rows_pad = a.nsplit ? (a.c.R + 127) & ~127 : (a.c.R + 8 * U - 1) & ~(8 * U - 1);
  const int chunks = a.nsplit ? ((a.c.C + 127) & ~127) >> 3 : (a.c.C + 7) >> 3;
  const int cgroups = (chunks + 3) >> 2;                       // groups of 4 chunks (32 columns)
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // warp: (row group of 8 U, chunk group), chunk groups fastest
  const int lane = threadIdx.x & 31;
  const int64_t rg = w / cgroups;
  const int r0 = (int)(rg * 8 * U) + (lane & 7), c8 = (int)(w - rg * cgroups) * 4 + (lane >> 3), c0 = c8 * 8;
  if (rg * 8 * U >= rows_pad || c8 >= chunks) return;
  const bool colok = c0 < a.c.C;
  const int n = colok ? min(8, a.c.C - c0) : 0;
  float z[U][8], dy[U][8];
  bool ok[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int r = r0 + 8 * u;
    ok[u] = colok && r < a.c.R;
    if (ok[u]) load8(a.c.Z + (int64_t)r * a.c.ldz + c0, n, z[u]);
  }
  if (POOLED) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ok[u]) {
        const int r = r0 + 8 * u;
        const int b = r / a.N, pt = r - b * a.N;
        const int32_t* ip = a.idx + (int64_t)b * a.c.C + c0;
        float g[8];
        load8(a.dG + (int64_t)b * a.ldg + c0, n, g);
        if (n == 8 && (reinterpret_cast<uintptr_t>(ip) & 15) == 0) {
          const int4 i0 = *reinterpret_cast<const int4*>(ip), i1 = *reinterpret_cast<const int4*>(ip + 4);
          const int ix[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) dy[u][e] = ix[e] == pt ? g[e] : 0.f;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) dy[u][e] = (e < n && ip[e] == pt) ? g[e] : 0.f;
        }
      }
    }
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (ok[u]) load8(a.c.dA + (int64_t)(r0 + 8 * u) * a.c.ldd + c0, n, dy[u]);
  }
  float sc[8], sh[8], mean[8], cb[8], cc[8];
  if (colok) {
    load8(a.c.scale + c0, n, sc);
    load8(a.c.shift + c0, n, sh);
    load8(a.c.mean + c0, n, mean);
    load8(a.cb + c0, n, cb);
    load8(a.cc + c0, n, cc);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int r = r0 + 8 * u;
    if (r >= rows_pad) continue;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (ok[u]) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (e < n) {
          const float g = fmaf(z[u][e], sc[e], sh[e]) > 0.f ? dy[u][e] : 0.f;
          v[e] = fmaf(sc[e], g, fmaf(cc[e], z[u][e] - mean[e], cb[e]));
        }
      }
    }
    if (a.nsplit == 0) {
      if (ok[u]) {
        float* op = a.dZ + (int64_t)r * a.ldo + c0;
        if (n == 8 && (reinterpret_cast<uintptr_t>(op) & 15) == 0) {
          *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (e < n) op[e] = v[e];
        }
      }
      continue;
    }
    const int64_t off = ((int64_t)(r >> 7) * (chunks >> 4) + (c8 >> 4)) * 16384 + ((c8 & 15) * 128 + (r & 127)) * 8;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      if (s < a.nsplit) {
        __nv_bfloat162 b[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          b[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
          v[2 * e] -= __bfloat162float(b[e].x);
          v[2 * e + 1] -= __bfloat162float(b[e].y);
        }
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&b[0]); o.y = *reinterpret_cast<uint32_t*>(&b[1]);
        o.z = *reinterpret_cast<uint32_t*>(&b[2]); o.w = *reinterpret_cast<uint32_t*>(&b[3]);
        *reinterpret_cast<uint4*>(a.img[s] + off) = o;
      }
    }
  }
}

// Image-writing form of the kernel above with DRAM-friendly access on both sides: a CTA owns one 128 x 128 image block.
// It reads its 128 rows in 512-byte runs (a warp = 2 rows x 16 chunks), and the 16-byte image pieces -- which the lane
// mapping above scattered as 128-byte lines over 16 planes 2 KB apart, each read and each write landing in a different
// DRAM page at 1024 channels (3.8 TB/s) -- are staged in shared memory in block layout (plane stride padded by one chunk
// against bank conflicts) and leave as one contiguous 32 KB run per image.
constexpr int kBlkPlane = 129;                                   // 16-byte units per staged plane
constexpr size_t kBlkImgBytes = 16 * kBlkPlane * 16;             // one staged image block
template <bool POOLED>
static __global__ void __launch_bounds__(256, 2) bn_bwd_apply_block_kernel(SplitBwdArgs a) {
  extern __shared__ __align__(16) uint8_t blk_smem[];
  const int ncb = ((a.c.C + 127) & ~127) >> 7;
  const int rb = blockIdx.x / ncb, cb = blockIdx.x - rb * ncb;
  const int chunk = threadIdx.x & 15, c8 = cb * 16 + chunk, c0 = c8 * 8;
  const bool colok = c0 < a.c.C;
  const int n = colok ? min(8, a.c.C - c0) : 0;
  float sc[8], sh[8], mean[8], cbv[8], ccv[8];
  if (colok) {
    load8(a.c.scale + c0, n, sc);
    load8(a.c.shift + c0, n, sh);
    load8(a.c.mean + c0, n, mean);
    load8(a.cb + c0, n, cbv);
    load8(a.cc + c0, n, ccv);
  }
#pragma unroll 1
  for (int it0 = 0; it0 < 8; it0 += 4) {
    float z[4][8], dy[4][8];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = rb * 128 + (it0 + u) * 16 + (threadIdx.x >> 4);
      ok[u] = colok && r < a.c.R;
      if (ok[u]) load8(a.c.Z + (int64_t)r * a.c.ldz + c0, n, z[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const int r = rb * 128 + (it0 + u) * 16 + (threadIdx.x >> 4);
      if (POOLED) {
        const int b = r / a.N, pt = r - b * a.N;
        const int32_t* ip = a.idx + (int64_t)b * a.c.C + c0;
        float g[8];
        load8(a.dG + (int64_t)b * a.ldg + c0, n, g);
        if (n == 8 && (reinterpret_cast<uintptr_t>(ip) & 15) == 0) {
          const int4 i0 = *reinterpret_cast<const int4*>(ip), i1 = *reinterpret_cast<const int4*>(ip + 4);
          const int ix[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) dy[u][e] = ix[e] == pt ? g[e] : 0.f;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) dy[u][e] = (e < n && ip[e] == pt) ? g[e] : 0.f;
        }
      } else {
        load8(a.c.dA + (int64_t)r * a.c.ldd + c0, n, dy[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = (it0 + u) * 16 + (threadIdx.x >> 4);
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (ok[u]) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (e < n) {
            const float g = fmaf(z[u][e], sc[e], sh[e]) > 0.f ? dy[u][e] : 0.f;
            v[e] = fmaf(sc[e], g, fmaf(ccv[e], z[u][e] - mean[e], cbv[e]));
          }
        }
      }
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        if (s < a.nsplit) {
          __nv_bfloat162 b[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            b[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            v[2 * e] -= __bfloat162float(b[e].x);
            v[2 * e + 1] -= __bfloat162float(b[e].y);
          }
          uint4 o;
          o.x = *reinterpret_cast<uint32_t*>(&b[0]); o.y = *reinterpret_cast<uint32_t*>(&b[1]);
          o.z = *reinterpret_cast<uint32_t*>(&b[2]); o.w = *reinterpret_cast<uint32_t*>(&b[3]);
          *reinterpret_cast<uint4*>(blk_smem + s * kBlkImgBytes + (size_t)(chunk * kBlkPlane + row) * 16) = o;
        }
      }
    }
  }
  __syncthreads();
  const int64_t blk_off = ((int64_t)rb * ncb + cb) * 16384;      // elements
  for (int s = 0; s < a.nsplit; ++s) {
    uint4* dst = reinterpret_cast<uint4*>(a.img[s] + blk_off);
    const uint8_t* src = blk_smem + s * kBlkImgBytes;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int u16 = k * 256 + threadIdx.x;                      // 16-byte unit of the block: plane = u16 / 128, row = u16 % 128
      dst[u16] = *reinterpret_cast<const uint4*>(src + (size_t)((u16 >> 7) * kBlkPlane + (u16 & 127)) * 16);
    }
  }
}

// One conv layer's BN + ReLU backward in the tensor-core modes.  pooled: dG / idx describe the (sparse) gradient of the
// layer's activation, else dA [R, C] is dense.  to_img: dZ lands as p.tc_split images in the gradient slot (*dz_img
// describes them), else as fp32 in dZ (leading dim C; may alias dA).
static int bn_relu_backward_split(const PlanF32& p, const BnRef& v, const float* Z, int R, bool pooled, const float* dA,
                                  const float* dG, int64_t ldg, const int32_t* idx, bool to_img, tcg::SplitMat* dz_img,
                                  float* dZ, cudaStream_t st) {
  SplitBwdArgs a;
  a.c.Z = Z; a.c.ldz = v.ch; a.c.R = R; a.c.C = v.ch; a.c.mean = v.mean; a.c.inv = v.inv; a.c.scale = v.scale; a.c.shift = v.shift;
  a.c.dA = dA; a.c.ldd = v.ch; a.c.acc0 = v.acc0; a.c.acc1 = v.acc1;
  a.dG = dG; a.ldg = ldg; a.idx = idx; a.N = p.N; a.inv_rows = 1.0 / R;
  if (pooled) {
    dim3 grid((v.ch + 31) / 32, (p.B + 63) / 64), block(32, 8);
    col_dy_pooled_kernel<<<grid, block, 0, st>>>(a, p.B);
    AN3D_LAUNCH_CHECK();
  } else {
    AN3D_TRY(launch_col_reduce(a.c, COL_DY, st));
  }
  if (!p.bwd_coef) {
    set_error("bn_relu_backward_split: no coefficient scratch in this plan");
    return AN3D_ERR_WORKSPACE;
  }
  float* cb = p.bwd_coef;
  float* cc = p.bwd_coef + ((v.ch + 3) & ~3);
  bn_bwd_coef_kernel<<<(v.ch + 127) / 128, 128, 0, st>>>(v.acc0, v.acc1, 1.0 / R, v.scale, v.inv, cb, cc, v.ch);
  AN3D_LAUNCH_CHECK();
  a.cb = cb; a.cc = cc;
  int64_t warps;
  const int rows_per_warp = pooled ? 8 : 16;     // (bn_bwd_apply_split_kernel's U)
  if (to_img) {
    const int64_t elems = fc_image_elems(R, v.ch);
    if (elems * p.tc_split > p.tcbuf_elems[tcg::SLOT_DZ] || !p.tcbuf[tcg::SLOT_DZ]) {
      set_error("bn_relu_backward_split: %d x %d does not fit the gradient image scratch", R, v.ch);
      return AN3D_ERR_WORKSPACE;
    }
    a.nsplit = p.tc_split;
    dz_img->n = p.tc_split;
    for (int s = 0; s < p.tc_split; ++s) {
      a.img[s] = p.tcbuf[tcg::SLOT_DZ] + s * elems;
      dz_img->img[s].g = a.img[s]; dz_img->img[s].rows = R; dz_img->img[s].cols = v.ch;
    }
    warps = (int64_t)(((R + 127) & ~127) / rows_per_warp) * ((((v.ch + 127) & ~127) >> 3) / 4);
  } else {
    a.nsplit = 0; a.dZ = dZ; a.ldo = v.ch;
    warps = (int64_t)((R + rows_per_warp - 1) / rows_per_warp) * ((((v.ch + 7) >> 3) + 3) / 4);
  }
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  if (to_img) {
    // one CTA per 128 x 128 image block, staged in shared memory
    static bool attr_set = false;
    if (!attr_set) {
      AN3D_CUDA_CHECK(cudaFuncSetAttribute(bn_bwd_apply_block_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * kBlkImgBytes)));
      AN3D_CUDA_CHECK(cudaFuncSetAttribute(bn_bwd_apply_block_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * kBlkImgBytes)));
      attr_set = true;
    }
    const unsigned nblk = (unsigned)(((R + 127) >> 7) * ((v.ch + 127) >> 7));
    const size_t smem = (size_t)p.tc_split * kBlkImgBytes;
    if (pooled) bn_bwd_apply_block_kernel<true><<<nblk, 256, smem, st>>>(a);
    else bn_bwd_apply_block_kernel<false><<<nblk, 256, smem, st>>>(a);
  } else if (pooled) {
    bn_bwd_apply_split_kernel<true><<<blocks, 256, 0, st>>>(a);
  } else {
    bn_bwd_apply_split_kernel<false><<<blocks, 256, 0, st>>>(a);
  }
  AN3D_LAUNCH_CHECK();
  bn_bwd_params_kernel<<<(v.ch + 127) / 128, 128, 0, st>>>(v.acc0, v.acc1, v.dgamma, v.dbeta, v.ch);
  AN3D_LAUNCH_CHECK();
  return AN3D_OK;
}

// bf16 mode: the operands of an FC layer's backward GEMMs as plane-major bf16 images (fc2_gemm.cuh): the layer's input
// activations and weights were packed by the forward pass, the gradient image is packed here.
struct FcImages {
  const __nv_bfloat16* x = nullptr;     // [R, cin]  input activations (forward)
  const __nv_bfloat16* w = nullptr;     // [cin, cout] weights (forward)
  __nv_bfloat16* dz = nullptr;          // [R, cout] scratch for the gradient image
  bool dz_packed = false;               // the image already holds dZ (written by the BN backward)
};

// wgrad (grads.W += X^T dZ, split-K with reductions) and dgrad (dX = dZ W^T) of one FC layer on the tensor cores, for
// one branch or for the two siamese branches as one batched launch each.
static int fc_backward_bf16(const Lin& L, int R, int nbr, const FcImages img[2], const float* const dZ[2], float* gradW,
                            float* const dX[2], int64_t lddx, cudaStream_t st, bool dz_packed = false) {
  fc2::PackArgs pa[2];
  fc2::Params w[2], d[2];
  for (int i = 0; i < nbr; ++i) {
    pa[i].src = dZ[i]; pa[i].ld = L.cout; pa[i].rows = R; pa[i].cols = L.cout; pa[i].dst = img[i].dz;
    fc2::Params& f = w[i];
    f.A.g = img[i].x; f.A.rows = R; f.A.cols = L.cin; f.a_mn = 1;          // contraction over the batch rows
    f.B.g = img[i].dz; f.B.rows = R; f.B.cols = L.cout; f.b_mn = 1;
    f.C = gradW; f.ldc = L.cout; f.M = L.cin; f.N = L.cout; f.K = R; f.accumulate = 1;
    const int tiles = nbr * ((L.cin + 127) / 128) * ((L.cout + 127) / 128);
    f.ksplit = std::max(1, std::min((R + 255) / 256, (296 + tiles - 1) / tiles));   // fill the SMs: wgrad has few output tiles
    fc2::Params& g = d[i];
    g.A.g = img[i].dz; g.A.rows = R; g.A.cols = L.cout; g.a_mn = 0;
    g.B.g = img[i].w; g.B.rows = L.cin; g.B.cols = L.cout; g.b_mn = 0;    // rows = output index (cin), cols = K (cout)
    g.C = dX[i]; g.ldc = lddx; g.M = R; g.N = L.cin; g.K = L.cout;
  }
  if (!dz_packed) AN3D_TRY(fc2::pack(pa[0], st, nbr == 2 ? &pa[1] : nullptr));   // (the BN backward wrote the image itself)
  // wgrad and dgrad of a layer read the same images and write disjoint outputs; neither fills the GPU (a few dozen CTAs
  // each), so wgrad runs on the side stream next to dgrad -- fork / join with events, legal inside a capture
#ifndef AN3D_FC_NO_SIDE
  SideStream* ss = (dX[0] && two_streams_enabled()) ? side_stream() : nullptr;
#else
  SideStream* ss = nullptr;
#endif
  if (ss) {
    AN3D_CUDA_CHECK(cudaEventRecord(ss->fork, st));
    AN3D_CUDA_CHECK(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
    const int r0 = fc2::launch(w[0], ss->stream, nbr == 2 ? &w[1] : nullptr);
    const int r1 = fc2::launch(d[0], st, nbr == 2 ? &d[1] : nullptr);
    AN3D_CUDA_CHECK(cudaEventRecord(ss->join, ss->stream));
    AN3D_CUDA_CHECK(cudaStreamWaitEvent(st, ss->join, 0));
    return r0 != AN3D_OK ? r0 : r1;
  }
  AN3D_TRY(fc2::launch(w[0], st, nbr == 2 ? &w[1] : nullptr));
  if (dX[0]) AN3D_TRY(fc2::launch(d[0], st, nbr == 2 ? &d[1] : nullptr));
  return AN3D_OK;
}

// ---- the 3-wide first conv layer (utils/tf_util.py:157 with the [1,3] kernel): its wgrad and dgrad read dZ [R, cout] once.
// The tiled SGEMM spent 260 / 233 us per launch on them at R = 819,200 (64 x 64 output tiles for a 3-row / 3-column
// result); these stream dZ at HBM speed.  Used by every mode of the materialised path.
// wgrad: grads.W[k, c] += sum_r X[r, k] dZ[r, c]   (k < 3).  Block = 256 threads = 4 row lanes x 64 channels... generalised:
// thread (ry, c) walks rows ry, ry + RY, ... of the block's row range for channel c; fp32 partials per thread, fp64 in the block.
static __global__ void __launch_bounds__(256) l0_wgrad_kernel(const float* X, const float* dZ, float* gW, int R, int C, int rows_per_block) {
  const int cpb = min(C, 64);                       // channels per block pass
  const int ry = threadIdx.x / cpb, ny = 256 / cpb;
  const int cl = threadIdx.x - ry * cpb;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  __shared__ double red[3][256];
  for (int cb = blockIdx.y * cpb; cb < C; cb += gridDim.y * cpb) {
    const int c = cb + cl;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (ry < ny && c < C) {
      // four independent rows in flight per thread; fp32 partials over the block's short row range, fp64 across blocks' lanes
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
      int r = r0 + ry;
      for (; r + 3 * ny < r1; r += 4 * ny) {
        float g[4], x[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          g[u] = dZ[(int64_t)(r + u * ny) * C + c];
          const float* xp = X + (int64_t)(r + u * ny) * 3;
          x[u][0] = xp[0]; x[u][1] = xp[1]; x[u][2] = xp[2];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { s0 = fmaf(x[u][0], g[u], s0); s1 = fmaf(x[u][1], g[u], s1); s2 = fmaf(x[u][2], g[u], s2); }
      }
      for (; r < r1; r += ny) {
        const float g = dZ[(int64_t)r * C + c];
        const float* xp = X + (int64_t)r * 3;
        s0 = fmaf(xp[0], g, s0); s1 = fmaf(xp[1], g, s1); s2 = fmaf(xp[2], g, s2);
      }
      a0 = s0; a1 = s1; a2 = s2;
    }
    red[0][threadIdx.x] = a0; red[1][threadIdx.x] = a1; red[2][threadIdx.x] = a2;
    __syncthreads();
    if (ry == 0 && c < C) {
      for (int i = 1; i < ny; ++i) { a0 += red[0][i * cpb + cl]; a1 += red[1][i * cpb + cl]; a2 += red[2][i * cpb + cl]; }
      atomicAdd(gW + c, (float)a0);
      atomicAdd(gW + C + c, (float)a1);
      atomicAdd(gW + 2 * C + c, (float)a2);
    }
    __syncthreads();
  }
}

// dgrad: dX[r, k] = sum_c dZ[r, c] W[k, c]   (k < 3): a warp per row group, lanes across channels (coalesced), shuffle reduce
static __global__ void __launch_bounds__(256) l0_dgrad_kernel(const float* dZ, const float* W, float* dX, int R, int C) {
  extern __shared__ float sw[];                     // [3][C]
  for (int i = threadIdx.x; i < 3 * C; i += 256) sw[i] = W[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < R; r += (int64_t)gridDim.x * 8) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float g = dZ[r * C + c];
      s0 = fmaf(g, sw[c], s0); s1 = fmaf(g, sw[C + c], s1); s2 = fmaf(g, sw[2 * C + c], s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) { dX[r * 3] = s0; dX[r * 3 + 1] = s1; dX[r * 3 + 2] = s2; }
  }
}

// Z = pro(X) W + b.  Given dZ [R,cout] (dense, ld = cout):
//   grads.W += pro(X)^T dZ ; grads.b += colsum(dZ) ; dX = dZ W^T (if dX != nullptr).
int linear_backward(const Lin& L, const float* X, int64_t ldx, const float* psc, const float* psh, const float* pmask,
                    float pmask_scale, const float* dZ, int R, const float* params, float* grads, float* dX,
                    int64_t lddx, double* bias_acc, cudaStream_t st, const FcImages* img = nullptr,
                    const PlanF32* tp = nullptr, const tcg::SplitMat* dz_img = nullptr, const __nv_bfloat16* x_kept = nullptr) {
  const bool bf16 = img != nullptr;
  bool wgrad_done = false, dgrad_done = false;
  bool skip_bias = bf16 && L.bn >= 0;
  if (!bf16 && tp && tcg::use_tensor_cores(*tp, L.cin, L.cout, R)) {
    skip_bias = L.bn >= 0 && dz_img != nullptr;   // (no fp32 dZ to reduce; the gradient is identically zero, see below)
    // materialised path on the tensor cores (gemm_tc.cuh): the layer's input (BN + ReLU + dropout of the producing layer
    // applied while packing), the gradient at its output and its weights are packed once and serve wgrad and dgrad
    tcg::SplitMat x, dz, w;
    if (x_kept) x = tcg::kept_images(*tp, x_kept, R, L.cin);     // the forward kept them
    else AN3D_TRY(tcg::pack_slot(*tp, tcg::SLOT_X, X, ldx, R, L.cin, psc, psh, pmask, pmask_scale, &x, st));
    if (dz_img) dz = *dz_img;     // the BN backward wrote the images itself
    else AN3D_TRY(tcg::pack_slot(*tp, tcg::SLOT_DZ, dZ, L.cout, R, L.cout, nullptr, nullptr, nullptr, 1.f, &dz, st));
    tcg::Params f;
    f.A = x; f.a_mn = 1; f.B = dz; f.b_mn = 1;                                   // contraction over the rows
    f.C = grads + L.w; f.ldc = L.cout; f.M = L.cin; f.N = L.cout; f.K = R; f.accumulate = 1;
    AN3D_TRY(tcg::launch(f, st));
    if (dX) {
      AN3D_TRY(tcg::pack_slot(*tp, tcg::SLOT_W, params + L.w, L.cout, L.cin, L.cout, nullptr, nullptr, nullptr, 1.f, &w, st));
      tcg::Params g;
      g.A = dz; g.a_mn = 0; g.B = w; g.b_mn = 0;                                 // W image: rows = cin (output index), cols = cout (K)
      g.C = dX; g.ldc = lddx; g.M = R; g.N = L.cin; g.K = L.cout;
      AN3D_TRY(tcg::launch(g, st));
    }
    wgrad_done = dgrad_done = true;
  }
  if (bf16) {
    const FcImages im[2] = {*img, FcImages()};
    const float* dz[2] = {dZ, nullptr};
    float* dx[2] = {dX, nullptr};
    AN3D_TRY(fc_backward_bf16(L, R, 1, im, dz, grads + L.w, dx, lddx, st, img->dz_packed));
    wgrad_done = dgrad_done = true;
  }
  // the 3-wide first conv layer at conv-stack row counts: dedicated streaming kernels (no prologue, dense operands)
  const bool l0 = !bf16 && L.cin == 3 && ldx == 3 && !psc && !pmask && R >= 4096 && L.cout <= 1024;
  if (!wgrad_done && l0) {
    const int rpb = 512;
    dim3 grid((R + rpb - 1) / rpb, 1);
    l0_wgrad_kernel<<<grid, 256, 0, st>>>(X, dZ, grads + L.w, R, L.cout, rpb);
    AN3D_LAUNCH_CHECK();
    wgrad_done = true;
  }
  if (dX && !dgrad_done && l0 && lddx == 3) {
    const unsigned blocks = (unsigned)std::min<int64_t>(((int64_t)R + 7) / 8, 148 * 16);
    l0_dgrad_kernel<<<blocks, 256, 3 * L.cout * sizeof(float), st>>>(dZ, params + L.w, dX, R, L.cout);
    AN3D_LAUNCH_CHECK();
    dgrad_done = true;
  }
  if (!wgrad_done) {  // wgrad: [cin, cout] += X^T [cin, R] * dZ [R, cout]
    GemmArgs g;
    g.A = X; g.lda = ldx; g.B = dZ; g.ldb = L.cout; g.C = grads + L.w; g.ldc = L.cout;
    g.M = L.cin; g.N = L.cout; g.K = R; g.pro_scale = psc; g.pro_shift = psh; g.pro_mask = pmask;
    g.pro_mask_scale = pmask_scale; g.accumulate = 1;
    const int tiles = ((L.cin + 63) / 64) * ((L.cout + 63) / 64);
    g.ksplit = std::max(1, std::min((R + 255) / 256, (592 + tiles - 1) / tiles));
    AN3D_TRY(launch_gemm(g, true, false, st));
  }
  // bias grad.  A bias that feeds a batch-statistics BN has an identically zero gradient (TF computes rounding
  // noise there); the bf16 path leaves it at zero, as it does for the conv stacks.
  if (!skip_bias) {
    AN3D_CUDA_CHECK(cudaMemsetAsync(bias_acc, 0, sizeof(double) * L.cout, st));
    ColArgs a;
    a.Z = dZ; a.ldz = L.cout; a.R = R; a.C = L.cout; a.acc0 = bias_acc;
    AN3D_TRY(launch_col_reduce(a, COL_SUM, st));
    add_double_to_float_kernel<<<(L.cout + 127) / 128, 128, 0, st>>>(bias_acc, grads + L.b, L.cout);
    AN3D_LAUNCH_CHECK();
  }
  if (dX && !dgrad_done) {  // dgrad: [R, cin] = dZ [R, cout] * W^T
    GemmArgs g;
    g.A = dZ; g.lda = L.cout; g.B = params + L.w; g.ldb = L.cout; g.C = dX; g.ldc = lddx;
    g.M = R; g.N = L.cin; g.K = L.cout;
    AN3D_TRY(launch_gemm(g, false, true, st));
  }
  return AN3D_OK;
}

// Backward through a conv stack + max-pool.  dG: [B, C_last] with leading dim ldg.
// Leaves d(stage input) in p.dpin when want_input_grad.
int conv_stack_backward(const Model& m, const PlanF32& p, int s, int br, const float* dG, int64_t ldg,
                        const float* params, float* grads, bool want_input_grad, cudaStream_t st) {
  const int64_t M = p.M;
  const int nl = (int)m.conv[s].size();
  int cur = 0;
  const bool tc = p.tc_split > 0;
  if (!tc) {
    const int C = m.conv[s].back().cout;
    AN3D_CUDA_CHECK(cudaMemsetAsync(p.dbuf[cur], 0, sizeof(float) * M * C, st));
    const int64_t total = (int64_t)p.B * C;
    pool_bwd_scatter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dG, ldg, p.gidx[s][br], p.dbuf[cur], p.N,
                                                                             C, p.B);
    AN3D_LAUNCH_CHECK();
  }
  for (int l = nl - 1; l >= 0; --l) {
    const Lin& L = m.conv[s][l];
    BnRef v = bn_ref(m, p, grads, false, br, L.bn);
    tcg::SplitMat dz_img;
    bool to_img = false;
    if (tc) {
      // tensor-core modes: the pooled layer's gradient stays sparse (dG at the arg rows), dZ of a layer whose GEMMs run
      // on the tensor cores is written as the split images those GEMMs read
      to_img = tcg::use_tensor_cores(p, L.cin, L.cout, (int)M);
      AN3D_TRY(bn_relu_backward_split(p, v, p.z[s][l][br], (int)M, l == nl - 1, p.dbuf[cur], dG, ldg, p.gidx[s][br], to_img,
                                      &dz_img, p.dbuf[cur], st));
    } else {
      AN3D_TRY(bn_relu_backward(v, p.z[s][l][br], (int)M, p.dbuf[cur], L.cout, nullptr, 1.f, p.dbuf[cur], L.cout, st));
    }
    const float *X, *psc = nullptr, *psh = nullptr;
    float* dX = nullptr;
    int64_t lddx = 0;
    if (l > 0) {
      const Lin& P = m.conv[s][l - 1];
      X = p.z[s][l - 1][br];
      const int64_t sl = m.bn_slot_off(false, br, P.bn);
      psc = p.bn.scale + sl;
      psh = p.bn.shift + sl;
      dX = p.dbuf[cur ^ 1];
      lddx = L.cin;
    } else {
      X = p.pin[s][br];
      if (want_input_grad) { dX = p.dpin; lddx = 3; }
    }
    AN3D_TRY(linear_backward(L, X, L.cin, psc, psh, nullptr, 1.f, p.dbuf[cur], (int)M, params, grads, dX, lddx,
                             p.dbias_acc, st, nullptr, &p, to_img ? &dz_img : nullptr, p.tcx[s][l][br]));
    cur ^= 1;
  }
  return AN3D_OK;
}

// Backward through get_mlp.  dOut: [B, out] dense.  x/ldx: the MLP input (already activated).
// Writes d(input) to dIn (leading dim lddin).
int mlp_backward(const Model& m, const PlanF32& p, int s, int br, const float* x, int64_t ldx, const float* dOut,
                 const float* params, float* grads, float* dIn, int64_t lddin, const float* mask, cudaStream_t st) {
  const bool head = s == HEAD;
  const int nl = (int)m.fc[s].size();
  const float* dZ = dOut;
  int cur = 0;
  const float mask_scale = mask ? 1.0f / m.arch.keep_prob[s] : 1.f;
  for (int l = nl - 1; l >= 0; --l) {
    const Lin& L = m.fc[s][l];
    if (L.bn >= 0) {
      BnRef v = bn_ref(m, p, grads, head, br, L.bn);
      // dZ currently holds dA (grad wrt this layer's post-activation); dropout only after the last hidden layer
      const bool dropped = (l == nl - 2) && mask;
      float* buf = const_cast<float*>(dZ);
      AN3D_TRY(bn_relu_backward(v, p.fz[s][l][br], p.B, dZ, L.cout, dropped ? mask : nullptr, mask_scale, buf, L.cout, st,
                                p.bf16 ? p.fcdz[0] : nullptr));
    }
    const float *X, *psc = nullptr, *psh = nullptr, *pm = nullptr;
    int64_t lx;
    float* dX;
    int64_t lddx;
    if (l > 0) {
      const Lin& P = m.fc[s][l - 1];
      X = p.fz[s][l - 1][br];
      lx = P.cout;
      const int64_t sl = m.bn_slot_off(head, br, P.bn);
      psc = p.bn.scale + sl;
      psh = p.bn.shift + sl;
      if (l == nl - 1) pm = mask;
      dX = p.dfc[cur];
      lddx = L.cin;
    } else {
      X = x;
      lx = ldx;
      dX = dIn;
      lddx = lddin;
    }
    FcImages im;
    im.x = p.fcx[s][l][br]; im.w = p.fcw[s][l]; im.dz = p.fcdz[0]; im.dz_packed = L.bn >= 0;
    AN3D_TRY(linear_backward(L, X, lx, psc, psh, pm, mask_scale, dZ, p.B, params, grads, dX, lddx, p.dbias_acc, st,
                             p.bf16 ? &im : nullptr, &p));
    dZ = dX;
    cur ^= 1;
  }
  return AN3D_OK;
}

// bf16 mode, both siamese branches of one FC stack at once: the BN/ReLU backward runs per branch, the wgrad and dgrad
// GEMMs of a layer are ONE launch each with the two branches as a batch (wgrad accumulates both into grads.W).
int mlp_backward_pair(const Model& m, const PlanF32& p, int s, const float* const x[2], int64_t ldx, const float* const dOut[2],
                      const float* params, float* grads, float* const dIn[2], int64_t lddin, const float* const mask[2],
                      cudaStream_t st) {
  const int nl = (int)m.fc[s].size();
  const float* dZ[2] = {dOut[0], dOut[1]};
  float* const scratch[2][2] = {{p.dfc[0], p.dfc[1]}, {p.dfc_b[0], p.dfc_b[1]}};
  int cur = 0;
  for (int l = nl - 1; l >= 0; --l) {
    const Lin& L = m.fc[s][l];
    const int R = p.B;
    for (int br = 0; br < 2; ++br) {
      if (L.bn < 0) continue;
      BnRef v = bn_ref(m, p, grads, false, br, L.bn);
      const bool dropped = (l == nl - 2) && mask[br];
      const float ms = mask[br] ? 1.0f / m.arch.keep_prob[s] : 1.f;
      AN3D_TRY(bn_relu_backward(v, p.fz[s][l][br], R, dZ[br], L.cout, dropped ? mask[br] : nullptr, ms,
                                const_cast<float*>(dZ[br]), L.cout, st, p.fcdz[br]));
    }
    float* dX[2];
    int64_t lddx;
    for (int br = 0; br < 2; ++br) {
      if (l > 0) {
        dX[br] = scratch[br][cur];
        lddx = L.cin;
      } else {
        dX[br] = dIn[br];
        lddx = lddin;
      }
    }
    if (L.bn < 0) {   // only a bias that does not feed a batch-statistics BN has a non-zero gradient
      for (int br = 0; br < 2; ++br) {
        AN3D_CUDA_CHECK(cudaMemsetAsync(p.dbias_acc, 0, sizeof(double) * L.cout, st));
        ColArgs a;
        a.Z = dZ[br]; a.ldz = L.cout; a.R = R; a.C = L.cout; a.acc0 = p.dbias_acc;
        AN3D_TRY(launch_col_reduce(a, COL_SUM, st));
        add_double_to_float_kernel<<<(L.cout + 127) / 128, 128, 0, st>>>(p.dbias_acc, grads + L.b, L.cout);
        AN3D_LAUNCH_CHECK();
      }
    }
    FcImages img[2];
    for (int br = 0; br < 2; ++br) {
      img[br].x = p.fcx[s][l][br]; img[br].w = p.fcw[s][l]; img[br].dz = p.fcdz[br];
    }
    AN3D_TRY(fc_backward_bf16(L, R, 2, img, dZ, grads + L.w, dX, lddx, st, L.bn >= 0));
    dZ[0] = dX[0];
    dZ[1] = dX[1];
    cur ^= 1;
  }
  return AN3D_OK;
}

// dOh[:, :3] = dpred_t ; dOh[:, 3:] = drem ; dc2[0] = ds2c1 - dpred_t ; dc2[1] = ds2c2 + dpred_t  (tp8.py:155)
// (one warp per sample)
__global__ void assemble_head_grad_kernel(const float* dend, float* dOh, float* dc2a, float* dc2b, int B, int nb) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* ds2c1 = dend + (int64_t)2 * B * 3;
  const float* ds2c2 = dend + (int64_t)3 * B * 3;
  const float* dpt = dend + (int64_t)4 * B * 3;
  const float* drem = dend + (int64_t)5 * B * 3 + (int64_t)2 * B * 2 * nb;
  float* o = dOh + (int64_t)b * (3 + 2 * nb);
  if (lane < 3) {
    const float g = dpt[b * 3 + lane];
    o[lane] = g;
    dc2a[b * 3 + lane] = ds2c1[b * 3 + lane] - g;
    dc2b[b * 3 + lane] = ds2c2[b * 3 + lane] + g;
  }
  for (int j = lane; j < 2 * nb; j += 32) o[3 + j] = drem[(int64_t)b * 2 * nb + j];
}

// dO2[:, :3] = dc2 ; dO2[:, 3:] = dlg (+ dang*pi/nb at nb+k, tp8.py:298) ; dc1 = ds1c + dc2  (tp8.py:109,117)
__global__ void assemble_s2_grad_kernel(const float* dlg, const float* dc2, const float* dang, const int32_t* angk,
                                        const float* ds1c, float* dO2, float* dc1, int B, int nb) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  float* o = dO2 + (int64_t)b * (3 + 2 * nb);
  if (lane < 3) {
    o[lane] = dc2[b * 3 + lane];
    dc1[b * 3 + lane] = ds1c[b * 3 + lane] + dc2[b * 3 + lane];
  }
  const int kk = nb + angk[b];
  const float extra = dang[b] * (3.14159265358979323846f / (float)nb);
  for (int j = lane; j < 2 * nb; j += 32) o[3 + j] = dlg[(int64_t)b * 2 * nb + j] + (j == kk ? extra : 0.f);
}

}  // namespace

// the two branches of a conv stage backward: independent (their own scratch, parameter gradients accumulate with
// reductions), so branch 1 runs on the side stream and its ~8 small launches hide under branch 0's persistent kernels
template <typename F>
static int two_branches(cudaStream_t st, F&& run) {
  SideStream* ss = two_streams_enabled() ? side_stream() : nullptr;
  if (!ss) {
    AN3D_TRY(run(0, st));
    return run(1, st);
  }
  AN3D_CUDA_CHECK(cudaEventRecord(ss->fork, st));
  AN3D_CUDA_CHECK(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
  const int r0 = run(0, st);
  const int r1 = run(1, ss->stream);
  AN3D_CUDA_CHECK(cudaEventRecord(ss->join, ss->stream));      // always join: a capture must not end with a dangling fork
  AN3D_CUDA_CHECK(cudaStreamWaitEvent(st, ss->join, 0));
  return r0 != AN3D_OK ? r0 : r1;
}

int backward_impl(const Model& m, const float* params, const float* pcs1, const float* pcs2, const an3d_labels* labels,
                  const an3d_outputs* out, int B, int N, int flags, float* grads, float* loss_out, void* workspace,
                  int64_t workspace_bytes, cudaStream_t st) {
  PlanF32 p;
  AN3D_TRY(plan_f32(m, B, N, flags | AN3D_TRAINING, workspace, &p));
  if (p.bytes > workspace_bytes) {
    set_error("workspace too small: need %lld bytes, got %lld", (long long)p.bytes, (long long)workspace_bytes);
    return AN3D_ERR_WORKSPACE;
  }
  const int nb = m.nb;
  const bool bf16 = p.bf16;
  const float* pcs[2] = {pcs1, pcs2};
  const float* c1o[2] = {out->pred_s1_pc1centers, out->pred_s1_pc2centers};
  const float* c2o[2] = {out->pred_s2_pc1centers, out->pred_s2_pc2centers};
  if (bf16) AN3D_TRY(pack_weights_bf16_bwd(m, p, params, st));
  prof_mark(PROF_LOSS, true, st);
  AN3D_TRY(run_loss(m, labels, out, B, loss_out, p.loss_scratch, p.dend, st));
  prof_mark(PROF_LOSS, false, st);
  AN3D_CUDA_CHECK(cudaMemsetAsync(grads, 0, sizeof(float) * m.n_trainable, st));
  AN3D_CUDA_CHECK(cudaMemsetAsync(p.bn.acc0, 0, sizeof(double) * m.bn_total_ch(), st));
  AN3D_CUDA_CHECK(cudaMemsetAsync(p.bn.acc1, 0, sizeof(double) * m.bn_total_ch(), st));
  const unsigned w_blocks = (unsigned)((B + 3) / 4);   // one warp per sample
  const float* masks[5];
  for (int i = 0; i < 5; ++i) {
    const int s = i < 2 ? S1 : (i < 4 ? S2 : HEAD);
    masks[i] = m.arch.keep_prob[s] < 1.f ? p.mask[i] : nullptr;
  }
  const int c_emb = m.conv[EMB].back().cout;
  // head
  assemble_head_grad_kernel<<<w_blocks, 128, 0, st>>>(p.dend, p.dout, p.dc2[0], p.dc2[1], B, nb);
  AN3D_LAUNCH_CHECK();
  AN3D_TRY(mlp_backward(m, p, HEAD, 0, p.feat, 2 * c_emb, p.dout, params, grads, p.dfeat, 2 * c_emb, masks[4], st));
  const float* dlg[2] = {p.dend + (int64_t)5 * B * 3, p.dend + (int64_t)5 * B * 3 + (int64_t)B * 2 * nb};
  const float* ds1c[2] = {p.dend, p.dend + (int64_t)B * 3};
  if (bf16) {
    // stage-major order (mirrors the forward): both branches' FC layers share launches
    const int c2w = m.conv[S2].back().cout, c1w = m.conv[S1].back().cout;
    float* const dO[2] = {p.dout, p.dout_b};
    float* const dGs[2] = {p.dg, p.dg_b};
    AN3D_TRY(two_branches(st, [&](int br, cudaStream_t s2) -> int {
      AN3D_CUDA_CHECK(cudaMemsetAsync(p.dang[br], 0, sizeof(float) * B, s2));
      AN3D_TRY(conv_stack_backward_bf16(m, p, EMB, br, pcs[br], c2o[br], p.ang[br], p.dfeat + (int64_t)br * c_emb,
                                        2 * c_emb, params, grads, true, p.dc2[br], p.dang[br], s2));
      assemble_s2_grad_kernel<<<w_blocks, 128, 0, s2>>>(dlg[br], p.dc2[br], p.dang[br], p.angk[br], ds1c[br], dO[br],
                                                        p.dc1[br], B, nb);
      AN3D_LAUNCH_CHECK();
      return AN3D_OK;
    }));
    {
      const float* x[2] = {p.g[S2][0], p.g[S2][1]};
      const float* dOut[2] = {dO[0], dO[1]};
      const float* mk[2] = {masks[2], masks[3]};
      AN3D_TRY(mlp_backward_pair(m, p, S2, x, c2w, dOut, params, grads, dGs, c2w, mk, st));
    }
    AN3D_TRY(two_branches(st, [&](int br, cudaStream_t s2) -> int {
      return conv_stack_backward_bf16(m, p, S2, br, pcs[br], c1o[br], nullptr, dGs[br], c2w, params, grads, true, p.dc1[br],
                                      nullptr, s2);
    }));
    {
      // stage 1: d(delta1) = dc1 (tp8.py:109); its input p - mean(p) carries no parameter gradient
      const float* x[2] = {p.g[S1][0], p.g[S1][1]};
      const float* dOut[2] = {p.dc1[0], p.dc1[1]};
      const float* mk[2] = {masks[0], masks[1]};
      AN3D_TRY(mlp_backward_pair(m, p, S1, x, c1w, dOut, params, grads, dGs, c1w, mk, st));
    }
    AN3D_TRY(two_branches(st, [&](int br, cudaStream_t s2) -> int {
      return conv_stack_backward_bf16(m, p, S1, br, pcs[br], p.mu[br], nullptr, dGs[br], c1w, params, grads, false, nullptr,
                                      nullptr, s2);
    }));
    return AN3D_OK;
  }
  for (int br = 0; br < 2; ++br) {
    // final embedding stack: input q = Rz(a)(p - c2)
    if (bf16) {
      AN3D_CUDA_CHECK(cudaMemsetAsync(p.dang[br], 0, sizeof(float) * B, st));
      AN3D_TRY(conv_stack_backward_bf16(m, p, EMB, br, pcs[br], c2o[br], p.ang[br], p.dfeat + (int64_t)br * c_emb,
                                        2 * c_emb, params, grads, true, p.dc2[br], p.dang[br], st));
    } else {
      AN3D_TRY(conv_stack_backward(m, p, EMB, br, p.dfeat + (int64_t)br * c_emb, 2 * c_emb, params, grads, true, st));
      input_bwd_kernel<<<(B + 3) / 4, 128, 0, st>>>(p.dpin, p.pin[EMB][br], p.ang[br], p.dc2[br], p.dang[br], N, B);
      AN3D_LAUNCH_CHECK();
    }
    // stage 2
    assemble_s2_grad_kernel<<<w_blocks, 128, 0, st>>>(dlg[br], p.dc2[br], p.dang[br], p.angk[br], ds1c[br], p.dout,
                                                      p.dc1[br], B, nb);
    AN3D_LAUNCH_CHECK();
    const int c2w = m.conv[S2].back().cout;
    AN3D_TRY(mlp_backward(m, p, S2, br, p.g[S2][br], c2w, p.dout, params, grads, p.dg, c2w, masks[2 + br], st));
    if (bf16) {
      AN3D_TRY(conv_stack_backward_bf16(m, p, S2, br, pcs[br], c1o[br], nullptr, p.dg, c2w, params, grads, true,
                                        p.dc1[br], nullptr, st));
    } else {
      AN3D_TRY(conv_stack_backward(m, p, S2, br, p.dg, c2w, params, grads, true, st));
      input_bwd_kernel<<<(B + 3) / 4, 128, 0, st>>>(p.dpin, p.pin[S2][br], nullptr, p.dc1[br], nullptr, N, B);
      AN3D_LAUNCH_CHECK();
    }
    // stage 1: d(delta1) = dc1 (tp8.py:109); its input p - mean(p) carries no parameter gradient
    const int c1w = m.conv[S1].back().cout;
    AN3D_TRY(mlp_backward(m, p, S1, br, p.g[S1][br], c1w, p.dc1[br], params, grads, p.dg, c1w, masks[br], st));
    if (bf16) {
      AN3D_TRY(conv_stack_backward_bf16(m, p, S1, br, pcs[br], p.mu[br], nullptr, p.dg, c1w, params, grads, false,
                                        nullptr, nullptr, st));
    } else {
      AN3D_TRY(conv_stack_backward(m, p, S1, br, p.dg, c1w, params, grads, false, st));
    }
  }
  return AN3D_OK;
}

}  // namespace an3d
