// Internal declarations shared by the translation units of libalignnet_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/alignnet_b200.h"

namespace an3d {

constexpr float kBnEps = 1e-3f;  // utils/tf_util.py:491

void set_error(const char* fmt, ...);

#define AN3D_CUDA_CHECK(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      an3d::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return AN3D_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

#define AN3D_TRY(expr)          \
  do {                          \
    int _r = (expr);            \
    if (_r != AN3D_OK) return _r; \
  } while (0)

extern unsigned long long g_launch_count;   // kernels launched by this library (an3d_launch_count)
#define AN3D_LAUNCH_CHECK()                  \
  do {                                       \
    ++an3d::g_launch_count;                  \
    AN3D_CUDA_CHECK(cudaGetLastError());     \
  } while (0)

// Optional per-kernel timing (an3d_profile_begin/end): CUDA events recorded on the launch stream
// around the tagged heavy kernels.
enum ProfTag { PROF_CONV_STATS2 = 0, PROF_CONV_FULL = 1, PROF_BWD_T1 = 2, PROF_BWD_DGRAD3 = 3, PROF_BWD_L2 = 4,
               PROF_FC = 5, PROF_LOSS = 6, PROF_GRAM2 = 7, PROF_NTAGS = 8 };
void prof_mark(int tag, bool begin, cudaStream_t st);
bool prof_active();

// One linear layer (1x1 conv or FC).  Weights are stored as the reference stores them:
// [Cin, Cout] row-major (conv kernels [1,1,Cin,Cout] / [1,3,1,Cout], FC [Cin,Cout]).
struct Lin {
  int cin = 0, cout = 0;
  int64_t w = 0, b = 0;  // offsets into the flat parameter buffer
  int bn = -1;           // index into the per-branch (or head) BN table, -1 = no BN
  std::string scope;     // TF scope below the branch prefix
};

// One BN layer of a branch (or of the head).
struct BnLayer {
  int ch = 0;
  int64_t choff = 0;  // channel offset inside the per-branch BN block
  std::string scope;
};

enum StageId { S1 = 0, S2 = 1, EMB = 2, HEAD = 2 };

struct TensorInfo {
  std::string name;
  int64_t offset;
  int ndim;
  int64_t shape[4];
};

struct Model {
  an3d_arch arch;
  int nb = 0;
  std::vector<Lin> conv[3];  // S1, S2, EMB
  std::vector<Lin> fc[3];    // S1, S2, HEAD
  std::vector<BnLayer> bn_branch;  // BN layers of one siamese branch, execution order
  std::vector<BnLayer> bn_head;
  int64_t bn_branch_ch = 0, bn_head_ch = 0;
  int64_t bn_base = 0;        // offset of the BN gamma/beta block inside params
  int64_t n_trainable = 0, n_state = 0;
  std::vector<TensorInfo> trainable, state;

  // gamma at returned offset, beta at offset + ch (params); mean at offset, var at + ch (state)
  int64_t bn_param_off(bool head, int branch, int bn) const {
    if (head) return bn_base + 2 * 2 * bn_branch_ch + 2 * bn_head[bn].choff;
    return bn_base + (int64_t)branch * 2 * bn_branch_ch + 2 * bn_branch[bn].choff;
  }
  int64_t bn_state_off(bool head, int branch, int bn) const {
    if (head) return 2 * 2 * bn_branch_ch + 2 * bn_head[bn].choff;
    return (int64_t)branch * 2 * bn_branch_ch + 2 * bn_branch[bn].choff;
  }
  // running slot index for per-step BN scratch (scale/shift/mean/inv): channels before this layer
  int64_t bn_slot_off(bool head, int branch, int bn) const {
    if (head) return 2 * bn_branch_ch + bn_head[bn].choff;
    return (int64_t)branch * bn_branch_ch + bn_branch[bn].choff;
  }
  int64_t bn_total_ch() const { return 2 * bn_branch_ch + bn_head_ch; }
};

int build_model(const an3d_arch* arch, Model* m);

// Bump allocator over the caller's workspace; the same code path sizes it and carves it.
struct Arena {
  char* base = nullptr;
  int64_t off = 0;
  template <typename T>
  T* take(int64_t count) {
    off = (off + 255) & ~int64_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * (int64_t)sizeof(T);
    return p;
  }
};

// Per-step BN scratch, one float per BN channel in each array (indexed by Model::bn_slot_off).
struct BnScratch {
  float* scale = nullptr;   // gamma * rsqrt(var + eps)
  float* shift = nullptr;   // beta - mean * scale
  float* mean = nullptr;    // batch (or moving) mean
  float* inv = nullptr;     // rsqrt(var + eps)
  double* acc0 = nullptr;   // reduction scratch (sum / sum dy)
  double* acc1 = nullptr;   // reduction scratch (sum sq diff / sum dy*xhat)
};

// Extra buffers of the bf16 tcgen05 path (conv stacks); FC layers, loss and the inter-stage
// glue reuse the PlanF32 buffers.
// Backward scratch of ONE siamese branch (the two branches of a stage run on two streams, so each has its own).
struct BwdScratch {
  float* dyext = nullptr;                   // [B, C3max] gradient at the pooled arg rows (after the ReLU mask)
  double* red3 = nullptr;                   // [2][C3max] sum dy, sum dy*xhat of layer 3
  float* coef3 = nullptr;                   // [4][C3max]: q, p', (unused), (unused)
  __nv_bfloat16* gq = nullptr;              // Gq image, two K halves [2][128][64]
  float* gq_f32 = nullptr;                  // [128][128] fp32 accumulation of Gq
  float* uvec = nullptr;                    // [128]
  __nv_bfloat16* dy2img = nullptr;          // per item image of dy2, TRANSPOSED (convbwd::dy2_img_bytes; fits the a2img size)
  double* red2 = nullptr;                   // [128][2]
  float* coef2 = nullptr;                   // [128][2]  m0, m1
  float* l1sums = nullptr;                  // [items][64][4]: sum_p dy1 * (1, x, y, z) per item and channel
  double* red1 = nullptr;                   // [64][2]
  float* coef1 = nullptr;                   // [64][2]
};

constexpr int kMaxParts1 = 640, kMaxParts = 160;   // CTAs of the statistics / Gram passes (>= SMs x CTAs per SM)

struct PlanBf16 {
  int PC = 0, npc = 0;                      // points per work item (multiple of 16) and items per cloud
  uint32_t idx_mask = 0;                    // low mantissa bits carrying the arg-max point index
  __nv_bfloat16* w2t[3];                    // W2^T plane image [128 ch][64 k]
  __nv_bfloat16* w3t[3][2];                 // sign-folded W3^T plane images, per branch (BN gamma is per branch)
  float* w1f[3][2];                         // layer-1 weights with BN scale folded [3][64]
  float* c1f[3][2];                         // folded layer-1 bias [64]
  float* t2f[3][2];                         // folded layer-2 shift [128]
  double* moments[3][2];                    // first/second moments of the stage input points
  double* stats2[3][2];                     // sum / sum-of-squares of the raw layer-2 accumulator [128][2]
  double* stats3[3][2];                     // same for layer 3 [C3][2]
  float* gram1[3][2];                       // [64][80] Gram matrix + column sums of the layer-1 activations (layer-2 BN statistics)
  float* gram1_parts[2];                    // per branch: [kMaxParts1][64*80] per-CTA partial sums (added in CTA order)
  float* gram_parts[2];                     // per branch: [kMaxParts][128*132] per-CTA partial Gram matrices + column sums
  uint32_t* zext[3][2];                     // packed pooled extreme [B][C3]
  __nv_bfloat16* a2img[3][2];               // saved layer-2 activations (training): per item the A2 smem tile image
  int64_t img_bytes = 0;                    // bytes of one item image (16 planes)
  double* sa2[3][2];                        // column sums of a2 [128]
  // ---- backward scratch (training) ----
  BwdScratch bw[2];                         // per branch
  __nv_bfloat16* w3n[3];                    // W3 (unfolded) half-chunk images [C3/64][128 k][64 c], K-major in c
  __nv_bfloat16* w2p[3];                    // W2 padded to 128 rows, image [128 k1][128 k2]
  float* gram[3][2];                        // [128][128] a2^T a2 per (stage, branch): BN3 statistics and wgrad3
  float* gw[3][2];                          // [128][C3] gram * W3 (bf16-rounded weights), forward -> backward
};

// Workspace plan of the fp32 (parity) path: everything the backward needs is materialised.
struct PlanF32 {
  int B = 0, N = 0;
  bool bf16 = false;                        // AN3D_PRECISION_BF16 with [64,128,C] conv stacks: the fused tensor-core kernels
  int tc_split = 0;                         // materialised path: 0 = CUDA-core SGEMM, 1-3 = bf16 images per GEMM operand (gemm_tc.cuh)
  __nv_bfloat16* tcbuf[3] = {nullptr, nullptr, nullptr};   // image scratch: layer input, output gradient, weights
  int64_t tcbuf_elems[3] = {0, 0, 0};       // capacity of each (all images of the operand)
  __nv_bfloat16* tcx[3][AN3D_MAX_LAYERS][2];  // training: input images of conv layer l >= 1 of (stage, branch), kept for the backward (else null)
  float* bwd_coef = nullptr;                // [2][max conv width]: per-channel constants of the BN backward (tensor-core modes, training)
  bool deterministic = false;               // AN3D_DETERMINISTIC: no split-K fp32 reductions in inference FC layers
  bool prepared = false;                    // AN3D_WEIGHTS_PREPARED (bf16 inference): folded BN / weight images are reused
  int64_t M = 0;  // rows per branch = B*N
  float* pin[3][2];                         // stage input points [M,3]
  float* z[3][AN3D_MAX_LAYERS][2];          // pre-BN conv outputs [M,C]
  float* g[3][2];                           // pooled features [B,C3] (EMB: both point into feat)
  int32_t* gidx[3][2];                      // arg rows of the pool [B,C3]
  float* feat = nullptr;                    // [B, 2*C_emb] head input
  float* fz[3][AN3D_MAX_LAYERS + 1][2];     // FC pre-BN outputs [B,C] (head uses branch 0)
  // bf16 mode: plane-major bf16 images of every matrix that enters an FC GEMM (fc2_gemm.cuh)
  __nv_bfloat16* fcx[3][AN3D_MAX_LAYERS + 1][2];   // input activations of FC layer l of stage s, per branch (head: [0])
  __nv_bfloat16* fcw[3][AN3D_MAX_LAYERS + 1];      // weights of FC layer l of stage s
  __nv_bfloat16* fcdz[2] = {nullptr, nullptr};     // gradient wrt the current layer's output, per branch (training)
  float* mask[5];                           // dropout keep masks [B, width]
  float* mu[2];                             // centroids [B,3]
  float* ang[2];                            // decoded stage-2 yaw [B]
  int32_t* angk[2];                         // argmax bin [B]
  BnScratch bn;
  // backward scratch
  float* dbuf[2];                           // ping-pong activation gradients [max rows * max ch]
  float* dfc[2];                            // FC-sized gradients [2B * max fc width]
  float* dfeat = nullptr;                   // [B, 2*C_emb]
  float* dpin = nullptr;                    // [M,3]
  float* dg = nullptr;                      // [B, max C3] pooled-feature gradient
  float* dout = nullptr;                    // [B, 3+2nb] MLP output gradient
  double* dbias_acc = nullptr;              // [max ch] bias-gradient reduction scratch
  // second set for branch 2 when the bf16 path walks the backward stage by stage with both branches' FC layers batched
  float* dfc_b[2] = {nullptr, nullptr};
  float* dg_b = nullptr;
  float* dout_b = nullptr;
  float* dend = nullptr;                    // gradients of the 8 end_points, packed (see loss.cu)
  float* dc1[2];                            // [B,3] accumulated centre gradients
  float* dc2[2];
  float* dang[2];                           // [B]
  float* loss_scratch = nullptr;            // loss partial sums
  PlanBf16 bf;                              // only carved when AN3D_PRECISION_BF16 is set
  int64_t bytes = 0;
};

int plan_f32(const Model& m, int B, int N, int flags, void* workspace, PlanF32* p);

// elements of the bf16 block image of a [rows, cols] matrix (fc2_gemm.cuh): whole 128 x 128 blocks
inline int64_t fc_image_elems(int rows, int cols) { return (int64_t)((rows + 127) >> 7) * ((cols + 127) >> 7) * 16384; }

struct Ctx {
  Model model;
};

int check_device();

// The two siamese branches of a stage on two streams (forward_f32.cu): fork / join with events, legal inside a capture.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool ok = false;
};
SideStream* side_stream();
bool two_streams_enabled();

}  // namespace an3d

struct an3d_ctx {
  an3d::Ctx impl;
};
