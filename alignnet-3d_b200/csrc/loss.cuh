#pragma once
#include "common.cuh"

namespace an3d {

// Scratch of the loss kernels (carved from PlanF32::loss_scratch / the bf16 plan).
struct LossScratch {
  double* sums;  // [16]: 5 huber sums, 6 cross-entropy sums (inst*2 + variant)
  double* S;     // [6][B] per-column huber sums of the [B,B] residual loss
  double* G;     // [6][B] per-column huber' sums
  float* pd;     // [B] decoded pc2 - pc1 stage-2 yaw (tp8.py:325-327)
  float* gt3;    // [B] pc2_angles - pc1_angles
  float* pred;   // [6][B] residual logit at the target class
  float* lab;    // [4][B] normalised residual labels of the two stage-2 losses
  int* cls;      // [6][B] target classes
  int* k1;       // [B] argmax bins of the stage-2 logits
  int* k2;
  int* sel;      // [3] selected variant (0 = target, 1 = target + pi)
  int64_t bytes;
};

LossScratch carve_loss_scratch(float* base, int B);
int64_t loss_scratch_floats(int B);

// dend layout (floats): ds1c1[B,3] ds1c2[B,3] ds2c1[B,3] ds2c2[B,3] dpred_t[B,3] dlg1[B,2nb] dlg2[B,2nb] drem[B,2nb]
int run_loss(const Model& m, const an3d_labels* lb, const an3d_outputs* out, int B, float* loss_out, float* scratch,
             float* dend, cudaStream_t st);

}  // namespace an3d
