#pragma once
#include "common.cuh"

namespace an3d {

int bf16_supported(const Model& m);
void plan_bf16(const Model& m, int B, int N, int flags, Arena& a, PlanBf16* q);
int pack_weights_bf16(const Model& m, const PlanF32& p, const float* params, cudaStream_t st);
// One conv stack (3 -> 64 -> 128 -> C3 + max-pool) of one branch on the tensor cores.  Fills
// p.g[s][br] (+ p.gidx in training), the BN scratch of its three BN layers and the EMA state.
int conv_stack_forward_bf16(const Model& m, const PlanF32& p, int s, int br, const float* pcs, const float* center,
                            const float* angle, const float* params, float* state, bool training, float decay,
                            cudaStream_t st);

int pack_weights_bf16_bwd(const Model& m, const PlanF32& p, const float* params, cudaStream_t st);
int conv_stack_backward_bf16(const Model& m, const PlanF32& p, int s, int br, const float* pcs, const float* center,
                             const float* angle, const float* dG, int64_t lddg, const float* params, float* grads,
                             bool want_input_grad, float* dcenter, float* dangle, cudaStream_t st);

}  // namespace an3d
