// placeholder until the tcgen05 path lands
#include "common.cuh"
namespace an3d {
int plan_bf16_bytes(const Model&, int, int, int, int64_t*) { set_error("bf16 path not built yet"); return AN3D_ERR_UNSUPPORTED; }
int forward_bf16(const Model&, const float*, float*, const float*, const float*, int, int, int, float, const an3d_dropout*, const an3d_outputs*, void*, int64_t, cudaStream_t) { set_error("bf16 path not built yet"); return AN3D_ERR_UNSUPPORTED; }
int backward_bf16(const Model&, const float*, const float*, const float*, const an3d_labels*, const an3d_outputs*, int, int, int, float*, float*, void*, int64_t, cudaStream_t) { set_error("bf16 path not built yet"); return AN3D_ERR_UNSUPPORTED; }
}
