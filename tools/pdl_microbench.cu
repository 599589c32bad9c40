// Programmatic dependent launch inside a captured CUDA graph: time per node of a chain of N dependent small kernels,
// launched normally vs with cudaLaunchAttributeProgrammaticStreamSerialization (early trigger + grid-dependency wait at the
// top of every kernel).  Decides whether the library's chains of small kernels (BN folds, partial sums, packs: ~270 launches
// per c3 step) are worth converting.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/pdl_microbench tools/pdl_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int WORK>
__global__ void chain_kernel(float* p, int n) {
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = p[i];
#pragma unroll 1
    for (int k = 0; k < WORK; ++k) v = fmaf(v, 1.0000001f, 1e-7f);
    p[i] = v;
  }
}

template <int WORK>
static int run(const char* label, int blocks, int threads, int nodes, bool pdl, float* buf, cudaStream_t st) {
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
  for (int k = 0; k < nodes; ++k) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    if (pdl) { cfg.attrs = attr; cfg.numAttrs = 1; }
    CK(cudaLaunchKernelEx(&cfg, chain_kernel<WORK>, buf, blocks * threads));
  }
  CK(cudaStreamEndCapture(st, &graph));
  CK(cudaGraphInstantiate(&exec, graph, 0));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int w = 0; w < 5; ++w) CK(cudaGraphLaunch(exec, st));
  CK(cudaStreamSynchronize(st));
  const int reps = 50;
  CK(cudaEventRecord(a, st));
  for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(exec, st));
  CK(cudaEventRecord(b, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("%-34s grid %4d x %3d  work %5d  %s: %7.3f us per node\n", label, blocks, threads, WORK, pdl ? "PDL   " : "normal", ms * 1000.0 / (reps * nodes));
  cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
  return 0;
}

int main() {
  float* buf;
  CK(cudaMalloc(&buf, 148 * 8 * 256 * sizeof(float)));
  CK(cudaMemset(buf, 0, 148 * 8 * 256 * sizeof(float)));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  for (int pdl = 0; pdl < 2; ++pdl) {
    if (run<1>("one tiny block", 1, 128, 256, pdl, buf, st)) return 1;
    if (run<1>("one wave of tiny blocks", 148, 256, 256, pdl, buf, st)) return 1;
    if (run<2000>("one wave, ~5 us of work", 148, 256, 256, pdl, buf, st)) return 1;
    if (run<2000>("four waves-worth (8 blocks / SM)", 148 * 8, 256, 256, pdl, buf, st)) return 1;
  }
  // eager (no graph) launches of the same chain
  for (int pdl = 0; pdl < 2; ++pdl) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a, st));
    for (int k = 0; k < 2000; ++k) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      if (pdl) { cfg.attrs = attr; cfg.numAttrs = 1; }
      CK(cudaLaunchKernelEx(&cfg, chain_kernel<1>, buf, 148 * 256));
    }
    CK(cudaEventRecord(b, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    printf("eager stream launches, one wave     %s: %7.3f us per launch\n", pdl ? "PDL   " : "normal", ms * 1000.0 / 2000);
  }
  return 0;
}
