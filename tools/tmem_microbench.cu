// TMEM read-back microbenchmark (sm_100a): how fast can W warps drain accumulator columns with tcgen05.ld, alone
// and with the max-reduction of the conv-stack forward epilogue behind it?  Answers whether the forward kernel's
// epilogue is bound by TMEM read bandwidth, by CUDA-core issue, or by latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tmem_microbench tools/tmem_microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t t, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(cols) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ld16(uint32_t a, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(a) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t a, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(a) : "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// MODE 0: loads only (one register of each load is consumed).  1: + plain max (fmax3, 2 elements / instr).
// 2: + arg index in the low 4 mantissa bits (LOP3 per element) as the training pass does.
// X: 16 or 32 columns per tcgen05.ld.  DEPTH: loads in flight before the first wait (1 = ld, wait, use).
template <int MODE, int X, int DEPTH>
__global__ void __launch_bounds__(512, 1) bench(int cols, int reps, float* out, long long* cycles) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tbase, 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  const int nw = blockDim.x >> 7;                 // warps per lane quarter
  const int wq = warp >> 2;                       // index of this warp within its quarter
  float m = -INFINITY;
  __syncthreads();
  const long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    // the warps of a quarter deal the X-column groups round-robin
    uint32_t r[DEPTH][X];
    int g = wq * X;
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) {
      if (g + d * nw * X < cols) { if (X == 16) ld16(t + g + d * nw * X, (uint32_t(&)[16])r[d]); else ld32(t + g + d * nw * X, (uint32_t(&)[32])r[d]); }
    }
    int slot = 0;
    for (; g < cols; g += nw * X) {
      const int gl = g + (DEPTH - 1) * nw * X;
      const int sl = (slot + DEPTH - 1) % DEPTH;
      if (DEPTH == 1) { if (X == 16) ld16(t + g, (uint32_t(&)[16])r[0]); else ld32(t + g, (uint32_t(&)[32])r[0]); ld_wait(); }
      else {
        ld_wait();
        if (gl < cols) { if (X == 16) ld16(t + gl, (uint32_t(&)[16])r[sl]); else ld32(t + gl, (uint32_t(&)[32])r[sl]); }
      }
      const uint32_t* q = r[slot];
      if (MODE == 0) m = fmaxf(m, __uint_as_float(q[0]));
      else {
        float gm = -INFINITY;
#pragma unroll
        for (int i = 0; i < X; i += 2) {
          if (MODE == 2) gm = fmax3(gm, __uint_as_float((q[i] & ~15u) | (uint32_t)(i & 15)), __uint_as_float((q[i + 1] & ~15u) | (uint32_t)((i + 1) & 15)));
          else gm = fmax3(gm, __uint_as_float(q[i]), __uint_as_float(q[i + 1]));
        }
        if (MODE == 2) { if (gm > m) m = __uint_as_float((__float_as_uint(gm) & ~0xf0u) | (uint32_t)(g & 0xf0)); }
        else m = fmaxf(m, gm);
      }
      slot = (slot + 1) % DEPTH;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = m;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

template <int MODE, int X, int DEPTH>
void run(int warps, int cols, int ctas, float* out, long long* cyc) {
  const int reps = 2000;
  bench<MODE, X, DEPTH><<<ctas, warps * 32>>>(cols, reps, out, cyc);
  cudaDeviceSynchronize();
  bench<MODE, X, DEPTH><<<ctas, warps * 32>>>(cols, reps, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per_rep = (double)mx / reps;
  const double bytes = 128.0 * cols * 4;
  printf("mode %d  x%-2d depth %d  warps %2d  cols %3d  ctas %3d : %8.1f cycles / pass  = %6.1f B/clk/SM  (%s)\n", MODE, X, DEPTH, warps,
         cols, ctas, per_rep, bytes / per_rep, cudaGetErrorString(e));
}

// tagging variants of the training epilogue: (x & ~15) | Q
template <int Q> __device__ __forceinline__ float tag_lop(uint32_t x, uint32_t mask) {      // one LOP3 (ALU pipe)
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEC;" : "=r"(d) : "r"(x), "n"(Q), "r"(mask));
  return __uint_as_float(d);
}
template <int Q> __device__ __forceinline__ float tag_fma(uint32_t x, uint32_t sixteen) {   // IMAD.HI + IMAD (FMA pipe)
  uint32_t hi, d;
  asm("mul.hi.u32 %0, %1, 268435456;" : "=r"(hi) : "r"(x));
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(sixteen), "n"(Q));
  return __uint_as_float(d);
}
// MODE 2: all LOP3, two chains.  MODE 3: odd columns through the FMA pipe.  MODE 4: all through the FMA pipe.
template <int MODE> __device__ __forceinline__ float group_max(const uint32_t* r, uint32_t mask, uint32_t sixteen) {
#define TL(Q) tag_lop<Q>(r[Q], mask)
#define TF(Q) tag_fma<Q>(r[Q], sixteen)
#define TA(Q) (MODE == 2 ? TL(Q) : MODE == 4 ? TF(Q) : ((Q & 1) ? TF(Q) : TL(Q)))
  float g0 = fmax3(TA(0), TA(1), TA(2)), g1 = fmax3(TA(8), TA(9), TA(10));
  g0 = fmax3(g0, TA(3), TA(4)); g1 = fmax3(g1, TA(11), TA(12));
  g0 = fmax3(g0, TA(5), TA(6)); g1 = fmax3(g1, TA(13), TA(14));
  return fmax3(g0, g1, fmaxf(TA(7), TA(15)));
#undef TL
#undef TF
#undef TA
}

// ---- TMEM read-back WHILE the tensor pipe is busy: warp `nwarps` of the CTA issues back-to-back 128 x N x 16 bf16 MMAs (zeroed
// shared-memory operands) into TMEM columns [256, 256 + N) while the other warps drain columns [0, cols) as above.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
template <int MODE>
__global__ void __launch_bounds__(544, 1) bench_mma(int cols, int reps, int mma_n, int mma_on, float* out, long long* cycles,
                                                    uint32_t mask = ~15u, uint32_t sixteen = 16u) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tbase;
  __shared__ volatile int stop;
  __shared__ uint64_t done_bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nread = (blockDim.x >> 5) - 1;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) stop = 0;
  if (warp == 0) tmem_alloc(&tbase, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float m = -INFINITY;
  long long t0 = 0, t1 = 0;
  if (warp == nread) {
    if (mma_on && lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(mma_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t ad = make_desc(smem_u32(smem), 2048, 128), bd = make_desc(smem_u32(smem) + 8192, 4096, 128);
      while (!stop) {
        for (int k = 0; k < 8; ++k)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase + 256), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
      }
      // every queued MMA has retired before TMEM is released
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&done_bar)) : "memory");
    }
  } else {
    const uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const int nw = nread >> 2, wq = warp >> 2;
    asm volatile("bar.sync 1, %0;" ::"r"(nread * 32) : "memory");
    t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      uint32_t ra[16], rb[16];
      int g = wq * 16;
      if (g < cols) ld16(t + g, ra);
      for (; g < cols; g += 2 * nw * 16) {
        ld_wait();
        const int g2 = g + nw * 16;
        if (g2 < cols) ld16(t + g2, rb);
        if (MODE == 0) m = fmaxf(m, __uint_as_float(ra[0]));
        else if (MODE >= 2) m = fmaxf(m, group_max<MODE>(ra, mask, sixteen));
        else { float gm = -INFINITY;
#pragma unroll
          for (int i = 0; i < 16; i += 2) gm = fmax3(gm, __uint_as_float((ra[i] & ~15u) | (uint32_t)i), __uint_as_float((ra[i + 1] & ~15u) | (uint32_t)(i + 1)));
          m = fmaxf(m, gm); }
        if (g2 < cols) {
          ld_wait();
          if (g2 + nw * 16 < cols) ld16(t + g2 + nw * 16, ra);
          if (MODE == 0) m = fmaxf(m, __uint_as_float(rb[0]));
          else if (MODE >= 2) m = fmaxf(m, group_max<MODE>(rb, mask, sixteen));
          else { float gm = -INFINITY;
#pragma unroll
            for (int i = 0; i < 16; i += 2) gm = fmax3(gm, __uint_as_float((rb[i] & ~15u) | (uint32_t)i), __uint_as_float((rb[i + 1] & ~15u) | (uint32_t)(i + 1)));
            m = fmaxf(m, gm); }
        }
      }
    }
    t1 = clock64();
    asm volatile("bar.sync 1, %0;" ::"r"(nread * 32) : "memory");
    if (threadIdx.x == 0) stop = 1;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = m;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

template <int MODE>
void run_mma(int readers, int cols, int mma_n, int mma_on, float* out, long long* cyc) {
  const int reps = 2000;
  cudaFuncSetAttribute(bench_mma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  for (int i = 0; i < 2; ++i) {
    bench_mma<MODE><<<1, (readers + 1) * 32, 48 * 1024>>>(cols, reps, mma_n, mma_on, out, cyc, ~15u, 16u);
    cudaDeviceSynchronize();
  }
  cudaError_t e = cudaGetLastError();
  long long h = 0;
  cudaMemcpy(&h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
  const double per_rep = (double)h / reps;
  printf("pipelined x16, mode %d, readers %2d, cols %3d, concurrent MMA 128x%dx16 %s : %8.1f cycles / pass = %6.1f B/clk/SM (%s)\n", MODE,
         readers, cols, mma_n, mma_on ? "ON " : "off", per_rep, 128.0 * cols * 4 / per_rep, cudaGetErrorString(e));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  printf("pass = all 128 lanes x `cols` fp32 columns read once; modes: 0 load only, 1 + fmax3, 2 + fmax3 with packed arg index\n");
  for (int warps : {4, 8, 16}) {
    for (int cols : {112, 256}) {
      run<0, 16, 1>(warps, cols, 1, out, cyc);
      run<0, 16, 2>(warps, cols, 1, out, cyc);
      run<0, 32, 1>(warps, cols, 1, out, cyc);
      run<0, 32, 2>(warps, cols, 1, out, cyc);
      run<1, 16, 2>(warps, cols, 1, out, cyc);
      run<1, 32, 2>(warps, cols, 1, out, cyc);
      run<2, 16, 2>(warps, cols, 1, out, cyc);
      run<2, 32, 2>(warps, cols, 1, out, cyc);
    }
  }
  printf("--- read-back with and without a concurrent MMA stream (depth-2 results above are void: dynamically indexed register arrays)\n");
  for (int readers : {8, 16}) {
    for (int on : {0, 1}) {
      run_mma<0>(readers, 112, 112, on, out, cyc);
      run_mma<1>(readers, 112, 112, on, out, cyc);
      run_mma<2>(readers, 112, 112, on, out, cyc);
      run_mma<3>(readers, 112, 112, on, out, cyc);
      run_mma<4>(readers, 112, 112, on, out, cyc);
      run_mma<0>(readers, 224, 112, on, out, cyc);
      run_mma<1>(readers, 224, 208, on, out, cyc);
    }
  }
  run<0, 32, 2>(8, 256, 148, out, cyc);
  run<2, 32, 2>(8, 256, 148, out, cyc);
  run<2, 16, 2>(16, 256, 148, out, cyc);
  return 0;
}
