// Floor of one Gauss-Seidel step INSIDE a CTA on sm_100a: 8 warps, a step is relaxed by one warp (rotating), hand-off through a
// shared-memory mbarrier (arrive after the store, wait before the next load) or bar.sync; the relaxing warp does the dependent
// chain of block_gs.cuh: 8 LDS -> 8 DMUL -> 3-level DADD tree -> +1 DADD -> log2(T) x (SHFL + DADD) -> DSUB -> DMUL, 2 DFMA -> STS.
//   nvcc -arch=sm_100a -O3 step.cu -o step
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n\t.reg .pred P1;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(b)), "r"(ph) : "memory");
}
template <int T, int MODE>   // MODE 0: mbarrier arrive/wait, 1: bar.sync, 2: one warp only, no hand-off (pure chain)
__global__ void __launch_bounds__(256) k(int nsteps, double* out, long long* cyc) {
  __shared__ double win[2048];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2048; i += 256) win[i] = 1.0 + 1e-9 * i;
  if (tid == 0) mb_init(&bar, 8);
  __syncthreads();
  uint32_t ph = 0;
  const double v = 0.125, d = 3.0, y = 1.0 / 3.0, b = 1.0;
  const long long t0 = clock64();
  for (int s = 0; s < nsteps; ++s) {
    const bool mine = MODE == 2 ? wid == 0 : (s & 7) == wid;
    if (mine) {
      double pr[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) pr[j] = __dmul_rn(v, win[(s * 8 + j * 37 + lane) & 2047]);
      double r = __dadd_rn(__dadd_rn(__dadd_rn(pr[0], pr[1]), __dadd_rn(pr[2], pr[3])), __dadd_rn(__dadd_rn(pr[4], pr[5]), __dadd_rn(pr[6], pr[7])));
      r = __dadd_rn(r, 0.5);
#pragma unroll
      for (int o = T / 2; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o, T);
      r = __dsub_rn(b, r);
      const double q = __dmul_rn(r, y);
      const double rem = __fma_rn(-d, q, r);
      const double q2 = __fma_rn(y, rem, q);
      if ((lane & (T - 1)) == 0) win[((s + 1) * 8 + lane) & 2047] = q2;
    }
    if (MODE == 0) {
      __syncwarp();
      if (lane == 0) mb_arrive(&bar);
      mb_wait(&bar, ph);
      ph ^= 1;
    } else if (MODE == 1) {
      __syncthreads();
    } else {
      __syncwarp();
    }
  }
  const long long t1 = clock64();
  if (tid == 0) { cyc[0] = t1 - t0; out[0] = win[5]; }
}
template <int T, int MODE> void run(const char* name, double* out, long long* cyc) {
  const int n = 4000;
  for (int rep = 0; rep < 2; ++rep) k<T, MODE><<<1, 256>>>(n, out, cyc);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-46s T=%2d  %.0f cycles/step\n", name, T, (double)c / n);
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
  run<1, 2>("chain only (one warp, no hand-off)", out, cyc);
  run<8, 2>("chain only (one warp, no hand-off)", out, cyc);
  run<32, 2>("chain only (one warp, no hand-off)", out, cyc);
  run<1, 0>("rotating warps, mbarrier arrive/wait", out, cyc);
  run<8, 0>("rotating warps, mbarrier arrive/wait", out, cyc);
  run<32, 0>("rotating warps, mbarrier arrive/wait", out, cyc);
  run<1, 1>("rotating warps, bar.sync", out, cyc);
  run<8, 1>("rotating warps, bar.sync", out, cyc);
  run<32, 1>("rotating warps, bar.sync", out, cyc);
  return 0;
}
