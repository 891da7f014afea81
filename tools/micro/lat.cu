// Dependent-issue latencies that bound a Gauss-Seidel hop on sm_100a: FP64 add / mul / fma / div chains,
// double shuffle, L2 load (ld.cg) — one warp, cycles per operation.   nvcc -arch=sm_100a -O3 lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, const double* buf, int n) {
  double a = out[0], b = out[1];
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) a = __dadd_rn(a, b);
  long long t1 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) a = __dmul_rn(a, b);
  long long t2 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) a = __fma_rn(a, b, b);
  long long t3 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) a = __ddiv_rn(b, a + 1.5);
  long long t4 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) a += __shfl_down_sync(0xffffffffu, a, 1);
  long long t5 = clock64();
  const double* p = buf;
  long long idx = 0;
#pragma unroll 1
  for (int i = 0; i < n; ++i) { double v = __ldcg(p + idx); idx = (long long)v; a += v; }
  long long t6 = clock64();
  out[2] = a;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; }
}
int main() {
  const int n = 2000, N = 1 << 22;
  double *out, *buf; long long* cyc;
  cudaMalloc(&out, 64); cudaMalloc(&cyc, 64); cudaMalloc(&buf, sizeof(double) * N);
  double h[3] = {1.0000001, 1.0000002, 0};
  cudaMemcpy(out, h, 24, cudaMemcpyHostToDevice);
  double* hb = new double[N];
  for (int i = 0; i < N; ++i) hb[i] = (double)((i * 7919LL + 4099) % N);   // pointer chase through 32 MB (L2 resident)
  cudaMemcpy(buf, hb, sizeof(double) * N, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(out, cyc, buf, n);
  long long c[6]; cudaMemcpy(c, cyc, 48, cudaMemcpyDeviceToHost);
  const char* nm[6] = {"DADD", "DMUL", "DFMA", "DDIV(+DADD)", "SHFL+DADD", "ld.cg L2 chase(+cvt,DADD)"};
  for (int i = 0; i < 6; ++i) printf("%-28s %.1f cycles/op\n", nm[i], (double)c[i] / n);
  return 0;
}
