// Cross-SM hand-off latency on sm_100a: two CTAs (different SMs) bounce a counter N times.
//   mode 0: st.relaxed.gpu / ld.relaxed.gpu poll           (flag only)
//   mode 1: st.volatile / ld.volatile poll
//   mode 2: data store + __threadfence + red.relaxed, consumer polls then ld.cg data  (the counter protocol)
//   mode 3: 16-byte {data,flag,data,flag} mailbox store/poll (the LL protocol)
//   mode 4: same CTA pair inside one CLUSTER, flag in distributed shared memory (st.shared::cluster / ld.shared::cluster)
// prints ns per one-way hop.        nvcc -gencode arch=compute_100a,code=sm_100a -O3 pingpong.cu -o pingpong
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ unsigned ldr(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0,[%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void str(unsigned* p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0],%1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ldv(const unsigned* p) { unsigned v; asm volatile("ld.volatile.global.u32 %0,[%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void stv(unsigned* p, unsigned v) { asm volatile("st.volatile.global.u32 [%0],%1;" ::"l"(p), "r"(v) : "memory"); }
__global__ void pp(int mode, int n, unsigned* flags, double* data, uint4* mail, long long* out) {
  // block 0 and block gridDim.x-1 play; the blocks in between only make sure the two land on different SMs
  const int me = blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x - 1 ? 1 : -1);
  if (me < 0 || threadIdx.x != 0) return;
  unsigned* mine = flags + 64 * me;
  unsigned* other = flags + 64 * (1 - me);
  long long t0 = 0;
  for (int i = 1; i <= n; ++i) {
    if (i == 2 && me == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if ((i & 1) == me) {   // my turn to publish step i
      if (mode == 0) str(mine, i);
      else if (mode == 1) stv(mine, i);
      else if (mode == 2) { data[me] = (double)i; __threadfence(); asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(mine) : "memory"); }
      else { asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(mail + me), "r"(i), "r"(i), "r"(i), "r"(i) : "memory"); }
    } else {                // wait for the peer's step i
      if (mode == 0) { while (ldr(other) != (unsigned)i) {} }
      else if (mode == 1) { while (ldv(other) != (unsigned)i) {} }
      else if (mode == 2) { while (ldr(other) < (unsigned)((i + 1) / 2)) {} double v = __ldcg(data + (1 - me)); if (v != (double)i) out[3] = -1; }
      else { uint4 m; do { asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(m.x), "=r"(m.y), "=r"(m.z), "=r"(m.w) : "l"(mail + (1 - me)) : "memory"); } while (m.y != (unsigned)i || m.w != (unsigned)i); }
    }
  }
  if (me == 0) { long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); out[0] = t1 - t0; }
}
__global__ void __cluster_dims__(2, 1, 1) pp_cluster(int n, long long* out) {
  __shared__ unsigned flag;
  cg::cluster_group cl = cg::this_cluster();
  const unsigned me = cl.block_rank();
  if (threadIdx.x == 0) flag = 0;
  cl.sync();
  unsigned* peer = cl.map_shared_rank(&flag, 1 - me);
  if (threadIdx.x == 0) {
    long long t0 = 0;
    for (int i = 1; i <= n; ++i) {
      if (i == 2 && me == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      if ((unsigned)(i & 1) == me) { *(volatile unsigned*)peer = i; }                 // write INTO the peer's shared memory
      else { while (*(volatile unsigned*)&flag != (unsigned)i) {} }                   // spin on my own shared memory
    }
    if (me == 0) { long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); out[0] = t1 - t0; }
  }
  cl.sync();
}
int main() {
  const int n = 20001;
  unsigned* flags; double* data; uint4* mail; long long* out;
  cudaMalloc(&flags, 1024); cudaMalloc(&data, 64); cudaMalloc(&mail, 64); cudaMalloc(&out, 64);
  const char* nm[4] = {"relaxed.gpu flag", "volatile flag", "data+fence+red, poll+ld.cg", "16B mailbox (LL)"};
  for (int grid : {2, 75, 148}) {
    for (int mode = 0; mode < 4; ++mode) {
      cudaMemset(flags, 0, 1024); cudaMemset(mail, 0, 64); cudaMemset(out, 0, 64);
      pp<<<grid, 32>>>(mode, n, flags, data, mail, out);
      long long h[4]; cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
      printf("grid %3d  %-30s %7.1f ns/hop%s\n", grid, nm[mode], (double)h[0] / (n - 1), h[3] ? "  DATA MISMATCH" : "");
    }
  }
  cudaMemset(out, 0, 64);
  pp_cluster<<<2, 32>>>(n, out);
  long long h[1]; cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
  printf("cluster pair, flag in DSMEM            %7.1f ns/hop  (%s)\n", (double)h[0] / (n - 1), cudaGetErrorString(cudaGetLastError()));
  return 0;
}
