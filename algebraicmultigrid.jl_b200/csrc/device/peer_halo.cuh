// Halo exchange of a row-partitioned level over PEER MEMORY (NVLink 5 / NVSwitch): the rank that owns an entry stores it
// straight into the halo segment of every neighbour that needs it — no pack buffer, no NCCL kernel, no rendezvous.
//
// Every rank exports its exchangeable vectors (x, temp, res of every partitioned level) and one block of 64-bit words
// (flags) as CUDA IPC handles; the neighbours map them (cudaIpcOpenMemHandle) once, at finalize.  One exchange of a vector
// ("channel") is then three small kernels, all capturable into the whole-cycle CUDA graph:
//   push  (sender, communication stream)   waits until each neighbour has consumed the PREVIOUS exchange of this channel (ack),
//                                          gathers v[send_idx[i]] and stores it at the neighbour's halo position, and — last block
//                                          to finish — fence.sys + one release store of the exchange number into the
//                                          neighbour's flag word for (channel, me);
//   wait  (receiver, after its own push)   spins until the flag of every neighbour it receives from shows this exchange;
//   ack   (receiver, compute stream, after the last kernel that reads the halo) stores the exchange number into each
//                                          sender's ack word for (channel, me).
// Exchange numbers live in device memory (sent / expected per channel) and advance by one per exchange on every rank, so a
// captured graph replays correctly.  Every rank runs the same sequence of exchanges, so the waits are matched by
// construction; a watchdog turns a protocol error into a reported fault instead of a hang.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200amg {

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerMaxLevels = 8;
constexpr int kPeerChannels = 3;                              // x, temp, res
constexpr int kPeerWords = 2 * kPeerMaxWorld + 2;             // per channel: flag[world] | ack[world] | sent | expected
constexpr int kPeerSyncWords = kPeerMaxLevels * kPeerChannels * kPeerWords + 8;

__device__ __forceinline__ unsigned long long peer_ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool peer_spin_until(const unsigned long long* p, unsigned long long want, int* fault) {
  long long t0 = 0;
  unsigned spins = 0;
  while (peer_ld_acquire_sys(p) < want) {
    if ((++spins & 0x3ffu) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000ll) { atomicExch(fault, 2); return false; }
    }
  }
  return true;
}

struct PeerTables {                        // device arrays of one (level, channel)
  double* dst[kPeerMaxWorld];              // where my entries for rank q start inside q's halo of this channel (peer-mapped)
  unsigned long long* flag_at[kPeerMaxWorld];   // q's flag word for (channel, me)
  unsigned long long* ack_at[kPeerMaxWorld];    // q's ack word for (channel, me)
  int send_off[kPeerMaxWorld + 1];
  int recv_cnt[kPeerMaxWorld];
  int world;
};

// sync: my words of this channel (flag | ack | sent | expected); ticket: one zero-initialised word per channel
__global__ void __launch_bounds__(256) halo_push_kernel(const PeerTables* __restrict__ tab, const int* __restrict__ send_idx,
                                                        const double* __restrict__ v, unsigned long long* sync, unsigned* ticket, int* fault) {
  __shared__ PeerTables T;
  __shared__ int s_last;
  if (threadIdx.x == 0) T = *tab;
  __syncthreads();
  const int world = T.world;
  const unsigned long long sent = *reinterpret_cast<volatile unsigned long long*>(sync + 2 * kPeerMaxWorld);
  if ((int)threadIdx.x < world && T.send_off[threadIdx.x + 1] > T.send_off[threadIdx.x])   // the previous exchange of this channel was consumed
    peer_spin_until(sync + kPeerMaxWorld + threadIdx.x, sent, fault);
  __syncthreads();
  const int nsend = T.send_off[world];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nsend; i += gridDim.x * blockDim.x) {
    int q = 0;
    while (i >= T.send_off[q + 1]) ++q;
    T.dst[q][i - T.send_off[q]] = v[send_idx[i]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    if ((int)threadIdx.x < world && T.send_off[threadIdx.x + 1] > T.send_off[threadIdx.x]) {
      __threadfence_system();
      peer_st_release_sys(T.flag_at[threadIdx.x], sent + 1);
    }
    if (threadIdx.x == 0) {
      *ticket = 0u;
      *reinterpret_cast<volatile unsigned long long*>(sync + 2 * kPeerMaxWorld) = sent + 1;
    }
  }
}
__global__ void halo_wait_kernel(const PeerTables* __restrict__ tab, unsigned long long* sync, int* fault) {
  const int q = threadIdx.x;
  const unsigned long long want = *reinterpret_cast<volatile unsigned long long*>(sync + 2 * kPeerMaxWorld + 1) + 1;
  if (q < tab->world && tab->recv_cnt[q] > 0) peer_spin_until(sync + q, want, fault);
}
__global__ void halo_ack_kernel(const PeerTables* __restrict__ tab, unsigned long long* sync) {
  const int q = threadIdx.x;
  const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(sync + 2 * kPeerMaxWorld + 1) + 1;
  if (q < tab->world && tab->recv_cnt[q] > 0) peer_st_release_sys(tab->ack_at[q], e);
  __syncwarp();
  if (q == 0) *reinterpret_cast<volatile unsigned long long*>(sync + 2 * kPeerMaxWorld + 1) = e;
}

}  // namespace b200amg
