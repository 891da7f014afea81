// Gauss-Seidel / SOR on SMALL and MID-SIZE levels (up to ~260 000 rows): ONE thread-block cluster sweeps the
// level, x lives in DISTRIBUTED shared memory, the matrix arrives through a per-CTA TMA ring and wavefronts are
// separated by a dataflow hand-off through shared memory instead of an L2 round trip or a cluster-wide barrier.
//
// Why a third cluster variant: on these levels a wavefront has 5-250 rows of 30-130 entries, and there are
// 850-1700 of them per sweep, so the sweep time is (hand-off latency) x (wavefronts).  Measured on B200:
//   through L2 (gs_dataflow_kernel)                      2.2-2.5 us per wavefront
//   cluster barrier + __ldg row fetch (gs_cluster_kernel) 2.4 us   (row data one wavefront ahead is not far enough:
//                                                                   two dependent global loads are ~1.5 us)
//   one CTA, x through L2 (gs_cta_kernel<T,false>)        4.4-5.6 us
//   one CTA, x in shared memory (gs_cta_kernel<T,true>)   1.0-1.5 us but only <= 12 288 rows
// Here:
//   * x lives in the shared memory of NC = 1, 2, 4, 8 or 16 CTAs of one cluster, each value in the CTA that relaxes
//     its row; gathers are (remote) shared-memory loads, ~0.2 us instead of ~0.5 us through L2, and publishing a
//     value is a local shared-memory store;
//   * the level is cut at upload into wavefront-aligned tiles of <= 256/T rows and <= kDsmTileNnz entries
//     (rows of a tile are mutually independent); tiles are dealt round-robin over the CTAs in sweep order, and
//     every CTA streams ITS tiles (row pointers, column indices, values, b) through a kDsmStages-deep bulk-copy
//     ring, so matrix data is in shared memory several wavefronts before it is needed;
//   * every CTA keeps one mbarrier per wavefront in its OWN shared memory; a CTA that finished a tile of
//     wavefront w arrives on barrier[w] of EVERY CTA (remote mbarrier arrive, SASS SYNCS.ARRIVE.TRANS64.RED);
//     a tile of wavefront w + 1 waits (try_wait) on its local barrier[w].
//     Only CTAs that hold rows of a wavefront take part in its hand-off; nobody executes a cluster-wide barrier.
// Exact lexicographic semantics (gs! smoother.jl:73-90, sor_step! :205-221): the schedule is the level schedule
// of the symmetrised pattern, so every earlier-ordered neighbour sits in an earlier wavefront and every
// later-ordered one in a later wavefront (its old value is still in place when this row reads it).
// Deadlock freedom: the CTAs of a cluster are co-scheduled, each walks its tiles in sweep order, and the lowest
// unfinished tile only waits on finished wavefronts.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200amg {
namespace cg = cooperative_groups;

constexpr int kDsmThreads = 256;          // consumer threads (relax rows) by default; 512 is the other instantiation;
                                          // one more warp feeds the TMA ring
constexpr int kDsmMaxThreads = 512;
constexpr int kDsmTileNnz = 1024;
constexpr int kDsmTileRows = 128;   // = kDsmMaxThreads / 4: at least 4 lanes per row
constexpr int kDsmStages = 6;
constexpr int kDsmBurst = 8;        // entries per lane gathered back to back
constexpr int kDsmMaxDynSmem = 229376;   // 224 KB of dynamic shared memory (227 KB minus the static part)

struct __align__(16) DsmStage {
  double val[kDsmTileNnz + 8];
  double b[kDsmTileRows + 8];
  int col[kDsmTileNnz + 8];
  int rp[kDsmTileRows + 8];
};
static_assert(sizeof(DsmStage) % 16 == 0, "stage must keep 16-byte alignment");

// dynamic shared memory one CTA needs when the fullest CTA holds nslots values of x and the level has nlev wavefronts
static inline size_t dsm_smem_bytes(int64_t nslots_, int nlev) {
  const size_t nslots = (size_t)nslots_;
  return (size_t)kDsmStages * sizeof(DsmStage) + nslots * sizeof(double) + ((size_t)nlev + 2) * sizeof(uint64_t);
}

__device__ __forceinline__ void dsm_issue(DsmStage& S, uint64_t* bar, const int4 m, const int* __restrict__ rowptr,
                                          const int* __restrict__ col, const double* __restrict__ val,
                                          const double* __restrict__ b) {
  const int ka = m.z & ~3, kcnt = (m.w - ka + 3) & ~3;
  const int ra = m.x & ~3, rcnt = (m.y + 1 - ra + 3) & ~3;
  mbar_expect_tx(bar, (uint32_t)(kcnt * 12 + rcnt * 4 + rcnt * 8));
  if (kcnt) {
    bulk_g2s(S.val, val + ka, (uint32_t)kcnt * 8u, bar);
    bulk_g2s(S.col, col + ka, (uint32_t)kcnt * 4u, bar);
  }
  bulk_g2s(S.rp, rowptr + ra, (uint32_t)rcnt * 4u, bar);
  bulk_g2s(S.b, b + ra, (uint32_t)rcnt * 8u, bar);
}

// eight shared-memory loads issued back to back: LOCAL (ld.shared::cta, SASS LDS) when the cluster is one CTA, else
// through the cluster window (ld.shared::cluster, SASS LD.E on the shared window: local or remote shared memory)
template <bool LOCAL>
__device__ __forceinline__ void dsm_burst8(double (&out)[kDsmBurst], const uint32_t (&a)[kDsmBurst]) {
  static_assert(kDsmBurst == 8, "the burst is written for 8 slots");
  if (LOCAL)
    asm volatile(
        "ld.shared::cta.f64 %0, [%8];\n\t"
        "ld.shared::cta.f64 %1, [%9];\n\t"
        "ld.shared::cta.f64 %2, [%10];\n\t"
        "ld.shared::cta.f64 %3, [%11];\n\t"
        "ld.shared::cta.f64 %4, [%12];\n\t"
        "ld.shared::cta.f64 %5, [%13];\n\t"
        "ld.shared::cta.f64 %6, [%14];\n\t"
        "ld.shared::cta.f64 %7, [%15];"
        : "=d"(out[0]), "=d"(out[1]), "=d"(out[2]), "=d"(out[3]), "=d"(out[4]), "=d"(out[5]), "=d"(out[6]), "=d"(out[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
        : "memory");
  else
    asm volatile(
        "ld.shared::cluster.f64 %0, [%8];\n\t"
        "ld.shared::cluster.f64 %1, [%9];\n\t"
        "ld.shared::cluster.f64 %2, [%10];\n\t"
        "ld.shared::cluster.f64 %3, [%11];\n\t"
        "ld.shared::cluster.f64 %4, [%12];\n\t"
        "ld.shared::cluster.f64 %5, [%13];\n\t"
        "ld.shared::cluster.f64 %6, [%14];\n\t"
        "ld.shared::cluster.f64 %7, [%15];"
        : "=d"(out[0]), "=d"(out[1]), "=d"(out[2]), "=d"(out[3]), "=d"(out[4]), "=d"(out[5]), "=d"(out[6]), "=d"(out[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dsm_fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void dsm_fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t dsm_mapa(uint32_t addr, unsigned rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// one arrival on the mbarrier at shared::cluster address `addr` (any CTA of the cluster; SASS SYNCS.ARRIVE.TRANS64.RED).
// .relaxed: with .release.cluster ptxas emits MEMBAR.ALL.GPU in front (0.65 us measured); the caller orders its LOCAL
// shared-memory stores with bar.sync + fence.acq_rel.cta instead (see the kernel's comment).
__device__ __forceinline__ void dsm_arrive_remote(uint32_t addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
// non-blocking phase test (SASS SYNCS.PHASECHK without TRYWAIT: no hardware suspend, lowest wake-up latency)
__device__ __forceinline__ bool dsm_test_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.relaxed.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool dsm_try_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}

// Plan (built at upload for one cluster size NC, engine.cu):
//   meta[t] = {first row, end row, first nnz, end nnz}            tile t in forward sweep order, owned by CTA t % NC
//   aux[t]  = {wavefront (forward numbering), first slot of the tile}
//   code[k] = the column index of entry k rewritten as  slot * NC + owner  (where x of that column lives)
//   rowof[own_off[r] + s] = the row whose x lives in slot s of CTA r;  own_off[NC + 1] = slots of the fullest CTA
//   wave_tiles[w] = tiles of (forward) wavefront w = arrivals that complete its mbarrier
// OWNER COMPUTES: the CTA that relaxes a row also holds its x, so publishing a value is a LOCAL shared-memory store;
// only the gathers and the arrivals cross the cluster.  Hand-off: every CTA has one mbarrier per wavefront; a CTA that
// finished a tile of wavefront w arrives once on barrier[w] of EVERY CTA (remote mbarrier arrive), a tile of the next
// wavefront waits (try_wait) on its local barrier[w].  Ordering: the producer's stores are to its own shared memory;
// bar.sync + fence.acq_rel.cta put them before the arrivals, so a value is in the owner's shared memory before the
// arrival is even sent, and distributed shared memory is not cached anywhere.  flags bit 0 / bit 1 add
// fence.acq_rel.cluster on the producer / consumer side (the PTX-model-complete protocol; ptxas turns each into
// MEMBAR.ALL.GPU, +0.65 us per wavefront each, measured) for A/B runs; bit 2: wait with try_wait (hardware suspend)
// instead of spinning on test_wait.
template <int LOG_NC, int T, int BS>
__global__ void __launch_bounds__(BS + 32, 1)
    gs_dsm_kernel(int n, int ntiles, int nlev, const int4* __restrict__ meta, const int2* __restrict__ aux,
                  const int* __restrict__ rowptr, const int* __restrict__ code, const double* __restrict__ val,
                  const int* __restrict__ rowof, const int* __restrict__ own_off, const int* __restrict__ wave_tiles, double* x,
                  const double* __restrict__ b, double omega, int sor, int backward, int opaque_zero, int flags,
                  int* __restrict__ status, unsigned long long* __restrict__ dbg) {
  constexpr int NC = 1 << LOG_NC;
  constexpr bool LOCAL = NC == 1;
  constexpr int kDsmThreads = BS, kDsmBlock = BS + 32;   // (shadow the defaults)
  extern __shared__ __align__(128) unsigned char dsm_smem[];
  DsmStage* st = reinterpret_cast<DsmStage*>(dsm_smem);
  double* xs = reinterpret_cast<double*>(dsm_smem + kDsmStages * sizeof(DsmStage));
  __shared__ __align__(8) uint64_t full[kDsmStages];    // stage filled (TMA bytes landed)
  __shared__ uint32_t xaddr[NC];     // shared::cluster address of every CTA's x slots
  __shared__ uint32_t wbaddr[NC];    // shared::cluster address of every CTA's wavefront barriers
  __shared__ double s_zero;          // what idle slots of a gather burst read
  const int tid = threadIdx.x, g = tid / T, lane = tid % T;
  unsigned rank = 0;
  if (NC > 1) rank = cg::this_cluster().block_rank();
  const int off0 = __ldg(own_off + rank), nslots = __ldg(own_off + rank + 1) - off0;
  const int slots_max = __ldg(own_off + NC + 1);   // the same layout in every CTA
  uint64_t* wb = reinterpret_cast<uint64_t*>(xs + slots_max);
  const uint32_t wb0 = smem_u32(wb), zaddr = smem_u32(&s_zero);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kDsmStages; ++s) mbar_init(&full[s], 1);
    s_zero = 0.0;
  }
  if (NC > 1)
    for (int i = tid; i < nlev; i += kDsmBlock) mbar_init(&wb[i], (uint32_t)__ldg(wave_tiles + i));
  mbar_fence_init();
  __syncthreads();
  const int nown = ntiles > (int)rank ? (ntiles - (int)rank + NC - 1) >> LOG_NC : 0;   // my tiles: rank, rank + NC, ...
  auto tile_of = [&](int i) { return (int)rank + ((backward ? nown - 1 - i : i) << LOG_NC); };
#pragma unroll 4
  for (int i = tid; i < nslots; i += kDsmBlock) xs[i] = __ldcg(x + __ldg(rowof + off0 + i));
  if (tid < NC) {
    xaddr[tid] = NC > 1 ? dsm_mapa(smem_u32(xs), (unsigned)tid) : smem_u32(xs);
    wbaddr[tid] = NC > 1 ? dsm_mapa(wb0, (unsigned)tid) : wb0;
  }
  __syncthreads();
  if (NC > 1) cg::this_cluster().sync();   // every CTA's x slots and barriers are initialised before anyone touches them

  if (tid >= kDsmThreads) {
    // ---- producer warp: keeps the ring full; nothing of the matrix stream is on the consumers' critical path ----
    {
      int4 mt = nown > 0 ? __ldg(meta + tile_of(0)) : make_int4(0, 0, 0, 0);
      int ps = 0;
      for (int i = 0; i < nown; ++i) {
        const int4 mn = i + 1 < nown ? __ldg(meta + tile_of(i + 1)) : mt;
        if (i >= kDsmStages) {   // the consumers released the stage: bar.arrive there, bar.sync here (named barrier 2 + stage)
          asm volatile("bar.sync %0, %1;" ::"r"(2 + ps), "n"(kDsmBlock) : "memory");
        }
        if (tid == kDsmThreads) dsm_issue(st[ps], &full[ps], mt, rowptr, code, val, b);
        mt = mn;
        if (++ps == kDsmStages) ps = 0;
      }
    }
  } else {
  int s = 0;
  uint32_t parity = 0;
  bool dead = false;
  int4 m = make_int4(0, 0, 0, 0);
  int2 au = make_int2(0, 0);
  if (nown > 0) { m = __ldg(meta + tile_of(0)); au = __ldg(aux + tile_of(0)); }
  for (int i = 0; i < nown; ++i) {
    int4 m_next = m;
    int2 au_next = au;
    if (i + 1 < nown) { m_next = __ldg(meta + tile_of(i + 1)); au_next = __ldg(aux + tile_of(i + 1)); }   // off the critical path
    unsigned long long* stamp = (dbg && tid == 0) ? dbg + 8 * (size_t)tile_of(i) : nullptr;   // diagnostics (gs_timeline)
    if (stamp) stamp[0] = (unsigned long long)clock64();
    const int wf = au.x;                          // forward wavefront number
    const int wprev = backward ? wf + 1 : wf - 1;   // the wavefront the sweep relaxed before this one
    const int ka = m.z & ~3, ra = m.x & ~3;
    const int nrows = m.y - m.x;
    const bool active = g < nrows;
    const int row = active ? m.x + g : -1;
    const int mycode = active ? (((au.y + g) << LOG_NC) | (int)rank) : -1;
    mbar_wait(&full[s], parity);
    const DsmStage& S = st[s];
    int ks = 0, ke = 0;
    if (active) {
      ks = S.rp[row - ra] - ka;
      ke = S.rp[row - ra + 1] - ka;
    }
    int c[kDsmBurst];
    double v[kDsmBurst];
    uint32_t a[kDsmBurst];
#pragma unroll
    for (int j = 0; j < kDsmBurst; ++j) {
      const int k = ks + lane + j * T;
      const bool in = k < ke;
      c[j] = in ? S.col[k] : -1;
      v[j] = in ? S.val[k] : 0.0;
      a[j] = (c[j] >= 0 && c[j] != mycode) ? xaddr[c[j] & (NC - 1)] + 8u * (uint32_t)(c[j] >> LOG_NC) : zaddr;
    }
    const double bi = (active && lane == 0) ? S.b[row - ra] : 0.0;
    if (stamp) stamp[1] = (unsigned long long)clock64();
    // ---- wait until the previous wavefront of the sweep has been relaxed everywhere ----
    // (one CTA: the bar.sync that closed the previous tile already is the hand-off)
    if (NC > 1 && wprev >= 0 && wprev < nlev && !dead) {
      const uint32_t ba = wb0 + 8u * (uint32_t)wprev;
      long long t0 = 0;
      unsigned spins = 0;
      while (!((flags & 4) ? dsm_try_wait(ba, 0u) : dsm_test_wait(ba, 0u))) {
        if ((++spins & 0x3ffu) == 0u) {   // watchdog: a protocol error must not hang the device
          const long long now = clock64();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 6000000000ll) { dead = true; if (status) atomicExch(status, 1); break; }
        }
      }
    }
    if (NC > 1 && (flags & 2)) dsm_fence_cluster();
    if (stamp) stamp[2] = (unsigned long long)clock64();
    // ---- gather: all entries of the row in bursts of kDsmBurst per lane, two partial sums per lane ----
    double rs0 = 0.0, rs1 = 0.0, d = 0.0;
    {
      double xn[kDsmBurst];
      dsm_burst8<LOCAL>(xn, a);
      pin_burst8(xn, opaque_zero);
      if (stamp) stamp[3] = (unsigned long long)clock64() + (unsigned long long)(__double2loint(xn[0]) & opaque_zero);
#pragma unroll
      for (int j = 0; j < kDsmBurst; j += 2) {
        if (c[j] == mycode && c[j] >= 0) d = v[j];
        else rs0 = __dadd_rn(rs0, __dmul_rn(v[j], xn[j]));
        if (c[j + 1] == mycode && c[j + 1] >= 0) d = v[j + 1];
        else rs1 = __dadd_rn(rs1, __dmul_rn(v[j + 1], xn[j + 1]));
      }
    }
    for (int k0 = ks + kDsmBurst * T; k0 < ke; k0 += kDsmBurst * T) {   // rows longer than kDsmBurst * T entries
      double xn[kDsmBurst];
#pragma unroll
      for (int j = 0; j < kDsmBurst; ++j) {
        const int k = k0 + lane + j * T;
        const bool in = k < ke;
        c[j] = in ? S.col[k] : -1;
        v[j] = in ? S.val[k] : 0.0;
        a[j] = (c[j] >= 0 && c[j] != mycode) ? xaddr[c[j] & (NC - 1)] + 8u * (uint32_t)(c[j] >> LOG_NC) : zaddr;
      }
      dsm_burst8<LOCAL>(xn, a);
      pin_burst8(xn, opaque_zero);
#pragma unroll
      for (int j = 0; j < kDsmBurst; j += 2) {
        if (c[j] == mycode && c[j] >= 0) d = v[j];
        else rs0 = __dadd_rn(rs0, __dmul_rn(v[j], xn[j]));
        if (c[j + 1] == mycode && c[j + 1] >= 0) d = v[j + 1];
        else rs1 = __dadd_rn(rs1, __dmul_rn(v[j + 1], xn[j + 1]));
      }
    }
    double rsum = __dadd_rn(rs0, rs1);
    __syncwarp();   // lane groups of a warp may have walked different numbers of bursts
    if (T > 1) {
      rsum = group_lanes_sum<T>(rsum, 0xffffffffu);
      d = group_lanes_sum<T>(d, 0xffffffffu);
    }
    if (stamp) stamp[4] = (unsigned long long)clock64() + (unsigned long long)(__double2loint(rsum) & opaque_zero);
    if (active && lane == 0 && d != 0.0) {
      volatile double* slot = xs + (au.y + g);   // my own shared memory
      const double r = __dsub_rn(bi, rsum);
      *slot = sor ? __dadd_rn(__dmul_rn(1.0 - omega, *slot), __dmul_rn(__ddiv_rn(omega, d), r)) : __ddiv_rn(r, d);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kDsmThreads) : "memory");   // consumers: the tile is relaxed (its x is in my shared memory), stage s is free
    if (i + kDsmStages < nown) asm volatile("bar.arrive %0, %1;" ::"r"(2 + s), "n"(kDsmBlock) : "memory");   // tell the producer warp (non-blocking)
    if (stamp) stamp[5] = (unsigned long long)clock64();
    if (tid < 32) {
      if (NC > 1) {
        if (flags & 1) dsm_fence_cluster();
        else dsm_fence_cta();
        if (tid < NC) dsm_arrive_remote(wbaddr[tid] + 8u * (uint32_t)wf);
      }
      if (stamp) { stamp[6] = (unsigned long long)clock64(); stamp[7] = global_ns(); }
    }
    if (++s == kDsmStages) { s = 0; parity ^= 1u; }
    m = m_next;
    au = au_next;
  }
  }   // consumers
  // nobody may leave (and release its shared memory) while others still gather from it
  if (NC > 1) cg::this_cluster().sync();
  else __syncthreads();
#pragma unroll 4
  for (int i = tid; i < nslots; i += kDsmBlock) x[__ldg(rowof + off0 + i)] = xs[i];
}


// =============================================================================================
// gs_dsm2_kernel: the same sweep on the same plan with TWO consumer groups that alternate the tiles of a CTA.
// In gs_dsm_kernel everything between two wavefront hand-offs is serial in the same warps: stage wait, row pointers,
// column codes / values out of the ring, gather addresses (a table lookup per entry), diagonal search — and only then the
// gather, the sum, the division and the store that the NEXT wavefront is waiting for.  Here the group that owns tile
// i + 1 does all of that PREPARATION (including the lane reduction that finds the diagonal) while the other group relaxes
// tile i; what remains between two hand-offs is: gather burst -> products -> lane reduction -> division -> store.
//   one CTA    hand-off = split named barrier: the relaxing group `bar.arrive`s right after its stores and goes on to prepare
//              its next tile, the other group `bar.sync`s right before its gather (PTX producer / consumer pattern);
//   NC > 1     hand-off = the per-wavefront mbarriers of gs_dsm_kernel (remote arrive by the first warp of the group after
//              a group barrier); tiles of one CTA need no extra ordering: a tile of wavefront w waits for ALL tiles of
//              wavefront w - 1, local ones included.
// Stage s of the ring is always used by group s % 2 (kDsmStages is even), so the stage-release barrier 2 + s pairs that
// group with the producer warp.  Named barriers: 0 __syncthreads, 2..7 stage release, 8 / 9 group hand-off, 10 / 11 group
// barrier (NC > 1).
// =============================================================================================
static_assert(kDsmStages % 2 == 0, "gs_dsm2_kernel pairs stage parity with the consumer group");

template <int LOG_NC, int T, int BS>
__global__ void __launch_bounds__(2 * BS + 32, 1)
    gs_dsm2_kernel(int n, int ntiles, int nlev, const int4* __restrict__ meta, const int2* __restrict__ aux,
                   const int* __restrict__ rowptr, const int* __restrict__ code, const double* __restrict__ val,
                   const int* __restrict__ rowof, const int* __restrict__ own_off, const int* __restrict__ wave_tiles, double* x,
                   const double* __restrict__ b, double omega, int sor, int backward, int opaque_zero, int flags,
                   int* __restrict__ status, unsigned long long* __restrict__ dbg) {
  constexpr int NC = 1 << LOG_NC;
  constexpr bool LOCAL = NC == 1;
  constexpr int kCons = 2 * BS, kBlock = 2 * BS + 32, kRelease = BS + 32;
  extern __shared__ __align__(128) unsigned char dsm_smem[];
  DsmStage* st = reinterpret_cast<DsmStage*>(dsm_smem);
  double* xs = reinterpret_cast<double*>(dsm_smem + kDsmStages * sizeof(DsmStage));
  __shared__ __align__(8) uint64_t full[kDsmStages];    // stage filled (TMA bytes landed)
  __shared__ uint32_t xaddr[NC];     // shared::cluster address of every CTA's x slots
  __shared__ uint32_t wbaddr[NC];    // shared::cluster address of every CTA's wavefront barriers
  __shared__ double s_zero;          // what idle slots of a gather burst read
  const int tid = threadIdx.x;
  const int grp = tid / BS;          // 0 / 1: consumer groups, 2: producer warp
  const int gt = tid - grp * BS, g = gt / T, lane = gt % T;
  unsigned rank = 0;
  if (NC > 1) rank = cg::this_cluster().block_rank();
  const int off0 = __ldg(own_off + rank), nslots = __ldg(own_off + rank + 1) - off0;
  const int slots_max = __ldg(own_off + NC + 1);   // the same layout in every CTA
  uint64_t* wb = reinterpret_cast<uint64_t*>(xs + slots_max);
  const uint32_t wb0 = smem_u32(wb), zaddr = smem_u32(&s_zero);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kDsmStages; ++s) mbar_init(&full[s], 1);
    s_zero = 0.0;
  }
  if (NC > 1)
    for (int i = tid; i < nlev; i += kBlock) mbar_init(&wb[i], (uint32_t)__ldg(wave_tiles + i));
  mbar_fence_init();
  __syncthreads();
  const int nown = ntiles > (int)rank ? (ntiles - (int)rank + NC - 1) >> LOG_NC : 0;   // my tiles: rank, rank + NC, ...
  auto tile_of = [&](int i) { return (int)rank + ((backward ? nown - 1 - i : i) << LOG_NC); };
#pragma unroll 4
  for (int i = tid; i < nslots; i += kBlock) xs[i] = __ldcg(x + __ldg(rowof + off0 + i));
  if (tid < NC) {
    xaddr[tid] = NC > 1 ? dsm_mapa(smem_u32(xs), (unsigned)tid) : smem_u32(xs);
    wbaddr[tid] = NC > 1 ? dsm_mapa(wb0, (unsigned)tid) : wb0;
  }
  __syncthreads();
  if (NC > 1) cg::this_cluster().sync();   // every CTA's x slots and barriers are initialised before anyone touches them

  if (tid >= kCons) {
    // ---- producer warp: keeps the ring full ----
    int4 mt = nown > 0 ? __ldg(meta + tile_of(0)) : make_int4(0, 0, 0, 0);
    int ps = 0;
    for (int i = 0; i < nown; ++i) {
      const int4 mn = i + 1 < nown ? __ldg(meta + tile_of(i + 1)) : mt;
      if (i >= kDsmStages) asm volatile("bar.sync %0, %1;" ::"r"(2 + ps), "n"(kRelease) : "memory");   // released by the stage's group
      if (tid == kCons) dsm_issue(st[ps], &full[ps], mt, rowptr, code, val, b);
      mt = mn;
      if (++ps == kDsmStages) ps = 0;
    }
  } else {
    bool dead = false;
    int4 m = make_int4(0, 0, 0, 0);
    int2 au = make_int2(0, 0);
    if (grp < nown) { m = __ldg(meta + tile_of(grp)); au = __ldg(aux + tile_of(grp)); }
    for (int i = grp; i < nown; i += 2) {
      const int s = i % kDsmStages;
      const uint32_t parity = (uint32_t)(i / kDsmStages) & 1u;
      int4 m_next = m;
      int2 au_next = au;
      if (i + 2 < nown) { m_next = __ldg(meta + tile_of(i + 2)); au_next = __ldg(aux + tile_of(i + 2)); }
      unsigned long long* stamp = (dbg && gt == 0) ? dbg + 8 * (size_t)tile_of(i) : nullptr;   // diagnostics (gs_timeline)
      if (stamp) stamp[0] = (unsigned long long)clock64();
      // ---------------- PREPARE (overlaps the other group's relaxation) ----------------
      const int wf = au.x;                            // forward wavefront number
      const int wprev = backward ? wf + 1 : wf - 1;   // the wavefront the sweep relaxed before this one
      const int ka = m.z & ~3, ra = m.x & ~3;
      const int nrows = m.y - m.x;
      const bool active = g < nrows;
      const int row = active ? m.x + g : -1;
      const int mycode = active ? (((au.y + g) << LOG_NC) | (int)rank) : -1;
      mbar_wait(&full[s], parity);
      const DsmStage& S = st[s];
      int ks = 0, ke = 0;
      if (active) {
        ks = S.rp[row - ra] - ka;
        ke = S.rp[row - ra + 1] - ka;
      }
      double v[kDsmBurst];
      uint32_t a[kDsmBurst];
      double d = 0.0;
#pragma unroll
      for (int j = 0; j < kDsmBurst; ++j) {
        const int k = ks + lane + j * T;
        const bool in = k < ke;
        const int c = in ? S.col[k] : -1;
        const double vv = in ? S.val[k] : 0.0;
        const bool diag = c == mycode && c >= 0;
        if (diag) d = vv;
        const bool off = c >= 0 && !diag;
        v[j] = off ? vv : 0.0;                      // the diagonal and idle slots contribute 0 * 0
        a[j] = off ? xaddr[c & (NC - 1)] + 8u * (uint32_t)(c >> LOG_NC) : zaddr;
      }
      const bool long_row = ke - ks > kDsmBurst * T;   // entries beyond the first burst are handled after the hand-off
      const bool any_long = __any_sync(0xffffffffu, long_row);
      if (T > 1 && !any_long) d = group_lanes_sum<T>(d, 0xffffffffu);   // the diagonal is known before the hand-off
      const double bi = (active && lane == 0) ? S.b[row - ra] : 0.0;
      volatile double* slot = xs + (au.y + (active ? g : 0));   // my own shared memory
      if (stamp) stamp[1] = (unsigned long long)clock64();
      // ---------------- WAIT: the previous tile / wavefront has been relaxed ----------------
      if (NC == 1) {
        if (i > 0) asm volatile("bar.sync %0, %1;" ::"r"(8 + ((i - 1) & 1)), "n"(kCons) : "memory");
      } else if (wprev >= 0 && wprev < nlev && !dead) {
        const uint32_t ba = wb0 + 8u * (uint32_t)wprev;
        long long t0 = 0;
        unsigned spins = 0;
        while (!((flags & 4) ? dsm_try_wait(ba, 0u) : dsm_test_wait(ba, 0u))) {
          if ((++spins & 0x3ffu) == 0u) {   // watchdog: a protocol error must not hang the device
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 6000000000ll) { dead = true; if (status) atomicExch(status, 1); break; }
          }
        }
        if (flags & 2) dsm_fence_cluster();
      }
      if (stamp) stamp[2] = (unsigned long long)clock64();
      // ---------------- RELAX: gather burst, two partial sums per lane, lane reduction, division, store ----------------
      double rs0 = 0.0, rs1 = 0.0;
      {
        double xn[kDsmBurst];
        dsm_burst8<LOCAL>(xn, a);
        pin_burst8(xn, opaque_zero);
        if (stamp) stamp[3] = (unsigned long long)clock64() + (unsigned long long)(__double2loint(xn[0]) & opaque_zero);
#pragma unroll
        for (int j = 0; j < kDsmBurst; j += 2) {
          rs0 = __dadd_rn(rs0, __dmul_rn(v[j], xn[j]));
          rs1 = __dadd_rn(rs1, __dmul_rn(v[j + 1], xn[j + 1]));
        }
      }
      if (any_long) {
        for (int k0 = ks + kDsmBurst * T; k0 < ke; k0 += kDsmBurst * T) {   // rows longer than kDsmBurst * T entries
          double xn[kDsmBurst];
#pragma unroll
          for (int j = 0; j < kDsmBurst; ++j) {
            const int k = k0 + lane + j * T;
            const bool in = k < ke;
            const int c = in ? S.col[k] : -1;
            const double vv = in ? S.val[k] : 0.0;
            const bool diag = c == mycode && c >= 0;
            if (diag) d = vv;
            const bool off = c >= 0 && !diag;
            v[j] = off ? vv : 0.0;
            a[j] = off ? xaddr[c & (NC - 1)] + 8u * (uint32_t)(c >> LOG_NC) : zaddr;
          }
          dsm_burst8<LOCAL>(xn, a);
          pin_burst8(xn, opaque_zero);
#pragma unroll
          for (int j = 0; j < kDsmBurst; j += 2) {
            rs0 = __dadd_rn(rs0, __dmul_rn(v[j], xn[j]));
            rs1 = __dadd_rn(rs1, __dmul_rn(v[j + 1], xn[j + 1]));
          }
        }
        __syncwarp();   // lane groups of a warp may have walked different numbers of bursts
        if (T > 1) d = group_lanes_sum<T>(d, 0xffffffffu);
      }
      double rsum = __dadd_rn(rs0, rs1);
      if (T > 1) rsum = group_lanes_sum<T>(rsum, 0xffffffffu);
      if (stamp) stamp[4] = (unsigned long long)clock64() + (unsigned long long)(__double2loint(rsum) & opaque_zero);
      if (active && lane == 0 && d != 0.0) {
        const double r = __dsub_rn(bi, rsum);
        *slot = sor ? __dadd_rn(__dmul_rn(1.0 - omega, *slot), __dmul_rn(__ddiv_rn(omega, d), r)) : __ddiv_rn(r, d);
      }
      // ---------------- SIGNAL ----------------
      if (NC == 1) {
        if (i + 1 < nown) asm volatile("bar.arrive %0, %1;" ::"r"(8 + (i & 1)), "n"(kCons) : "memory");
      } else {
        asm volatile("bar.sync %0, %1;" ::"r"(10 + grp), "n"(BS) : "memory");   // the tile's x is in my shared memory
        if (gt < 32) {
          if (flags & 1) dsm_fence_cluster();
          else dsm_fence_cta();
          if (gt < NC) dsm_arrive_remote(wbaddr[gt] + 8u * (uint32_t)wf);
        }
      }
      if (stamp) { stamp[5] = (unsigned long long)clock64(); stamp[6] = stamp[5]; stamp[7] = global_ns(); }
      if (i + kDsmStages < nown) asm volatile("bar.arrive %0, %1;" ::"r"(2 + s), "n"(kRelease) : "memory");   // stage s is free
      m = m_next;
      au = au_next;
    }
  }   // consumers
  // nobody may leave (and release its shared memory) while others still gather from it
  if (NC > 1) cg::this_cluster().sync();
  else __syncthreads();
#pragma unroll 4
  for (int i = tid; i < nslots; i += kBlock) x[__ldg(rowof + off0 + i)] = xs[i];
}

}  // namespace b200amg
