// Gauss-Seidel / SOR on MID-SIZE levels (tens to hundreds of thousands of rows, 30-130 entries per row, wavefronts
// of 20-250 rows): ONE thread-block cluster sweeps the level.
//
// Why: such a level has ~1000-1700 dependent wavefronts; across SMs through L2 every wavefront costs 2.2-2.5 us
// (measured, gs_dataflow_kernel), inside one SM 1.0-1.5 us but x no longer fits one SM's shared memory.  A cluster
// of 8 or 16 CTAs keeps the whole x vector in DISTRIBUTED shared memory (x[p] lives in CTA p % NC, slot p / NC),
// exchanges it with remote shared-memory loads/stores, and separates wavefronts with the hardware cluster barrier
// (barrier.cluster arrive.release / wait.acquire) instead of an L2 hand-off (0.19 us vs 0.5-1.3 us per hop,
// tools/micro/pingpong.cu).  One warp relaxes one row (32 lanes, <= kClPrefetch entries each in registers); the
// row data of the NEXT wavefront is requested before the barrier, so a hop is barrier + DSMEM gather + arithmetic.
// Same exact lexicographic semantics as the other sweeps (gs! smoother.jl:73-90, sor_step! :205-221).
//
// MEASURED (256^3 RS hierarchy, level 5: 38 260 rows, 1702 wavefronts): 2.4 us per wavefront with 8 CTAs x 256 threads,
// 3.0-3.5 us with 1024 threads per CTA — no better than the wavefront-counter sweep through L2 (2.2 us), so the engine
// does not select this kernel by default (B200AMG_OPT_GS_CLUSTER turns it on); it stays as a tested alternative.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200amg {
namespace cg = cooperative_groups;

constexpr int kClThreads = 1024;
constexpr int kClPrefetch = 4;   // entries per lane held in registers: rows up to 128 entries are fully prefetched

struct ClRow {   // one warp's next row, lane-private parts
  int row, ks, ke;
  int c[kClPrefetch];
  double v[kClPrefetch];
  double b, xold;
};

__device__ __forceinline__ void cl_fetch(ClRow& r, int row, int lane, const int* __restrict__ rowptr, const int* __restrict__ col,
                                         const double* __restrict__ val, const double* __restrict__ b) {
  r.row = row;
  r.ks = r.ke = 0;
  r.b = 0.0;
  if (row >= 0) {
    r.ks = __ldg(rowptr + row);
    r.ke = __ldg(rowptr + row + 1);
    if (lane == 0) r.b = __ldg(b + row);
  }
#pragma unroll
  for (int j = 0; j < kClPrefetch; ++j) {
    const int k = r.ks + lane + 32 * j;
    const bool in = k < r.ke;
    r.c[j] = in ? __ldg(col + k) : -1;
    r.v[j] = in ? __ldg(val + k) : 0.0;
  }
}

// LOG_NC: log2 of the cluster size (x[p] lives in CTA p & (NC-1), slot p >> LOG_NC)
template <int LOG_NC, int BS>
__global__ void __launch_bounds__(BS, 1)
    gs_cluster_kernel(int n, int nlev, const int* __restrict__ lvlptr, const int* __restrict__ rowptr, const int* __restrict__ col,
                      const double* __restrict__ val, double* x, const double* __restrict__ b, double omega, int sor, int backward) {
  constexpr int NC = 1 << LOG_NC;
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();
  extern __shared__ __align__(16) double cl_xs[];   // my slots of x, then the wavefront boundaries
  const int nslots = (n + NC - 1) >> LOG_NC;
  int* slv = reinterpret_cast<int*>(cl_xs + nslots);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; (i << LOG_NC) + rank < n; i += BS) cl_xs[i] = __ldcg(x + (i << LOG_NC) + rank);
  for (int i = tid; i <= nlev; i += BS) slv[i] = __ldg(lvlptr + i);
  // remote bases of every CTA's x slice (generic addresses into distributed shared memory)
  __shared__ double* base[NC];
  if (tid < NC) base[tid] = cl.map_shared_rank(cl_xs, tid);
  __syncthreads();
  cl.sync();
  constexpr int WPC = BS / 32;          // warps per CTA
  const int gw = rank * WPC + warp;    // warp id across the cluster
  constexpr int NW = NC * WPC;
  auto wave_lo = [&](int ww) { const int w = backward ? nlev - 1 - ww : ww; return slv[w]; };
  auto wave_hi = [&](int ww) { const int w = backward ? nlev - 1 - ww : ww; return slv[w + 1]; };
  ClRow nxt;
  {
    const int r0 = wave_lo(0) + gw;
    cl_fetch(nxt, r0 < wave_hi(0) ? r0 : -1, lane, rowptr, col, val, b);
  }
  for (int ww = 0; ww < nlev; ++ww) {
    const int a = wave_lo(ww), e = wave_hi(ww);
    ClRow cur = nxt;
    // request the first row of the next wavefront now: its latency hides behind this wavefront and the barrier
    if (ww + 1 < nlev) {
      const int r1 = wave_lo(ww + 1) + gw;
      cl_fetch(nxt, r1 < wave_hi(ww + 1) ? r1 : -1, lane, rowptr, col, val, b);
    }
    for (int row = a + gw; row < e; row += NW) {
      if (row != cur.row) cl_fetch(cur, row, lane, rowptr, col, val, b);   // wide wavefront: more than one row per warp
      double rsum = 0.0, d = 0.0;
#pragma unroll
      for (int j = 0; j < kClPrefetch; ++j) {
        const int c = cur.c[j];
        if (c == row) d = cur.v[j];
        else if (c >= 0) rsum = __dadd_rn(rsum, __dmul_rn(cur.v[j], base[c & (NC - 1)][c >> LOG_NC]));
      }
      for (int k = cur.ks + lane + 32 * kClPrefetch; k < cur.ke; k += 32) {   // rows longer than 128 entries
        const int c = __ldg(col + k);
        const double v = __ldg(val + k);
        if (c == row) d = v;
        else rsum = __dadd_rn(rsum, __dmul_rn(v, base[c & (NC - 1)][c >> LOG_NC]));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        rsum += __shfl_down_sync(0xffffffffu, rsum, o);
        d += __shfl_down_sync(0xffffffffu, d, o);
      }
      if (lane == 0 && d != 0.0) {
        double* slot = base[row & (NC - 1)] + (row >> LOG_NC);
        const double r = __dsub_rn(cur.b, rsum);
        *slot = sor ? __dadd_rn(__dmul_rn(1.0 - omega, *slot), __dmul_rn(__ddiv_rn(omega, d), r)) : __ddiv_rn(r, d);
      }
    }
    cl.sync();   // the wavefront's x (remote shared-memory stores) is visible cluster-wide
  }
  for (int i = tid; (i << LOG_NC) + rank < n; i += BS) x[(i << LOG_NC) + rank] = cl_xs[i];
}

}  // namespace b200amg
