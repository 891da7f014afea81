// Gauss-Seidel / SOR with the reference's exact sweep order (gs! src/smoother.jl:73-90, sor_step! :205-221) as a PASS sweep:
// the tiles, renumbering and ticket order of the blocked plan (block_plan.h), relaxed by a kernel whose hand-off between
// dependent steps is one named barrier and whose arithmetic runs out of registers.  Layout and requirements: pass_plan.h.
//
// One CTA per SM relaxes one tile at a time.  Inside the CTA:
//   * TWO compute groups of 8 warps take the passes of the tile in turn (pass i -> group i & 1).  While one group relaxes
//     pass i — window loads, four products, the sum, the lane reduction, the division, the stores: the dependent chain,
//     nothing else — the other group prepares ITS next pass.  The hand-off is a split named barrier: the relaxing group
//     `bar.arrive`s right after its stores and goes on preparing, the other group `bar.sync`s right before its window loads.
//   * A pass is prepared in two stages, each two of the group's own passes ahead of the next: FAR (global loads of the x
//     values that do not come from the window, issued ~4 passes before they are used, results left in registers) and SUM
//     (those values times their matrix entries, summed; the near entries' values and shared-memory addresses put into
//     registers; right-hand side, diagonal, reciprocal).
//   * The matrix arrives as dense slabs (one word per thread and slot, consecutive threads consecutive words), moved by ONE
//     producer thread with bulk copies (TMA) into a byte ring of whole passes (up to 16 in flight: the ring has to cover the
//     passes the groups work on plus ~2 us of copy latency); b and the diagonal ride along.
//   * Warp 7 of each group relaxes nothing: it takes part in the barriers and, having no memory operation in flight, publishes
//     the number of finished passes to shared memory with a release store (a working warp would stall on its own far loads).
//   * A gate warp waits for the passes of OTHER tiles a pass depends on (32 passes polled at once, released in order) and
//     forwards the tile's own progress to global memory (fence.acq_rel.gpu + one store); the compute threads only check a
//     shared-memory counter before they issue far loads.
//
// What a far load may see.  FAR(q) of pass q is issued by its group after the group's bar.sync for pass q - 5, so every pass
// <= q - 5 of the tile is complete and visible (st.global.cg / ld.global.cg, both at L2, ordered by the CTA barrier); the plan
// marks values of passes q - 4 .. q - 1 NEAR (window).  Old values of later-ordered neighbours cannot have been overwritten:
// their rows depend on this row (symmetric pattern) and are relaxed after it — in this tile by the barrier chain, in another
// tile by that tile's requirement on this pass.
// Deadlock freedom: as in block_gs.cuh (tickets in a topological order of the tile graph, all CTAs resident, a CTA finishes
// its tile before it claims the next); a group that waits at the gate only delays its own tile.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "block_gs.cuh"
#include "pass_plan.h"

namespace b200amg {

constexpr int kPgCompute = 2 * kPassGroup;
constexpr int kPgThreads = kPgCompute + 64;   // + producer warp, gate / publisher warp
constexpr int kPgSlots = 16;                  // passes in flight in the ring (power of two)
// one pass in the ring: values | indices | b | diagonal (entries = (far + near slots) * width; rows padded to an even count)
// (kPgRingBytes, kPgWinOff, kPgZeroOff: pass_plan.h — the near slots of the plan hold shared-memory byte offsets)
constexpr int kPgSmemBytes = kPgZeroOff + 16;
static_assert(kPgSmemBytes + 1024 <= 232448, "ring + window exceed the shared memory of one SM");
static_assert((kPgSlots & (kPgSlots - 1)) == 0, "slot count must be a power of two");
static_assert(3 * ((kPassFar + kPassNearSlots) * kPassGroup * 12 + 2 * (kPassRows + 2) * 8 + 128) <= kPgRingBytes, "ring too small for three full passes");

__device__ __forceinline__ void pg_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kPgCompute) : "memory"); }
__device__ __forceinline__ void pg_bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kPgCompute) : "memory"); }

struct PgWork {    // what a thread holds for the row it relaxes next
  double v[kPassNearSlots];      // values of the near slots
  uint32_t a[kPassNearSlots];    // shared-memory addresses of their x values
  double farsum, d, ry, bval, xold;
  int row;
  bool warp_active, store;
};

template <int T>
__global__ void __launch_bounds__(kPgThreads, 1)
    gs_pass_kernel(int ntiles, const int2* __restrict__ ptile, const int4* __restrict__ ppass, const int2* __restrict__ preq,
                   const int2* __restrict__ req, const int* __restrict__ order, unsigned* ctl,
                   const double* __restrict__ val, const int* __restrict__ idx, const double* __restrict__ diag, double* x,
                   const double* __restrict__ b, double omega, int sor, int* __restrict__ fault, unsigned long long* __restrict__ dbg) {
  extern __shared__ __align__(128) unsigned char pg_smem[];
  double* win = reinterpret_cast<double*>(pg_smem + kPgWinOff);
  __shared__ __align__(8) uint64_t full[kPgSlots], freeb[kPgSlots];
  __shared__ int s_off[kPgSlots];   // where the pass in a slot starts in the ring
  __shared__ int s_tile, s_done, s_gated;
  const int tid = threadIdx.x, wid = tid >> 5, lane32 = tid & 31;
  unsigned* progress = ctl + kBgCtlProgress;
  const uint32_t base_addr = smem_u32(pg_smem);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kPgSlots; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&freeb[s], kPassRows / 32);   // the working warps of the group that relaxes the pass
    }
    *reinterpret_cast<double*>(pg_smem + kPgZeroOff) = 0.0;
    *reinterpret_cast<double*>(pg_smem + kPgZeroOff + 8) = 0.0;
    mbar_fence_init();
  }
  int qbase = 0;          // passes this CTA has walked so far (slots and phases continue across tiles)
  int p_head = 0, p_oldest = 0;   // producer: next free byte of the ring, oldest pass (sequence number) not yet given back
  for (;;) {
    if (tid == 0) {
      s_tile = (int)atomicAdd(&ctl[0], 1u);
      s_done = 0;
      s_gated = 0;
    }
    __syncthreads();
    const int tk = s_tile;
    if (tk >= ntiles) break;
    const int t = __ldg(order + tk);
    const int2 QT = __ldg(ptile + t);   // {first pass, end pass}
    const int np = QT.y - QT.x;

    if (wid == kPgCompute / 32) {
      // ------------------------------- producer -------------------------------
      if (lane32 == 0) {
        auto pass_rec = [&](int i) { return i < np ? __ldg(ppass + QT.x + i) : make_int4(0, 0, 0, 0); };
        int4 r0 = pass_rec(0), r1 = pass_rec(1), r2;
        for (int i = 0; i < np; ++i) {
          r2 = pass_rec(i + 2);
          const int rows = r0.y & 0xffff, slots = ((r0.y >> 16) & 0xff) + ((r0.y >> 24) & 0xff);
          const int entries = slots * ((rows * T + 31) & ~31);
          const int ra = r0.x & ~1, rcnt = (r0.x + rows - ra + 1) & ~1;
          const int need = (entries * 12 + rcnt * 16 + 127) & ~127;
          const int q = qbase + i, slot = q & (kPgSlots - 1);
          for (;;) {   // room in the ring: a circular first-in first-out buffer of whole passes
            const int outstanding = q - p_oldest;
            if (outstanding == 0) { p_head = 0; break; }
            if (outstanding < kPgSlots) {
              const int tail = s_off[p_oldest & (kPgSlots - 1)];
              if (p_head > tail) {
                if (p_head + need <= kPgRingBytes) break;
                if (need <= tail) { p_head = 0; break; }
              } else if (p_head < tail && p_head + need <= tail) {
                break;
              }
            }
            mbar_wait(&freeb[p_oldest & (kPgSlots - 1)], (uint32_t)((p_oldest / kPgSlots) & 1));
            ++p_oldest;
          }
          s_off[slot] = p_head;
          unsigned char* at = pg_smem + p_head;
          mbar_expect_tx(&full[slot], (uint32_t)(entries * 12 + rcnt * 16));
          if (entries) {
            bulk_g2s(at, val + r0.z, (uint32_t)entries * 8u, &full[slot]);
            bulk_g2s(at + entries * 8, idx + r0.z, (uint32_t)entries * 4u, &full[slot]);
          }
          bulk_g2s(at + entries * 12, b + ra, (uint32_t)rcnt * 8u, &full[slot]);
          bulk_g2s(at + entries * 12 + rcnt * 8, diag + ra, (uint32_t)rcnt * 8u, &full[slot]);
          p_head += need;
          r0 = r1;
          r1 = r2;
        }
      }
    } else if (wid == kPgCompute / 32 + 1) {
      // ------------------------------- gate + publisher -------------------------------
      int last_pub = 0;
      auto publish = [&]() {   // finished passes of this tile -> visible to other CTAs (lane 0; decoupled from the compute groups)
        if (lane32 == 0) {
          const int v = ld_acquire_cta_shared(&s_done);
          if (v > last_pub) {
            __threadfence();
            st_relaxed_gpu_u32(progress + t, (unsigned)v);
            last_pub = v;
          }
        }
      };
      bool bailed = false;
      int2 pq_next = lane32 < np ? __ldg(preq + QT.x + lane32) : make_int2(0, 0);
      for (int i0 = 0; i0 < np && !bailed; i0 += 32) {
        const int2 pq = pq_next;
        {
          const int nx = i0 + 32 + lane32;   // the next batch's records travel while this one is polled
          pq_next = nx < np ? __ldg(preq + QT.x + nx) : make_int2(0, 0);
        }
        const int rb = pq.x, rc = (i0 + lane32 < np) ? pq.y : 0;
        int j = 0, released = i0;
        bool done = rc == 0;
        int2 rq = rc > 0 ? __ldg(req + rb) : make_int2(0, 0);
        long long t0 = 0;
        unsigned spins = 0;
        for (;;) {
          if (!done) {
            const unsigned seen = ld_acquire_u32(progress + rq.x);
            if (seen >= (unsigned)rq.y) {
              ++j;
              done = j >= rc;
              if (!done) rq = __ldg(req + rb + j);
            }
          }
          const unsigned m = __ballot_sync(0xffffffffu, done);
          __syncwarp();
          const int nready = m == 0xffffffffu ? 32 : __ffs((int)~m) - 1;
          const int upto = min(i0 + nready, np);
          if (upto > released) {
            if (lane32 == 0) st_release_cta_shared(&s_gated, upto);
            released = upto;
          }
          publish();
          if (m == 0xffffffffu) break;
          __nanosleep(20);
          bool bail = false;
          if ((++spins & 0xfffu) == 0u) {   // watchdog: a protocol error must not hang the device (the host reports it)
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else bail = now - t0 > 8000000000ll;
          }
          if (__any_sync(0xffffffffu, bail)) {
            if (lane32 == 0) {
              atomicExch(fault, 1);
              st_release_cta_shared(&s_gated, np);
            }
            bailed = true;
            break;
          }
        }
      }
      for (;;) {
        publish();
        if (__shfl_sync(0xffffffffu, last_pub, 0) >= np) break;
        __nanosleep(100);
      }
    } else if ((wid & 7) == 7) {
      // ------------------------------- signalling warp of a compute group -------------------------------
      int i = wid >> 3;
      for (; i < np; i += 2) {
        if (i > 0) pg_bar_sync(1 + ((i - 1) & 1));
        if (lane32 == 0 && i > 0) st_release_cta_shared(&s_done, i);   // passes < i are complete (nothing of this warp is in flight)
        pg_bar_arrive(1 + (i & 1));
      }
      if (i == np) pg_bar_sync(1 + ((np - 1) & 1));
      asm volatile("bar.sync 3, %0;" ::"n"(kPgCompute) : "memory");
      if (wid == 7 && lane32 == 0) st_release_cta_shared(&s_done, np);
    } else {
      // ------------------------------- compute -------------------------------
      const int grp = wid >> 3;
      const int tg = tid & (kPassGroup - 1);   // < kPassRows
      const int lane = tg % T, rloc = tg / T;
      int gated_seen = 0;
      auto fetch = [&](int i) { return i < np ? __ldg(ppass + QT.x + i) : make_int4(0, 0, 0, -1); };
      // FAR: the x values that do not come from the window, into registers
      auto far = [&](const int4 r, int i, double (&xf)[kPassFar + 1]) {
#pragma unroll
        for (int j = 0; j <= kPassFar; ++j) xf[j] = 0.0;
        if (r.w < 0) return;
        while (gated_seen <= i) {   // the passes of other tiles this pass reads from are published
          gated_seen = ld_acquire_cta_shared(&s_gated);
          if (gated_seen <= i) __nanosleep(40);
        }
        const int q = qbase + i, slot = q & (kPgSlots - 1);
        mbar_wait(&full[slot], (uint32_t)((q / kPgSlots) & 1));
        const int rows = r.y & 0xffff, jf = (r.y >> 16) & 0xff, jn = (r.y >> 24) & 0xff;
        if (rloc < rows) {
          const int width = (rows * T + 31) & ~31;
          const int* fidx = reinterpret_cast<const int*>(pg_smem + s_off[slot] + (jf + jn) * width * 8) + tg;
          int c[kPassFar];
#pragma unroll
          for (int j = 0; j < kPassFar; ++j) c[j] = j < jf ? fidx[j * width] : 0;
#pragma unroll
          for (int j = 0; j < kPassFar; ++j)
            if (j < jf) xf[j] = __ldcg(x + c[j]);
          if (sor) xf[kPassFar] = __ldcg(x + r.x + rloc);
        }
      };
      // SUM: far products summed, near entries into registers, the row's right-hand side / diagonal / reciprocal
      auto sum = [&](const int4 r, int i, const double (&xf)[kPassFar + 1], PgWork& wk) {
        const int rows = r.w < 0 ? 0 : (r.y & 0xffff);
        const bool active = rloc < rows;
        wk.warp_active = __any_sync(0xffffffffu, active);
        wk.store = active && lane == 0;
        wk.row = r.x + rloc;
        if (!wk.warp_active) return;
        const int jf = (r.y >> 16) & 0xff, jn = (r.y >> 24) & 0xff;
        const int width = (rows * T + 31) & ~31;
        const int entries = (jf + jn) * width;
        const unsigned char* at = pg_smem + s_off[(qbase + i) & (kPgSlots - 1)];
        const double* sval = reinterpret_cast<const double*>(at) + tg;
        const int* sidx = reinterpret_cast<const int*>(at + entries * 8) + tg;
        double fp[kPassFar];
#pragma unroll
        for (int j = 0; j < kPassFar; ++j) fp[j] = __dmul_rn(j < jf ? sval[j * width] : 0.0, xf[j]);
        static_assert(kPassFar == 6 && kPassNearSlots == 4, "the sums below are written for six far and four near slots");
        wk.farsum = __dadd_rn(__dadd_rn(__dadd_rn(fp[0], fp[1]), __dadd_rn(fp[2], fp[3])), __dadd_rn(fp[4], fp[5]));
        const int nb = jf * width;
#pragma unroll
        for (int j = 0; j < kPassNearSlots; ++j) {
          wk.v[j] = j < jn ? sval[nb + j * width] : 0.0;
          wk.a[j] = base_addr + (uint32_t)(j < jn ? sidx[nb + j * width] : kPgZeroOff);
        }
        const int ra = r.x & ~1, rcnt = (r.x + rows - ra + 1) & ~1;
        const double* sb = reinterpret_cast<const double*>(at + entries * 12);
        const int rr = active ? wk.row - ra : 0;
        wk.bval = sb[rr];
        const double d = sb[rcnt + rr];
        wk.d = active ? d : 0.0;
        wk.ry = 0.0;
        if (wk.store && d != 0.0) wk.ry = sor ? __ddiv_rn(omega, d) : bg_rcp_refined(d);
        wk.xold = xf[kPassFar];
      };
      // after the hand-off: four window loads, four products, a two-level tree, the far part on top, the lane reduction, the
      // update.  (Summation order: far entries first, then the near ones, pairwise — the reference adds in entry order,
      // src/smoother.jl:81-86; the difference is rounding only, <= 1e-15 relative.)
      auto relax = [&](const PgWork& wk) {
        if (!wk.warp_active) return;
        double pr[kPassNearSlots];
#pragma unroll
        for (int j = 0; j < kPassNearSlots; ++j) pr[j] = __dmul_rn(wk.v[j], lds_f64(wk.a[j]));
        double rsum = __dadd_rn(__dadd_rn(pr[0], pr[1]), __dadd_rn(pr[2], pr[3]));
        rsum = __dadd_rn(wk.farsum, rsum);
        if (T > 1) rsum = bg_lanes_sum<T>(rsum);
        if (wk.store) {
          double xnew;
          const double d = wk.d;
          if (d != 0.0) {
            const double r = __dsub_rn(wk.bval, rsum);
            xnew = sor ? __dadd_rn(__dmul_rn(1.0 - omega, wk.xold), __dmul_rn(wk.ry, r)) : bg_div_finish(r, d, wk.ry);
            __stcg(x + wk.row, xnew);
          } else {
            xnew = __ldcg(x + wk.row);   // rows without a usable diagonal are left unchanged (smoother.jl:84-87)
          }
          win[wk.row & (kPassWindow - 1)] = xnew;
        }
      };
      auto give_back = [&](int i) {   // this warp no longer reads the ring data of pass i
        __syncwarp();
        if (lane32 == 0) bg_mbar_arrive(&freeb[(qbase + i) & (kPgSlots - 1)]);
      };
      const bool stamp = dbg != nullptr && tid == 0;
      const long long c_begin = stamp ? clock64() : 0;
      long long c_wait = 0, c_relax = 0, c_sum = 0;

      int i = grp;
      int4 rC, rB, rN;
      double xf[kPassFar + 1];
      PgWork wk;
      {
        const int4 r0 = fetch(i);
        rC = fetch(i + 2);
        rB = fetch(i + 4);
        rN = fetch(i + 6);
        far(r0, i, xf);
        sum(r0, i, xf, wk);
        if (i < np) give_back(i);
        far(rC, i + 2, xf);
      }
      for (; i < np; i += 2) {
        const long long w0 = stamp ? clock64() : 0;
        if (i > 0) pg_bar_sync(1 + ((i - 1) & 1));   // pass i - 1 is relaxed: its x is in the window and in global memory
        const long long w1 = stamp ? clock64() : 0;
        relax(wk);
        pg_bar_arrive(1 + (i & 1));
        const long long w2 = stamp ? clock64() : 0;
        sum(rC, i + 2, xf, wk);
        if (rC.w >= 0) give_back(i + 2);
        far(rB, i + 4, xf);
        rC = rB;
        rB = rN;
        rN = fetch(i + 8);
        if (stamp) { c_wait += w1 - w0; c_relax += w2 - w1; c_sum += clock64() - w2; }
      }
      if (i == np) pg_bar_sync(1 + ((np - 1) & 1));   // the other group's last pass
      asm volatile("bar.sync 3, %0;" ::"n"(kPgCompute) : "memory");
      if (stamp) {   // diagnostics (tools/block_timeline.py): SM cycles of compute thread 0 in this tile
        unsigned long long* o = dbg + 8 * (size_t)t;
        o[0] = (unsigned long long)(clock64() - c_begin);
        o[1] = (unsigned long long)c_wait;
        o[2] = (unsigned long long)c_relax;
        o[3] = (unsigned long long)c_sum;
        o[4] = 0ull;
        o[5] = (unsigned long long)((np + 1) / 2);
        o[6] = (unsigned long long)np;
        o[7] = global_ns();
      }
    }
    __syncthreads();
    qbase += np;
  }
}

}  // namespace b200amg
