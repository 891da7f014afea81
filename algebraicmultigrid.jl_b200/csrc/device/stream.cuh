// CSR "stream" kernels: the bandwidth path of the solve phase on sm_100a.
//
// The matrix streams (values, column indices, row pointers) are perfectly contiguous, so they are
// moved HBM -> shared memory by the TMA unit as 1-D bulk copies (cp.async.bulk, SASS UBLKCP)
// signalled on an mbarrier, three stages deep, by persistent CTAs; the threads only wait on the
// barrier, gather x through L1/L2 and reduce rows out of shared memory.  Nothing about the matrix
// ever occupies a register while it is in flight, which is what lets two CTAs per SM keep
// > 100 KB of HBM requests outstanding (Little's law needs ~31 KB per SM at 7.7 TB/s).
//
// A matrix is cut at upload into TILES of whole rows (<= kTileNnz non-zeros, <= rows_per_tile rows);
// tile t is described by one int4 {first row, end row, first nnz, end nnz}.  T lanes cooperate on
// a row (T = 1 for stencil-like matrices: one thread per row, sequential ascending-column
// accumulation with separate multiply and add roundings == the accumulation order of the
// reference's CSC scatter `y[rowval[k]] += nzval[k]*x[j]`, so T = 1 results are bit-identical to
// the CPU oracle; T > 1 differs by summation order only).
//
//   MODE 0  y  = A x                       mul!(res, A, x)                multilevel.jl:188,219,223
//   MODE 1  y  = b - A x                   res .= b .- res (fused)        multilevel.jl:189,220
//   MODE 2  y += A x                       mul!(res,P,cx); x .+= res      multilevel.jl:233-234
//   MODE 3  y  = jacobi_fast(x)            smooth!(::FastJacobiSmoother)  smoother.jl:113-141
//   MODE 4  y  = jacobi_general(x)         smooth!(::JacobiSmoother)      smoother.jl:157-171
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200amg {

constexpr int kStreamThreads = 256;
constexpr int kTileNnz = 2048;
constexpr int kTileRowsMax = 1024;
constexpr int kStages = 3;
constexpr int kStreamBurst = 8;   // x gathers a lane issues back to back for a short row (template default)

// VT: how the matrix values are STORED (double, or float when every value of the operator is exactly representable in
// binary32 — decided at upload, DevCsr::val32 — so that the products, computed in fp64 either way, are bit-identical: 8
// instead of 12 bytes per entry cross HBM)
template <typename VT>
struct __align__(16) StreamStageT {
  VT val[kTileNnz + 8];
  int col[kTileNnz + 8];
  int rp[kTileRowsMax + 8];
};
using StreamStage = StreamStageT<double>;
static_assert(sizeof(StreamStage) % 16 == 0 && sizeof(StreamStageT<float>) % 16 == 0, "stage must keep 16-byte alignment");
constexpr int kStreamSmemBytes = kStages * (int)sizeof(StreamStage);

// ---- mbarrier / bulk-copy primitives (PTX) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar` (TMA unit; SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <typename VT>
__device__ __forceinline__ void stream_issue(StreamStageT<VT>& S, uint64_t* bar, const int4 m, const int* __restrict__ rowptr,
                                             const int* __restrict__ col, const VT* __restrict__ val) {
  const int ka = m.z & ~3, kcnt = (m.w - ka + 3) & ~3;
  const int ra = m.x & ~3, rcnt = (m.y + 1 - ra + 3) & ~3;
  mbar_expect_tx(bar, (uint32_t)(kcnt * (4 + (int)sizeof(VT)) + rcnt * 4));
  if (kcnt) {
    bulk_g2s(S.val, val + ka, (uint32_t)kcnt * (uint32_t)sizeof(VT), bar);
    bulk_g2s(S.col, col + ka, (uint32_t)kcnt * 4u, bar);
  }
  bulk_g2s(S.rp, rowptr + ra, (uint32_t)rcnt * 4u, bar);
}

template <int T>
__device__ __forceinline__ double lanes_sum(double v) {
#pragma unroll
  for (int o = T / 2; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, T);
  return v;
}

// tiles are handed out in runs of `chunk` consecutive tiles per CTA (chunk = 1: round robin)
__device__ __forceinline__ int stream_tile_of(int i, int chunk) {
  return (i / chunk) * ((int)gridDim.x * chunk) + (int)blockIdx.x * chunk + (i % chunk);
}

template <int T, int MODE, int BURST = kStreamBurst, typename VT = double>
__global__ void __launch_bounds__(kStreamThreads, 2)
    csr_stream_kernel(int ntiles, int chunk, const int4* __restrict__ meta, const int* __restrict__ rowptr,
                      const int* __restrict__ col, const VT* __restrict__ val, const double* __restrict__ x,
                      const double* __restrict__ b, double* __restrict__ y, double omega,
                      const double* __restrict__ diagvals) {
  extern __shared__ __align__(128) unsigned char stream_smem[];
  StreamStageT<VT>* st = reinterpret_cast<StreamStageT<VT>*>(stream_smem);
  __shared__ __align__(8) uint64_t full[kStages];
  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      const int t = stream_tile_of(s, chunk);
      if (t < ntiles) stream_issue(st[s], &full[s], __ldg(meta + t), rowptr, col, val);
    }
  }
  constexpr int G = kStreamThreads / T;   // rows relaxed per pass
  const int g = tid / T, lane = tid % T;
  int s = 0;
  uint32_t parity = 0;
  // the CTA's tiles are a prefix-closed sequence: stop at the first index beyond the matrix
  for (int i = 0;; ++i) {
    const int t = stream_tile_of(i, chunk);
    if (t >= ntiles) {
      // with chunk > 1 a later run can still hold valid tiles only if this one did: runs are
      // ordered, so the first invalid tile ends the CTA's work.
      break;
    }
    const int4 m = __ldg(meta + t);
    const int nrows = m.y - m.x, ka = m.z & ~3, rofs = m.x - (m.x & ~3);
    // operands of the epilogue of the first pass: requested before the barrier wait
    double pre_b[2] = {0.0, 0.0}, pre_x[2] = {0.0, 0.0}, pre_d[2] = {0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (lane == 0 && g + q * G < nrows) {
        const int row = m.x + g + q * G;
        if (MODE == 1 || MODE == 3 || MODE == 4) pre_b[q] = __ldg(b + row);
        if (MODE == 2) pre_x[q] = y[row];
        if (MODE == 3 || MODE == 4) pre_x[q] = __ldg(x + row);
        if (MODE == 4) pre_d[q] = __ldg(diagvals + row);
      }
    mbar_wait(&full[s], parity);
    const StreamStageT<VT>& S = st[s];
    for (int rbase = 0; rbase < nrows; rbase += G) {
      const int r = rbase + g;
      double sum = 0.0, diag = 0.0;
      if (r < nrows) {
        const int ks = S.rp[rofs + r] - ka, ke = S.rp[rofs + r + 1] - ka;
        const int row = m.x + r;
        if (ke - ks <= BURST * T) {
          // short row (every stencil row): ALL its x gathers leave in one burst, so the row costs one L2
          // round trip; the products are then added in the reference's order (idle slots add an exact +0,
          // as the reference itself does for the diagonal in smooth!: `rsum += ifelse(row == i, z, ...)`)
          int c[BURST];
          double v[BURST], xv[BURST];
#pragma unroll
          for (int j = 0; j < BURST; ++j) {
            const int k = ks + lane + j * T;
            const bool in = k < ke;
            c[j] = in ? S.col[k] : -1;
            v[j] = in ? (double)S.val[k] : 0.0;
          }
#pragma unroll
          for (int j = 0; j < BURST; ++j) {
            const bool use = c[j] >= 0 && !(MODE == 3 && c[j] == row);
            xv[j] = use ? __ldg(x + c[j]) : 0.0;
            if (MODE == 3 && c[j] == row) diag = v[j];
          }
#pragma unroll
          for (int j = 0; j < BURST; ++j) sum = __dadd_rn(sum, __dmul_rn(v[j], xv[j]));
        } else {
#pragma unroll 4
          for (int k = ks + lane; k < ke; k += T) {
            const int c = S.col[k];
            const double v = (double)S.val[k];
            if (MODE == 3) {
              if (c == row) diag = v;
              else sum = __dadd_rn(sum, __dmul_rn(v, __ldg(x + c)));
            } else {
              sum = __dadd_rn(sum, __dmul_rn(v, __ldg(x + c)));
            }
          }
        }
      }
      if (T > 1) {
        sum = lanes_sum<T>(sum);
        if (MODE == 3) diag = lanes_sum<T>(diag);
      }
      if (lane == 0 && r < nrows) {
        const int row = m.x + r;
        double bv, xv, dv;
        if (rbase == 0) { bv = pre_b[0]; xv = pre_x[0]; dv = pre_d[0]; }
        else if (rbase == G) { bv = pre_b[1]; xv = pre_x[1]; dv = pre_d[1]; }
        else {
          bv = xv = dv = 0.0;
          if (MODE == 1 || MODE == 3 || MODE == 4) bv = __ldg(b + row);
          if (MODE == 2) xv = y[row];
          if (MODE == 3 || MODE == 4) xv = __ldg(x + row);
          if (MODE == 4) dv = __ldg(diagvals + row);
        }
        if (MODE == 0) y[row] = sum;
        else if (MODE == 1) y[row] = bv - sum;
        else if (MODE == 2) y[row] = xv + sum;
        else if (MODE == 3)   // (one - ω) * temp[i] + ω * ((b[i] - rsum) / diag), no FMA contraction
          y[row] = (diag == 0.0) ? xv
                                 : __dadd_rn(__dmul_rn(1.0 - omega, xv), __dmul_rn(omega, __ddiv_rn(__dsub_rn(bv, sum), diag)));
        else                  // x[i] -= ω * temp[i] / d   with temp = A x - b
          y[row] = (dv != 0.0) ? __dsub_rn(xv, __ddiv_rn(__dmul_rn(omega, __dsub_rn(sum, bv)), dv)) : xv;
      }
    }
    __syncthreads();   // every thread is done with stage s
    if (tid == 0) {
      const int tn = stream_tile_of(i + kStages, chunk);
      if (tn < ntiles) stream_issue(st[s], &full[s], __ldg(meta + tn), rowptr, col, val);
    }
    if (++s == kStages) { s = 0; parity ^= 1u; }
  }
}



// =============================================================================================
// Gauss-Seidel / SOR as a DATAFLOW sweep (exact lexicographic semantics of gs! smoother.jl:73-90 and
// sor_step! :205-221).
//
// Levels that are relaxed by Gauss-Seidel / SOR are RENUMBERED at upload into wavefront order (level
// schedule of the symmetrised pattern for the ascending-index sweep: rows of one wavefront are mutually
// independent, every earlier-ordered neighbour sits in an earlier wavefront, every later-ordered one
// in a later wavefront; the backward sweep walks the same wavefronts in reverse).  Inside a row the
// entries keep the reference's order.  A wavefront is then a contiguous row range of the ordinary CSR
// arrays, neighbouring rows of a wavefront read neighbouring x entries (coalesced), and the sweep is
// ONE persistent kernel:
//   * CTAs claim TASKS (<= BS/T consecutive rows of one wavefront) in sweep order from a ticket counter;
//   * everything that does not depend on this sweep's progress is requested BEFORE waiting: row
//     pointers, column indices, values, b, and the x entries of LATER-ordered neighbours (nobody
//     touches those until this task has published);
//   * one thread polls the previous wavefront's completion counter; then the EARLIER-ordered x
//     entries are gathered in a single burst straight from L2 (ld.cg: other SMs wrote them), the row is
//     relaxed, and the task publishes with fence + counter increment.
// Dependency latency per wavefront is one L2 hand-off instead of a kernel launch or a grid barrier,
// and the matrix streams of later wavefronts are already in registers while earlier ones resolve.
// Claiming by ticket makes it deadlock-free for any grid size: a task is only ever held by a resident
// CTA, and the lowest unfinished task never waits on anything unfinished.
//
// T = 1: one thread per row, sequential accumulation in the reference's order, separate multiply/add
// roundings and a true division -> bit-identical to the reference's sequential sweep.
// =============================================================================================
constexpr int kGsPrefetch = 8;   // (column, value) pairs a lane holds in registers across the wait
constexpr int kGsCounterStride = 64;   // uints between counters: 256 B apart = different L2 slices, no hot line
constexpr int kGsFarWaves = 4;   // CTAs this many wavefronts ahead of the front sleep-poll

__device__ double g_gs_zero = 0.0;   // what idle slots of the gather burst read

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_relaxed_inc(unsigned* p) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// eight L1-bypassing loads issued back to back (one asm block: NVVM cannot sink them)
__device__ __forceinline__ void ldcg_burst8(double (&out)[kGsPrefetch], const double* const (&a)[kGsPrefetch]) {
  static_assert(kGsPrefetch == 8, "the burst is written for 8 slots");
  asm volatile(
      "ld.global.cg.f64 %0, [%8];\n\t"
      "ld.global.cg.f64 %1, [%9];\n\t"
      "ld.global.cg.f64 %2, [%10];\n\t"
      "ld.global.cg.f64 %3, [%11];\n\t"
      "ld.global.cg.f64 %4, [%12];\n\t"
      "ld.global.cg.f64 %5, [%13];\n\t"
      "ld.global.cg.f64 %6, [%14];\n\t"
      "ld.global.cg.f64 %7, [%15];"
      : "=d"(out[0]), "=d"(out[1]), "=d"(out[2]), "=d"(out[3]), "=d"(out[4]), "=d"(out[5]), "=d"(out[6]), "=d"(out[7])
      : "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]), "l"(a[4]), "l"(a[5]), "l"(a[6]), "l"(a[7])
      : "memory");
}
// ptxas sinks each load of a burst down to its first use, and in-order issue then serialises the L2
// round trips (measured: 1.0 us of a 2.0 us hop).  Folding all results into one run-time zero that
// every later use depends on pins the whole burst in front of the first wait.
__device__ __forceinline__ void pin_burst8(double (&v)[kGsPrefetch], int opaque_zero) {
  int dep = __double2hiint(v[0]);
#pragma unroll
  for (int j = 1; j < kGsPrefetch; ++j) dep &= __double2hiint(v[j]);
  dep &= opaque_zero;
#pragma unroll
  for (int j = 0; j < kGsPrefetch; ++j) v[j] = __hiloint2double(__double2hiint(v[j]) | dep, __double2loint(v[j]));
}

__device__ __forceinline__ void st_mail(uint4* p, double v, unsigned e) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)__double2loint(v)), "r"(e),
               "r"((unsigned)__double2hiint(v)), "r"(e)
               : "memory");
}
// eight mailbox polls issued back to back
__device__ __forceinline__ void ld_mail8(uint4 (&m)[8], const uint4* const (&a)[8]) {
  asm volatile(
      "ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%32];\n\t"
      "ld.relaxed.gpu.global.v4.u32 {%4, %5, %6, %7}, [%33];\n\t"
      "ld.relaxed.gpu.global.v4.u32 {%8, %9, %10, %11}, [%34];\n\t"
      "ld.relaxed.gpu.global.v4.u32 {%12, %13, %14, %15}, [%35];\n\t"
      "ld.relaxed.gpu.global.v4.u32 {%16, %17, %18, %19}, [%36];\n\t"
      "ld.relaxed.gpu.global.v4.u32 {%20, %21, %22, %23}, [%37];\n\t"
      "ld.relaxed.gpu.global.v4.u32 {%24, %25, %26, %27}, [%38];\n\t"
      "ld.relaxed.gpu.global.v4.u32 {%28, %29, %30, %31}, [%39];"
      : "=r"(m[0].x), "=r"(m[0].y), "=r"(m[0].z), "=r"(m[0].w), "=r"(m[1].x), "=r"(m[1].y), "=r"(m[1].z), "=r"(m[1].w),
        "=r"(m[2].x), "=r"(m[2].y), "=r"(m[2].z), "=r"(m[2].w), "=r"(m[3].x), "=r"(m[3].y), "=r"(m[3].z), "=r"(m[3].w),
        "=r"(m[4].x), "=r"(m[4].y), "=r"(m[4].z), "=r"(m[4].w), "=r"(m[5].x), "=r"(m[5].y), "=r"(m[5].z), "=r"(m[5].w),
        "=r"(m[6].x), "=r"(m[6].y), "=r"(m[6].z), "=r"(m[6].w), "=r"(m[7].x), "=r"(m[7].y), "=r"(m[7].z), "=r"(m[7].w)
      : "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]), "l"(a[4]), "l"(a[5]), "l"(a[6]), "l"(a[7])
      : "memory");
}
// the same burst, but only the slots named in `mask` are loaded (the others keep their contents and cost no L2 request):
// a row polls ONLY the mailboxes it still waits for
__device__ __forceinline__ void ld_mail8_masked(uint4 (&m)[8], const uint4* const (&a)[8], unsigned mask) {
  asm volatile(
      "{\n\t"
      ".reg .pred q0, q1, q2, q3, q4, q5, q6, q7;\n\t"
      ".reg .b32 t;\n\t"
      "and.b32 t, %40, 1;\n\t   setp.ne.b32 q0, t, 0;\n\t"
      "and.b32 t, %40, 2;\n\t   setp.ne.b32 q1, t, 0;\n\t"
      "and.b32 t, %40, 4;\n\t   setp.ne.b32 q2, t, 0;\n\t"
      "and.b32 t, %40, 8;\n\t   setp.ne.b32 q3, t, 0;\n\t"
      "and.b32 t, %40, 16;\n\t  setp.ne.b32 q4, t, 0;\n\t"
      "and.b32 t, %40, 32;\n\t  setp.ne.b32 q5, t, 0;\n\t"
      "and.b32 t, %40, 64;\n\t  setp.ne.b32 q6, t, 0;\n\t"
      "and.b32 t, %40, 128;\n\t setp.ne.b32 q7, t, 0;\n\t"
      "@q0 ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%32];\n\t"
      "@q1 ld.relaxed.gpu.global.v4.u32 {%4, %5, %6, %7}, [%33];\n\t"
      "@q2 ld.relaxed.gpu.global.v4.u32 {%8, %9, %10, %11}, [%34];\n\t"
      "@q3 ld.relaxed.gpu.global.v4.u32 {%12, %13, %14, %15}, [%35];\n\t"
      "@q4 ld.relaxed.gpu.global.v4.u32 {%16, %17, %18, %19}, [%36];\n\t"
      "@q5 ld.relaxed.gpu.global.v4.u32 {%20, %21, %22, %23}, [%37];\n\t"
      "@q6 ld.relaxed.gpu.global.v4.u32 {%24, %25, %26, %27}, [%38];\n\t"
      "@q7 ld.relaxed.gpu.global.v4.u32 {%28, %29, %30, %31}, [%39];\n\t"
      "}"
      : "+r"(m[0].x), "+r"(m[0].y), "+r"(m[0].z), "+r"(m[0].w), "+r"(m[1].x), "+r"(m[1].y), "+r"(m[1].z), "+r"(m[1].w),
        "+r"(m[2].x), "+r"(m[2].y), "+r"(m[2].z), "+r"(m[2].w), "+r"(m[3].x), "+r"(m[3].y), "+r"(m[3].z), "+r"(m[3].w),
        "+r"(m[4].x), "+r"(m[4].y), "+r"(m[4].z), "+r"(m[4].w), "+r"(m[5].x), "+r"(m[5].y), "+r"(m[5].z), "+r"(m[5].w),
        "+r"(m[6].x), "+r"(m[6].y), "+r"(m[6].z), "+r"(m[6].w), "+r"(m[7].x), "+r"(m[7].y), "+r"(m[7].z), "+r"(m[7].w)
      : "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]), "l"(a[4]), "l"(a[5]), "l"(a[6]), "l"(a[7]), "r"(mask)
      : "memory");
}
// tasks[t] = {first row, rows, wavefront in sweep order, tasks of the previous wavefront}
// counters[0] = ticket, counters[(1 + w) * kGsCounterStride] = finished tasks of wavefront w (zeroed before the launch)
// MAIL: rows are additionally published as 16-byte {value, epoch} mailboxes (see gs_mail_kernel) and the
// earlier-ordered neighbours are read from there.  The flags vouch for the data, so the producer needs NO
// fence between its stores and the completion count (0.4 us off every hop); a mailbox that is not visible
// yet when the count is, is simply polled again.
template <int T, int BS, bool MAIL>
__global__ void __launch_bounds__(BS)
    gs_dataflow_kernel(int ntasks, const int4* __restrict__ tasks, unsigned* counters, const int* __restrict__ rowptr,
                       const int* __restrict__ col, const double* __restrict__ val, double* x, const double* __restrict__ b,
                       double omega, int sor, int backward, int acquire_mode, int opaque_zero, unsigned long long* dbg,
                       uint4* mail, const unsigned* mail_ctl) {
  __shared__ int s_task[2];
  const int tid = threadIdx.x, g = tid / T, lane = tid % T;
  const unsigned e = MAIL ? ld_relaxed_u32(mail_ctl + 1) : 0u;
  if (tid == 0) s_task[0] = (int)atomicAdd(&counters[0], 1u);
  __syncthreads();
  int cur = s_task[0], buf = 0;
  while (cur < ntasks) {
    if (tid == 0) s_task[buf ^ 1] = (int)atomicAdd(&counters[0], 1u);   // next ticket, off the critical path
    const int4 tk = __ldg(tasks + cur);
    if (dbg && tid == 0) dbg[(size_t)cur * 8 + 0] = global_ns();
    const bool active = g < tk.y;
    const int row = active ? tk.x + g : -1;
    int ks = 0, ke = 0;
    double bv = 0.0, xold = 0.0;
    if (active) {
      ks = __ldg(rowptr + row);
      ke = __ldg(rowptr + row + 1);
      if (lane == 0) {
        bv = __ldg(b + row);
        if (sor || MAIL) xold = __ldcg(x + row);   // only this row's own update ever writes x[row] during the sweep
      }
    }
    int c[kGsPrefetch];
    double v[kGsPrefetch];
#pragma unroll
    for (int j = 0; j < kGsPrefetch; ++j) {
      const int k = ks + lane + j * T;
      const bool in = k < ke;
      c[j] = in ? __ldg(col + k) : -1;
      v[j] = in ? __ldg(val + k) : 0.0;
    }
    // later-ordered neighbours keep their old value until this task has published: fetch them now
    double xl[kGsPrefetch];
    {
      const double* a[kGsPrefetch];
#pragma unroll
      for (int j = 0; j < kGsPrefetch; ++j) {
        const bool later = c[j] >= 0 && (backward ? c[j] < row : c[j] > row);
        a[j] = later ? x + c[j] : &g_gs_zero;
      }
      ldcg_burst8(xl, a);
    }
    if (tk.z > 0) {
      if (tid == 0) {
        const unsigned need = (unsigned)tk.w;   // tasks of wavefront tk.z - 1
        const unsigned* flag = counters + (size_t)tk.z * kGsCounterStride;
        if (tk.z >= kGsFarWaves) {   // far from the front: sleep instead of hammering L2
          const unsigned* far = flag - (size_t)(kGsFarWaves - 1) * kGsCounterStride;   // wavefront tk.z - kGsFarWaves
          while (ld_relaxed_u32(far) == 0u) __nanosleep(500);
        }
        while (ld_relaxed_u32(flag) < need) {}
        if (dbg) dbg[(size_t)cur * 8 + 2] = global_ns();
        // Acquire side.  The producers performed every x store at L2 (fence) before their count became
        // visible there, and the gathers below are L1-bypassing loads issued after this bar.sync, so
        // they read L2 no earlier than the observed count: no consumer fence is needed on this
        // hardware.  acquire_mode 1/2 add the formal PTX acquire (re-read with ld.acquire / full fence).
        if (acquire_mode == 1) (void)ld_acquire_u32(flag);
        else if (acquire_mode == 2) __threadfence();
      }
      __syncthreads();
    }
    if (dbg && tid == 0) dbg[(size_t)cur * 8 + 3] = global_ns();
    // earlier-ordered neighbours: one burst, one L2 round trip on the critical path
    double xe[kGsPrefetch];
    {
      const double* a[kGsPrefetch];
#pragma unroll
      for (int j = 0; j < kGsPrefetch; ++j) {
        const bool earlier = c[j] >= 0 && (backward ? c[j] > row : c[j] < row);
        a[j] = earlier ? x + c[j] : &g_gs_zero;
      }
      if (!MAIL) {
        ldcg_burst8(xe, a);
        pin_burst8(xe, opaque_zero);
      } else {
        unsigned need = 0u;
#pragma unroll
        for (int j = 0; j < kGsPrefetch; ++j) {
          xe[j] = 0.0;
          if (a[j] != &g_gs_zero) need |= 1u << j;
        }
        while (need) {
          const uint4* ma[kGsPrefetch];
          uint4 mm[kGsPrefetch];
#pragma unroll
          for (int j = 0; j < kGsPrefetch; ++j) ma[j] = ((need >> j) & 1u) ? mail + c[j] : mail + (row >= 0 ? row : 0);
          ld_mail8(mm, ma);
          {
            unsigned dep = mm[0].y;
#pragma unroll
            for (int j = 1; j < kGsPrefetch; ++j) dep &= mm[j].y;
            dep &= (unsigned)opaque_zero;
#pragma unroll
            for (int j = 0; j < kGsPrefetch; ++j) mm[j].y |= dep;
          }
#pragma unroll
          for (int j = 0; j < kGsPrefetch; ++j)
            if (((need >> j) & 1u) && mm[j].y == e && mm[j].w == e) {
              xe[j] = __hiloint2double((int)mm[j].z, (int)mm[j].x);
              need &= ~(1u << j);
            }
        }
        __syncwarp();
      }
    }
    if (dbg && tid == 0) dbg[(size_t)cur * 8 + 1] = global_ns() + (unsigned long long)(__double2loint(xe[0]) & opaque_zero);
    double rsum = 0.0, d = 0.0;
#pragma unroll
    for (int j = 0; j < kGsPrefetch; ++j) {
      if (c[j] == row && c[j] >= 0) d = v[j];
      else {
        const bool later = c[j] >= 0 && (backward ? c[j] < row : c[j] > row);
        rsum = __dadd_rn(rsum, __dmul_rn(v[j], later ? xl[j] : xe[j]));   // idle slots add an exact +0
      }
    }
    for (int k = ks + lane + kGsPrefetch * T; k < ke; k += T) {   // rows longer than T * kGsPrefetch
      const int cc = __ldg(col + k);
      const double vv = __ldg(val + k);
      if (cc == row) { d = vv; continue; }
      double xv;
      if (MAIL && (backward ? cc > row : cc < row)) {
        uint4 q;
        do {
          asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                       : "l"(mail + cc)
                       : "memory");
        } while (q.y != e || q.w != e);
        xv = __hiloint2double((int)q.z, (int)q.x);
      } else {
        xv = __ldcg(x + cc);
      }
      rsum = __dadd_rn(rsum, __dmul_rn(vv, xv));
    }
    if (T > 1) {
      rsum = lanes_sum<T>(rsum);
      d = lanes_sum<T>(d);
    }
    if (MAIL) {
      if (active && lane == 0) {
        double xnew = xold;
        if (d != 0.0) {
          const double r = __dsub_rn(bv, rsum);
          xnew = sor ? __dadd_rn(__dmul_rn(1.0 - omega, xold), __dmul_rn(__ddiv_rn(omega, d), r)) : __ddiv_rn(r, d);
        }
        st_mail(mail + row, xnew, e);
        __stcg(x + row, xnew);
      }
    } else if (active && lane == 0 && d != 0.0) {
      const double r = __dsub_rn(bv, rsum);
      __stcg(x + row, sor ? __dadd_rn(__dmul_rn(1.0 - omega, xold), __dmul_rn(__ddiv_rn(omega, d), r)) : __ddiv_rn(r, d));
    }
    if (dbg && tid == 0) dbg[(size_t)cur * 8 + 4] = global_ns();
    __syncthreads();
    if (tid == 0) {
      if (dbg) dbg[(size_t)cur * 8 + 5] = global_ns();
      if (!MAIL) __threadfence();   // release side: every row of this task (bar.sync above) is visible before the count
      if (dbg) dbg[(size_t)cur * 8 + 6] = global_ns();
      red_relaxed_inc(counters + (size_t)(1 + tk.z) * kGsCounterStride);
      if (dbg) dbg[(size_t)cur * 8 + 7] = global_ns();
    }
    cur = s_task[buf ^ 1];
    buf ^= 1;
  }
}



// =============================================================================================
// Gauss-Seidel / SOR, per-ROW dataflow ("mailbox" sweep) — the default for structurally symmetric
// matrices.  Same renumbered layout, tasks and ticket scheduling as gs_dataflow_kernel, but the
// hand-off between dependent rows carries the DATA WITH THE FLAG, the way NCCL's LL protocol does:
// every relaxed row publishes its new value as one 16-byte store {lo32, epoch, hi32, epoch} into
// mail[row]; a consumer polls the mailboxes of its earlier-ordered neighbours with 16-byte loads
// until both epoch words match (8-byte halves are single-copy atomic, so a matching pair of halves
// IS the value).  No fence, no completion counter, no CTA barrier on the critical path: one hop is
// store -> L2 -> poll, roughly a third of the counter protocol's fence + count + poll + gather.
// Rows advance as soon as THEIR neighbours are done; there is no wavefront-wide barrier at all.
//
// Throttle: a task of wavefront w first sleeps until some task of wavefront w - 2 has finished
// (started[] hint), so only ~2 wavefronts' worth of threads ever spin on mailboxes.
//
// Anti-dependencies (reading the OLD value of a later-ordered neighbour j) are safe because the
// pattern is symmetric: j's row contains this row, so j cannot be relaxed before this row publishes,
// which happens after the old value was read.
//
// ctl[0] = ticket, ctl[1] = epoch (gs_mail_prepare_kernel bumps it before every sweep),
// ctl[(2 + w) * kGsCounterStride] = epoch of the last sweep in which a task of wavefront w finished.
// =============================================================================================
__global__ void gs_mail_prepare_kernel(unsigned* ctl) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    ctl[0] = 0u;
    ctl[1] = ctl[1] + 1u;
  }
}
// four mailbox polls issued back to back
__device__ __forceinline__ void ld_mail4(uint4 (&m)[4], const uint4* const (&a)[4]) {
  asm volatile(
      "ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%16];\n\t"
      "ld.relaxed.gpu.global.v4.u32 {%4, %5, %6, %7}, [%17];\n\t"
      "ld.relaxed.gpu.global.v4.u32 {%8, %9, %10, %11}, [%18];\n\t"
      "ld.relaxed.gpu.global.v4.u32 {%12, %13, %14, %15}, [%19];"
      : "=r"(m[0].x), "=r"(m[0].y), "=r"(m[0].z), "=r"(m[0].w), "=r"(m[1].x), "=r"(m[1].y), "=r"(m[1].z), "=r"(m[1].w),
        "=r"(m[2].x), "=r"(m[2].y), "=r"(m[2].z), "=r"(m[2].w), "=r"(m[3].x), "=r"(m[3].y), "=r"(m[3].z), "=r"(m[3].w)
      : "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3])
      : "memory");
}

template <int T, int BS>
__global__ void __launch_bounds__(BS)
    gs_mail_kernel(int ntasks, const int4* __restrict__ tasks, unsigned* ctl, const int* __restrict__ rowptr,
                   const int* __restrict__ col, const double* __restrict__ val, double* x, const double* __restrict__ b,
                   uint4* mail, double omega, int sor, int backward, int opaque_zero, int poll_sleep, int gate_sleep) {
  __shared__ int s_task[2];
  const int tid = threadIdx.x, g = tid / T, lane = tid % T;
  const unsigned e = ld_relaxed_u32(ctl + 1);
  if (tid == 0) s_task[0] = (int)atomicAdd(&ctl[0], 1u);
  __syncthreads();
  int cur = s_task[0], buf = 0;
  while (cur < ntasks) {
    if (tid == 0) s_task[buf ^ 1] = (int)atomicAdd(&ctl[0], 1u);   // next ticket, off the critical path
    const int4 tk = __ldg(tasks + cur);
    const bool active = g < tk.y;
    const int row = active ? tk.x + g : -1;
    int ks = 0, ke = 0;
    double bv = 0.0, xold = 0.0;
    if (active) {
      ks = __ldg(rowptr + row);
      ke = __ldg(rowptr + row + 1);
      if (lane == 0) {
        bv = __ldg(b + row);
        xold = __ldcg(x + row);   // only this row's own update ever writes x[row] during the sweep
      }
    }
    int c[kGsPrefetch];
    double v[kGsPrefetch];
#pragma unroll
    for (int j = 0; j < kGsPrefetch; ++j) {
      const int k = ks + lane + j * T;
      const bool in = k < ke;
      c[j] = in ? __ldg(col + k) : -1;
      v[j] = in ? __ldg(val + k) : 0.0;
    }
    // later-ordered neighbours keep their old value until this row has published: fetch them now
    double xn[kGsPrefetch];   // neighbour values: later ones now, earlier ones from the mailboxes below
    unsigned need = 0u;       // slots still waiting for an earlier-ordered neighbour
    {
      const double* a[kGsPrefetch];
#pragma unroll
      for (int j = 0; j < kGsPrefetch; ++j) {
        const bool valid = c[j] >= 0 && c[j] != row;
        const bool earlier = valid && (backward ? c[j] > row : c[j] < row);
        if (earlier) need |= 1u << j;
        a[j] = (valid && !earlier) ? x + c[j] : &g_gs_zero;
      }
      ldcg_burst8(xn, a);
    }
    if (tk.z >= 2) {   // throttle: stay asleep until wavefront tk.z - 2 has begun to finish
      if (tid == 0) {
        const unsigned* hint = ctl + (size_t)tk.z * kGsCounterStride;   // (2 + (tk.z - 2))
        while (ld_relaxed_u32(hint) != e) if (gate_sleep) __nanosleep(gate_sleep);
      }
      __syncthreads();
    }
    // ---- poll the mailboxes of the earlier-ordered neighbours, four slots at a time ----
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      unsigned want = (need >> (4 * half)) & 0xfu;
      while (want) {
        const uint4* a[4];
        uint4 m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = ((want >> j) & 1u) ? mail + c[4 * half + j] : mail + (row >= 0 ? row : 0);
        ld_mail4(m, a);
        {   // pin the four polls in front of the first flag test (see pin_burst8)
          const unsigned dep = (m[0].y & m[1].y & m[2].y & m[3].y) & (unsigned)opaque_zero;
#pragma unroll
          for (int j = 0; j < 4; ++j) m[j].y |= dep;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (((want >> j) & 1u) && m[j].y == e && m[j].w == e) {
            xn[4 * half + j] = __hiloint2double((int)m[j].z, (int)m[j].x);
            want &= ~(1u << j);
          }
        if (want && poll_sleep) __nanosleep(poll_sleep);
      }
      __syncwarp();   // lanes leave the poll loop at different times: reconverge before going on
    }
    double rsum = 0.0, d = 0.0;
#pragma unroll
    for (int j = 0; j < kGsPrefetch; ++j) {
      if (c[j] == row && c[j] >= 0) d = v[j];
      else rsum = __dadd_rn(rsum, __dmul_rn(v[j], xn[j]));   // idle slots add an exact +0
    }
    for (int k = ks + lane + kGsPrefetch * T; k < ke; k += T) {   // rows longer than T * kGsPrefetch
      const int cc = __ldg(col + k);
      const double vv = __ldg(val + k);
      if (cc == row) { d = vv; continue; }
      double xv;
      if (backward ? cc > row : cc < row) {
        uint4 m;
        do {
          asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(m.x), "=r"(m.y), "=r"(m.z), "=r"(m.w)
                       : "l"(mail + cc)
                       : "memory");
        } while (m.y != e || m.w != e);
        xv = __hiloint2double((int)m.z, (int)m.x);
      } else {
        xv = __ldcg(x + cc);
      }
      rsum = __dadd_rn(rsum, __dmul_rn(vv, xv));
    }
    if (T > 1) {
      rsum = lanes_sum<T>(rsum);
      d = lanes_sum<T>(d);
    }
    if (active && lane == 0) {
      double xnew = xold;
      if (d != 0.0) {
        const double r = __dsub_rn(bv, rsum);
        xnew = sor ? __dadd_rn(__dmul_rn(1.0 - omega, xold), __dmul_rn(__ddiv_rn(omega, d), r)) : __ddiv_rn(r, d);
      }
      st_mail(mail + row, xnew, e);   // publish first: this is what dependants spin on
      __stcg(x + row, xnew);
    }
    __syncthreads();
    if (tid == 0) {
      volatile unsigned* mine = ctl + (size_t)(2 + tk.z) * kGsCounterStride;
      *mine = e;   // throttle hint only: no ordering required
    }
    cur = s_task[buf ^ 1];
    buf ^= 1;
  }
}


// =============================================================================================
// Gauss-Seidel / SOR on a SMALL level: one CTA (1024 threads) sweeps the whole level.
// Coarse levels have thousands of tiny wavefronts (20-200 rows each); across SMs every wavefront costs
// a 0.5-2 us L2 hand-off, inside one SM it costs a bar.sync.  The matrix (renumbered, so wavefronts are
// contiguous row ranges) is streamed through the same TMA tile pipeline as the SpMV kernels — row
// pointers, column indices, values and b are in shared memory before they are needed — and x lives in
// shared memory as well when the level fits (XS), else it is read/written through L2 (ld.cg / st.cg,
// ordered by the barrier).  T lanes cooperate on a row.
// lvlptr: forward wavefront boundaries (row indices); a backward sweep walks tiles and wavefronts in reverse.
// =============================================================================================
constexpr int kGsCtaThreads = 1024;
struct __align__(16) GsCtaStage {
  double val[kTileNnz + 8];
  double b[kTileRowsMax + 8];
  int col[kTileNnz + 8];
  int rp[kTileRowsMax + 8];
};
static_assert(sizeof(GsCtaStage) % 16 == 0, "stage must keep 16-byte alignment");

__device__ __forceinline__ void gs_cta_issue(GsCtaStage& S, uint64_t* bar, const int4 m, const int* __restrict__ rowptr,
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

template <int T, bool XS>
__global__ void __launch_bounds__(kGsCtaThreads, 1)
    gs_cta_kernel(int n, int ntiles, const int4* __restrict__ meta, const int* __restrict__ rowptr, const int* __restrict__ col,
                  const double* __restrict__ val, const int* __restrict__ lvlptr, int nlev, double* x, const double* __restrict__ b,
                  double omega, int sor, int backward, int opaque_zero) {
  extern __shared__ __align__(128) unsigned char gs_smem[];
  GsCtaStage* st = reinterpret_cast<GsCtaStage*>(gs_smem);
  double* xs = reinterpret_cast<double*>(gs_smem + kStages * sizeof(GsCtaStage));
  __shared__ __align__(8) uint64_t full[kStages];
  const int tid = threadIdx.x;
  constexpr int G = kGsCtaThreads / T;
  const int g = tid / T, lane = tid % T;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  if (XS)
    for (int i = tid; i < n; i += kGsCtaThreads) xs[i] = __ldcg(x + i);
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s)
      if (s < ntiles) gs_cta_issue(st[s], &full[s], __ldg(meta + (backward ? ntiles - 1 - s : s)), rowptr, col, val, b);
  }
  int w = backward ? nlev - 1 : 0;   // current wavefront (forward numbering)
  int s = 0;
  uint32_t parity = 0;
  for (int i = 0; i < ntiles; ++i) {
    const int4 m = __ldg(meta + (backward ? ntiles - 1 - i : i));
    const int ka = m.z & ~3, ra = m.x & ~3;
    mbar_wait(&full[s], parity);
    const GsCtaStage& S = st[s];
    int r = backward ? m.y : m.x;   // forward: next row to relax; backward: one past it
    while (backward ? r > m.x : r < m.y) {
      int a, e;   // segment [a, e): the part of wavefront w inside this tile
      if (!backward) {
        while (__ldg(lvlptr + w + 1) <= r) ++w;
        a = r;
        e = min(m.y, __ldg(lvlptr + w + 1));
      } else {
        while (__ldg(lvlptr + w) >= r) --w;
        e = r;
        a = max(m.x, __ldg(lvlptr + w));
      }
      for (int base = a; base < e; base += G) {   // uniform trip count: shuffles stay converged
        const int row = base + g;
        double rsum = 0.0, d = 0.0;
        if (XS) {
          if (row < e) {
            const int ks = S.rp[row - ra] - ka, ke = S.rp[row - ra + 1] - ka;
            for (int k = ks + lane; k < ke; k += T) {
              const int c = S.col[k];
              const double v = S.val[k];
              if (c == row) d = v;
              else rsum = __dadd_rn(rsum, __dmul_rn(v, xs[c]));
            }
          }
        } else {
          // x through L2: all gathers of the row leave in one burst (one round trip per wavefront, not one per entry)
          int ks = 0, ke = 0;
          if (row < e) {
            ks = S.rp[row - ra] - ka;
            ke = S.rp[row - ra + 1] - ka;
          }
          int c[kGsPrefetch];
          double v[kGsPrefetch], xv[kGsPrefetch];
          const double* a[kGsPrefetch];
#pragma unroll
          for (int j = 0; j < kGsPrefetch; ++j) {
            const int k = ks + lane + j * T;
            const bool in = k < ke;
            c[j] = in ? S.col[k] : -1;
            v[j] = in ? S.val[k] : 0.0;
            a[j] = (in && c[j] != row) ? x + c[j] : &g_gs_zero;
          }
          ldcg_burst8(xv, a);
          pin_burst8(xv, opaque_zero);
#pragma unroll
          for (int j = 0; j < kGsPrefetch; ++j) {
            if (c[j] == row && c[j] >= 0) d = v[j];
            else rsum = __dadd_rn(rsum, __dmul_rn(v[j], xv[j]));
          }
          for (int k = ks + lane + kGsPrefetch * T; k < ke; k += T) {
            const int cc = S.col[k];
            const double vv = S.val[k];
            if (cc == row) d = vv;
            else rsum = __dadd_rn(rsum, __dmul_rn(vv, __ldcg(x + cc)));
          }
        }
        if (T > 1) {
          rsum = lanes_sum<T>(rsum);
          d = lanes_sum<T>(d);
        }
        if (row < e && lane == 0 && d != 0.0) {
          const double rr = __dsub_rn(S.b[row - ra], rsum);
          const double xold = XS ? xs[row] : (sor ? __ldcg(x + row) : 0.0);
          const double xnew = sor ? __dadd_rn(__dmul_rn(1.0 - omega, xold), __dmul_rn(__ddiv_rn(omega, d), rr)) : __ddiv_rn(rr, d);
          if (XS) xs[row] = xnew;
          else __stcg(x + row, xnew);
        }
      }
      __syncthreads();   // the wavefront (segment) is relaxed: its x is visible to the next one
      r = backward ? a : e;
    }
    // every thread is past the barrier that closed the tile's last segment: refill the stage
    if (tid == 0 && i + kStages < ntiles)
      gs_cta_issue(st[s], &full[s], __ldg(meta + (backward ? ntiles - 1 - (i + kStages) : i + kStages)), rowptr, col, val, b);
    if (++s == kStages) { s = 0; parity ^= 1u; }
  }
  if (XS) {
    __syncthreads();
    for (int i = tid; i < n; i += kGsCtaThreads) x[i] = xs[i];
  }
}


// =============================================================================================
// Gauss-Seidel / SOR on LARGE levels: the mailbox protocol (gs_mail_kernel) fed by the TMA tile ring.
// Persistent CTAs own every gridDim-th tile of a wavefront-ALIGNED tile plan (a tile never crosses a
// wavefront boundary, so rows of a tile are mutually independent); tile data — row pointers, column
// indices, values, b — arrives in shared memory through a 3-stage bulk-copy ring, so between two
// tiles a CTA pays no ticket atomic and no dependent load chain.  Per row: later-ordered neighbours
// are read from x (old values), earlier-ordered ones are polled from their 16-byte mailboxes in ONE
// burst, the row is relaxed in the reference's accumulation order and published.
// Deadlock freedom: tiles are claimed by ticket in sweep order (three ahead, to keep the ring full), so
// a tile is only ever held by a resident CTA that walks its claims in order, and the lowest unfinished
// tile depends only on finished ones — for any grid size.
// =============================================================================================
constexpr int kGsTileThreads = 256;

template <int T>
__device__ __forceinline__ double group_lanes_sum(double v, unsigned mask) {
#pragma unroll
  for (int o = T / 2; o > 0; o >>= 1) v += __shfl_down_sync(mask, v, o, T);
  return v;
}

// meta[t] = {first row, end row, first nnz, end nnz}; tile_wave[t] = wavefront (forward numbering)
template <int T>
__global__ void __launch_bounds__(kGsTileThreads, 2)
    gs_tile_kernel(int ntiles, const int4* __restrict__ meta, const int* __restrict__ tile_wave, int nlev, unsigned* ctl,
                   const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, double* x,
                   const double* __restrict__ b, uint4* mail, double omega, int sor, int backward, int opaque_zero,
                   int poll_sleep, int gate_sleep, int poll_masked, int gate_dist, unsigned long long* __restrict__ dbg) {
  extern __shared__ __align__(128) unsigned char gs_tile_smem[];
  GsCtaStage* st = reinterpret_cast<GsCtaStage*>(gs_tile_smem);
  __shared__ __align__(8) uint64_t full[kStages];
  const int tid = threadIdx.x, g = tid / T, lane = tid % T;
  __shared__ int s_tile[kStages];   // tiles claimed (by ticket, in sweep order) for the stages of the ring
  const unsigned e = ld_relaxed_u32(ctl + 1);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      const int t = (int)atomicAdd(&ctl[0], 1u);
      s_tile[s] = t;
      if (t < ntiles) gs_cta_issue(st[s], &full[s], __ldg(meta + (backward ? ntiles - 1 - t : t)), rowptr, col, val, b);
    }
  }
  __syncthreads();
  int s = 0;
  uint32_t parity = 0;
  for (;;) {
    const int t = s_tile[s];
    if (t >= ntiles) break;   // tickets are handed out in order: the first invalid one ends this CTA's work
    int tnext = 0;
    if (tid == 0) tnext = (int)atomicAdd(&ctl[0], 1u);   // the ticket that will refill this stage, requested early
    const int tile = backward ? ntiles - 1 - t : t;
    const int4 m = __ldg(meta + tile);
    const int wf = __ldg(tile_wave + tile);
    const int w = backward ? nlev - 1 - wf : wf;   // wavefront in sweep order
    unsigned long long* stamp = (dbg && tid == 0) ? dbg + 8 * (size_t)tile : nullptr;   // diagnostics (gs_timeline): thread 0's view
    unsigned long long polls = 0;
    if (stamp) { stamp[0] = global_ns(); stamp[7] = (unsigned long long)w; }
    const int ka = m.z & ~3, ra = m.x & ~3;
    constexpr int G = kGsTileThreads / T;   // rows relaxed per pass
    const int nrows = m.y - m.x;
    const double xold0 = (g < nrows && lane == 0) ? __ldcg(x + m.x + g) : 0.0;   // first pass: requested before the waits
    mbar_wait(&full[s], parity);
    if (stamp) stamp[1] = global_ns();
    const GsCtaStage& S = st[s];
    for (int rbase = 0; rbase < nrows; rbase += G) {   // rows of a tile are mutually independent
      const bool active = rbase + g < nrows;
      const int row = active ? m.x + rbase + g : -1;
      const double xold = rbase == 0 ? xold0 : ((active && lane == 0) ? __ldcg(x + row) : 0.0);
      int ks = 0, ke = 0;
      if (active) {
        ks = S.rp[row - ra] - ka;
        ke = S.rp[row - ra + 1] - ka;
      }
      int c[kGsPrefetch];
      double v[kGsPrefetch];
#pragma unroll
      for (int j = 0; j < kGsPrefetch; ++j) {
        const int k = ks + lane + j * T;
        const bool in = k < ke;
        c[j] = in ? S.col[k] : -1;
        v[j] = in ? S.val[k] : 0.0;
      }
      double xn[kGsPrefetch];
      unsigned need = 0u;
      {
        const double* a[kGsPrefetch];
#pragma unroll
        for (int j = 0; j < kGsPrefetch; ++j) {
          const bool valid = c[j] >= 0 && c[j] != row;
          const bool earlier = valid && (backward ? c[j] > row : c[j] < row);
          if (earlier) need |= 1u << j;
          a[j] = (valid && !earlier) ? x + c[j] : &g_gs_zero;
        }
        ldcg_burst8(xn, a);   // old values of later-ordered neighbours: in flight while we wait / poll below
      }
      if (rbase == 0 && w >= gate_dist) {   // throttle: stay off the mailboxes until wavefront w - gate_dist (2) has begun to finish
        if (tid == 0) {
          const unsigned* hint = ctl + (size_t)(2 + w - gate_dist) * kGsCounterStride;
          while (ld_relaxed_u32(hint) != e)
            if (gate_sleep) __nanosleep(gate_sleep);
        }
        __syncthreads();
      }
      if (stamp && rbase == 0) stamp[2] = global_ns();
      while (need) {
        ++polls;
        const uint4* a[kGsPrefetch];
        uint4 mm[kGsPrefetch];
#pragma unroll
        for (int j = 0; j < kGsPrefetch; ++j) {
          a[j] = ((need >> j) & 1u) ? mail + c[j] : mail + (row >= 0 ? row : 0);
          mm[j] = make_uint4(0u, ~e, 0u, ~e);
        }
        if (poll_masked) ld_mail8_masked(mm, a, need);   // only the mailboxes still waited for: no poll traffic for the rest
        else ld_mail8(mm, a);
        {
          unsigned dep = mm[0].y;
#pragma unroll
          for (int j = 1; j < kGsPrefetch; ++j) dep &= mm[j].y;
          dep &= (unsigned)opaque_zero;
#pragma unroll
          for (int j = 0; j < kGsPrefetch; ++j) mm[j].y |= dep;
        }
#pragma unroll
        for (int j = 0; j < kGsPrefetch; ++j)
          if (((need >> j) & 1u) && mm[j].y == e && mm[j].w == e) {
            xn[j] = __hiloint2double((int)mm[j].z, (int)mm[j].x);
            need &= ~(1u << j);
          }
        if (need && poll_masked >= 2) {
          // focused spin (poll_masked = 2; parity-tested, not yet timed on hardware): a full round is ~175 instructions per warp, which
          // with 16 polling warps per SM is about as long as the L2 round trip itself; wait for ONE outstanding mailbox in
          // a five-instruction loop instead, then let the next masked round collect whatever else has arrived meanwhile
          const int jsel = __ffs((int)need) - 1;
          int csel = c[0];
#pragma unroll
          for (int j = 1; j < kGsPrefetch; ++j) csel = (jsel == j) ? c[j] : csel;
          const uint4* p = mail + csel;
          uint4 q;
          do {
            asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p) : "memory");
            ++polls;
          } while (q.y != e || q.w != e);
          const double got = __hiloint2double((int)q.z, (int)q.x);
#pragma unroll
          for (int j = 0; j < kGsPrefetch; ++j)
            if (jsel == j) xn[j] = got;
          need &= ~(1u << jsel);
        }
        if (need && poll_sleep) __nanosleep(poll_sleep);
      }
      if (stamp && rbase == 0) { stamp[3] = global_ns(); stamp[6] = polls; }
      __syncwarp();   // lanes leave the poll loop at different times: reconverge before the arithmetic
      double rsum = 0.0, d = 0.0;
#pragma unroll
      for (int j = 0; j < kGsPrefetch; ++j) {
        if (c[j] == row && c[j] >= 0) d = v[j];
        else rsum = __dadd_rn(rsum, __dmul_rn(v[j], xn[j]));
      }
      for (int k = ks + lane + kGsPrefetch * T; k < ke; k += T) {   // rows longer than T * kGsPrefetch
        const int cc = S.col[k];
        const double vv = S.val[k];
        if (cc == row) { d = vv; continue; }
        double xv;
        if (backward ? cc > row : cc < row) {
          uint4 q;
          do {
            asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                         : "l"(mail + cc)
                         : "memory");
          } while (q.y != e || q.w != e);
          xv = __hiloint2double((int)q.z, (int)q.x);
        } else {
          xv = __ldcg(x + cc);
        }
        rsum = __dadd_rn(rsum, __dmul_rn(vv, xv));
      }
      if (T > 1) {   // rows of a tile are independent (wavefront-aligned tiles): whole-warp shuffles cannot deadlock
        __syncwarp();
        rsum = group_lanes_sum<T>(rsum, 0xffffffffu);
        d = group_lanes_sum<T>(d, 0xffffffffu);
      }
      if (active && lane == 0) {
        double xnew = xold;
        if (d != 0.0) {
          const double r = __dsub_rn(S.b[row - ra], rsum);
          xnew = sor ? __dadd_rn(__dmul_rn(1.0 - omega, xold), __dmul_rn(__ddiv_rn(omega, d), r)) : __ddiv_rn(r, d);
        }
        st_mail(mail + row, xnew, e);
        __stcg(x + row, xnew);
        if (stamp && rbase == 0) stamp[4] = global_ns();
      }
    }
    __syncthreads();   // stage s is free; the tile is published
    if (stamp) stamp[5] = global_ns();
    if (tid == 0) {
      volatile unsigned* mine = ctl + (size_t)(2 + w) * kGsCounterStride;
      *mine = e;   // throttle hint only
      s_tile[s] = tnext;   // read again kStages tiles (>= kStages barriers) from now
      if (tnext < ntiles) gs_cta_issue(st[s], &full[s], __ldg(meta + (backward ? ntiles - 1 - tnext : tnext)), rowptr, col, val, b);
    }
    if (++s == kStages) { s = 0; parity ^= 1u; }
  }
}

}  // namespace b200amg
