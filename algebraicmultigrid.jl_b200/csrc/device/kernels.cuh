// Hand-written sm_100a kernels for the AMG solve phase.  All of them are HBM-bound sparse or
// vector streams (fp64 values, int32 indices): no tensor cores.  The rules that matter here are
// coalesced matrix streams, enough bytes in flight per SM, x gathers served by L1/L2, and grids
// sized to cover 148 SMs.  Reference semantics cited per kernel (paths under /root/reference).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200amg {

constexpr int kThreads = 256;
constexpr int kNumSM = 148;   // B200; only sizes compile-time tables (kRedBlocks) — grids use the SM count queried at b200amg_create

template <int T>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = T / 2; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, T);
  return v;
}

// ---------------------------------------------------------------------------------------------
// CSR "vector" kernel: T lanes cooperate on one row (T = 2..32 chosen per matrix from its mean row
// length), consecutive rows on consecutive lane groups so the value / column streams are read
// fully coalesced.  MODE 0: y = A x          (mul!(res, A, x)            multilevel.jl:188,219)
//                   MODE 1: y = b - A x      (res .= b .- res fused      multilevel.jl:189,220)
//                   MODE 2: y += A x         (mul!(res,P,cx); x .+= res  multilevel.jl:233-234)
// Also used for restriction (A := R by rows, multilevel.jl:223).
// ---------------------------------------------------------------------------------------------
template <int T, int MODE>
__global__ void __launch_bounds__(kThreads) csr_vec_kernel(int64_t nrows, const int* __restrict__ rowptr,
                                                           const int* __restrict__ col,
                                                           const double* __restrict__ val,
                                                           const double* __restrict__ x,
                                                           const double* __restrict__ b, double* __restrict__ y) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = gid / T;
  const int lane = (int)(gid % T);
  double sum = 0.0;
  if (row < nrows) {
    const int s = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
    for (int k = s + lane; k < e; k += T) sum += __ldg(val + k) * __ldg(x + __ldg(col + k));
  }
  sum = group_sum<T>(sum);
  if (lane == 0 && row < nrows) {
    if (MODE == 0) y[row] = sum;
    else if (MODE == 1) y[row] = b[row] - sum;
    else y[row] += sum;
  }
}

// ---------------------------------------------------------------------------------------------
// Damped Jacobi, "fast" (Hermitian) variant: smooth!(x, ::FastJacobiSmoother, b) smoother.jl:113-141.
// The matrix walked is the one whose ROW i is the reference's CSC COLUMN i.  The reference copies
// x -> temp and relaxes from temp; here xin/xout are distinct buffers (no copy pass).
//   x[i] = diag == 0 ? x[i] : (1-w) x[i] + w (b[i] - sum_{j != i} a_ij x[j]) / diag
// ---------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(kThreads) jacobi_fast_kernel(int64_t nrows, const int* __restrict__ rowptr,
                                                               const int* __restrict__ col,
                                                               const double* __restrict__ val,
                                                               const double* __restrict__ xin,
                                                               const double* __restrict__ b,
                                                               double* __restrict__ xout, double omega) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = gid / T;
  const int lane = (int)(gid % T);
  double rsum = 0.0, diag = 0.0;
  if (row < nrows) {
    const int s = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
    for (int k = s + lane; k < e; k += T) {
      const int c = __ldg(col + k);
      const double v = __ldg(val + k);
      if (c == row) diag = v; else rsum += v * __ldg(xin + c);
    }
  }
  rsum = group_sum<T>(rsum);
  diag = group_sum<T>(diag);
  if (lane == 0 && row < nrows) {
    const double xi = xin[row];
    xout[row] = (diag == 0.0) ? xi : (1.0 - omega) * xi + omega * ((b[row] - rsum) / diag);
  }
}

// Jacobi with a zero current iterate (the coarse-level pre-smoother always starts from
// coarse_x .= 0, multilevel.jl:226): every product val*0 is an exact zero, so
// x = w * (b / diag) needs the diagonal only.  diag = 0 rows keep x = 0.
__global__ void __launch_bounds__(kThreads) jacobi_zero_guess_kernel(int64_t n, const double* __restrict__ diag,
                                                                     const double* __restrict__ b,
                                                                     double* __restrict__ x, double omega, int general) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d = diag[i];
  // fast: (1-w)*0 + w*((b-0)/d) ; general (smoother.jl:157-171): 0 - w*(0 - b)/d
  x[i] = (d == 0.0) ? 0.0 : (general ? (0.0 - omega * (0.0 - b[i]) / d) : ((1.0 - omega) * 0.0 + omega * ((b[i] - 0.0) / d)));
}

// Jacobi, NoSymmetry variant: smooth!(x, ::JacobiSmoother, b) smoother.jl:157-171 on the TRUE rows of A
// with precomputed diagvals:  x[i] -= w (A x - b)[i] / d_i   (skipped where d_i == 0)
template <int T>
__global__ void __launch_bounds__(kThreads) jacobi_general_kernel(int64_t nrows, const int* __restrict__ rowptr,
                                                                  const int* __restrict__ col,
                                                                  const double* __restrict__ val,
                                                                  const double* __restrict__ diagvals,
                                                                  const double* __restrict__ xin,
                                                                  const double* __restrict__ b,
                                                                  double* __restrict__ xout, double omega) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = gid / T;
  const int lane = (int)(gid % T);
  double sum = 0.0;
  if (row < nrows) {
    const int s = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
    for (int k = s + lane; k < e; k += T) sum += __ldg(val + k) * __ldg(xin + __ldg(col + k));
  }
  sum = group_sum<T>(sum);
  if (lane == 0 && row < nrows) {
    const double d = diagvals[row];
    const double xi = xin[row];
    xout[row] = (d != 0.0) ? xi - omega * (sum - b[row]) / d : xi;
  }
}

// ---------------------------------------------------------------------------------------------
// Gauss-Seidel / SOR with the reference's exact lexicographic semantics (gs! smoother.jl:73-90,
// sor_step! :205-221, and the NoSymmetry L\(b-Ux) forms :421-582 which are the same row update on
// the true A).  Rows are grouped into wavefronts (level schedule built at upload from the
// symmetrised pattern): all rows of a wavefront are mutually independent, every earlier-ordered
// neighbour is in an earlier wavefront and every later-ordered neighbour in a later one, so
// relaxing wavefront by wavefront gives the sequential sweep's result up to summation order
// inside a row.  x is read with plain (coherent) loads: it is updated in place.
// ---------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ void relax_row(int row, int lane, const int* __restrict__ rowptr,
                                          const int* __restrict__ col, const double* __restrict__ val,
                                          double* x, const double* __restrict__ b, double omega, int sor) {
  double rsum = 0.0, d = 0.0;
  if (row >= 0) {
    const int s = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
    for (int k = s + lane; k < e; k += T) {
      const int c = __ldg(col + k);
      const double v = __ldg(val + k);
      if (c == row) d = v; else rsum += v * x[c];
    }
  }
  rsum = group_sum<T>(rsum);
  d = group_sum<T>(d);
  if (lane == 0 && row >= 0 && d != 0.0) {
    const double bi = b[row];
    x[row] = sor ? (1.0 - omega) * x[row] + (omega / d) * (bi - rsum) : (bi - rsum) / d;
  }
}

// one wavefront spread over the whole grid
template <int T>
__global__ void __launch_bounds__(kThreads) gs_wavefront_kernel(const int* __restrict__ rows, int count,
                                                                const int* __restrict__ rowptr,
                                                                const int* __restrict__ col,
                                                                const double* __restrict__ val, double* x,
                                                                const double* __restrict__ b, double omega, int sor) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t idx = gid / T;
  const int lane = (int)(gid % T);
  const int row = idx < count ? __ldg(rows + idx) : -1;
  relax_row<T>(row, lane, rowptr, col, val, x, b, omega, sor);
}

// a run of consecutive NARROW wavefronts relaxed by ONE CTA with __syncthreads() between them:
// replaces (lv_end - lv_begin) dependent launches by one.  lvlptr indexes `rows`.
constexpr int kCtaThreads = 1024;
template <int T>
__global__ void __launch_bounds__(kCtaThreads) gs_cta_levels_kernel(const int* __restrict__ rows,
                                                                    const int* __restrict__ lvlptr, int lv_begin,
                                                                    int lv_end, const int* __restrict__ rowptr,
                                                                    const int* __restrict__ col,
                                                                    const double* __restrict__ val, double* x,
                                                                    const double* __restrict__ b, double omega, int sor) {
  const int lane = threadIdx.x % T;
  const int grp = threadIdx.x / T;
  constexpr int kGroups = kCtaThreads / T;
  for (int lv = lv_begin; lv < lv_end; ++lv) {
    const int s = __ldg(lvlptr + lv), e = __ldg(lvlptr + lv + 1);
    for (int base = s; base < e; base += kGroups) {   // uniform trip count per warp: shuffles stay converged
      const int idx = base + grp;
      const int row = idx < e ? __ldg(rows + idx) : -1;
      relax_row<T>(row, lane, rowptr, col, val, x, b, omega, sor);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// reductions: deterministic two-stage dot / sum of squares (norm(b), norm(res) multilevel.jl:170,190;
// the dots of the device PCG).  Stage 1 writes one partial per CTA, stage 2 folds them in one CTA.
// ---------------------------------------------------------------------------------------------
constexpr int kRedBlocks = kNumSM * 4;

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[32];
  v = group_sum<32>(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (w == 0) v = group_sum<32>(v);
  __syncthreads();
  return v;
}

__global__ void __launch_bounds__(kThreads) dot_partial_kernel(int64_t n, const double* __restrict__ a,
                                                               const double* __restrict__ b,
                                                               double* __restrict__ partial) {
  double s = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s += a[i] * b[i];
  s = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out[0] = sum(partial) ; if take_sqrt, out[0] = sqrt(sum)
__global__ void __launch_bounds__(kThreads) reduce_final_kernel(int nparts, const double* __restrict__ partial,
                                                                double* __restrict__ out, int take_sqrt) {
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[i];
  s = block_sum(s);
  if (threadIdx.x == 0) out[0] = take_sqrt ? sqrt(s) : s;
}

// ---------------------------------------------------------------------------------------------
// coarse solve apply: x = M b, M dense n x n column-major ((p::Pinv)(x,b) = mul!(x, pinvA, b),
// coarse_solver.jl:16).  One thread per row; consecutive threads read consecutive addresses.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) dense_gemv_kernel(int n, const double* __restrict__ M,
                                                              const double* __restrict__ b, double* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int j = 0; j < n; ++j) s += M[i + (int64_t)j * n] * __ldg(b + j);
  x[i] = s;
}

// ---------------------------------------------------------------------------------------------
// vector updates of the device PCG (IterativeSolvers' PCGIterable restated): scalars are read from
// device memory so the whole iteration is enqueued without a host round trip.
// ---------------------------------------------------------------------------------------------
// u = c + (rho/rho_prev) u
__global__ void __launch_bounds__(kThreads) pcg_update_u_kernel(int64_t n, const double* __restrict__ c,
                                                                double* __restrict__ u,
                                                                const double* __restrict__ rho,
                                                                const double* __restrict__ rho_prev) {
  const double beta = rho[0] / rho_prev[0];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) u[i] = c[i] + beta * u[i];
}
// alpha = rho / (u.q) ; x += alpha u ; r -= alpha q
__global__ void __launch_bounds__(kThreads) pcg_update_xr_kernel(int64_t n, double* __restrict__ x,
                                                                 double* __restrict__ r,
                                                                 const double* __restrict__ u,
                                                                 const double* __restrict__ q,
                                                                 const double* __restrict__ rho,
                                                                 const double* __restrict__ uq) {
  const double alpha = rho[0] / uq[0];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    x[i] += alpha * u[i];
    r[i] -= alpha * q[i];
  }
}

__global__ void __launch_bounds__(kThreads) set_scalar_kernel(double* p, double v) {
  if (threadIdx.x == 0 && blockIdx.x == 0) p[0] = v;
}
__global__ void __launch_bounds__(kThreads) copy_scalar_kernel(double* dst, const double* src) {
  if (threadIdx.x == 0 && blockIdx.x == 0) dst[0] = src[0];
}

}  // namespace b200amg
