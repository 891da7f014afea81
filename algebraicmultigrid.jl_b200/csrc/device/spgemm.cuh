// Sparse matrix product C = A * B on the device — the Galerkin products R*A and (R*A)*P of the setup phase
// (`src/classical.jl:46`, `src/aggregation.jl:145`), with the semantics of the stdlib product the
// reference calls and of the host restatement (csrc/host/amg_setup.cpp: amgsetup_spgemm_begin): compressed-sparse-column
// in and out, sorted row indices inside a column, structural zeros from cancellation KEPT, and — so that a hierarchy
// built with it is bit-identical to the host-built one — the same accumulation order per entry:
// C(r, j) = sum over the entries k of B(:, j) in storage order, over the entries r of A(:, k) in storage order, first
// product assigned, later ones added, no FMA.
//
// One thread owns one output column.  Its distinct rows are collected in an open-addressing hash table that lives in a
// global scratch segment sized from the column's product count (power of two, load factor <= 1/2); a single thread
// inserts in the reference's order, so the sums are deterministic.  A second kernel compacts every table into the
// output column and sorts it by row index (columns of these products have tens to a few hundred entries).  Columns are
// processed in batches that fit the scratch budget; this is throughput work (millions of independent columns, random
// 12-byte accesses served by L2 / HBM), not a latency chain.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200amg {

constexpr int kSpgemmThreads = 128;

// products of column j = sum of the lengths of the columns of A that B(:, j) selects (upper bound of its distinct rows)
__global__ void spgemm_products_kernel(int64_t n, const int* __restrict__ Ap, const int* __restrict__ Bp, const int* __restrict__ Bj,
                                       long long* __restrict__ prod) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  long long p = 0;
  for (int bp = Bp[j]; bp < Bp[j + 1]; ++bp) {
    const int k = Bj[bp];
    p += Ap[k + 1] - Ap[k];
  }
  prod[j] = p;
}

// columns [j0, j0 + count): table of column j0 + t at keys/vals[off[t] .. off[t] + cap[t]), cap a power of two (or 0)
__global__ void spgemm_hash_kernel(int64_t j0, int64_t count, const int* __restrict__ Ap, const int* __restrict__ Aj,
                                   const double* __restrict__ Ax, const int* __restrict__ Bp, const int* __restrict__ Bj,
                                   const double* __restrict__ Bx, const long long* __restrict__ off, const int* __restrict__ cap,
                                   int* __restrict__ keys, double* __restrict__ vals, int* __restrict__ ucount) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int64_t j = j0 + t;
  const unsigned mask = (unsigned)cap[t] - 1u;
  int* K = keys + off[t];
  double* V = vals + off[t];
  int u = 0;
  for (int bp = Bp[j]; bp < Bp[j + 1]; ++bp) {
    const int k = Bj[bp];
    const double bv = Bx[bp];
    for (int ap = Ap[k]; ap < Ap[k + 1]; ++ap) {
      const int r = Aj[ap];
      const double prod = __dmul_rn(Ax[ap], bv);
      unsigned h = ((unsigned)r * 2654435761u) & mask;
      for (;;) {
        const int key = K[h];
        if (key == r) { V[h] = __dadd_rn(V[h], prod); break; }
        if (key < 0) { K[h] = r; V[h] = prod; ++u; break; }
        h = (h + 1u) & mask;
      }
    }
  }
  ucount[t] = u;
}

// compact the table of every column into C (cptr[t] = first output slot of column j0 + t inside this batch) and sort by row
__global__ void spgemm_emit_kernel(int64_t count, const long long* __restrict__ off, const int* __restrict__ cap,
                                   const int* __restrict__ keys, const double* __restrict__ vals, const long long* __restrict__ cptr,
                                   int* __restrict__ Cj, double* __restrict__ Cx) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int c = cap[t];
  const int* K = keys + off[t];
  const double* V = vals + off[t];
  int* oj = Cj + cptr[t];
  double* ox = Cx + cptr[t];
  int u = 0;
  for (int h = 0; h < c; ++h) {
    const int key = K[h];
    if (key < 0) continue;
    const double v = V[h];
    int q = u - 1;   // insertion sort while compacting
    while (q >= 0 && oj[q] > key) { oj[q + 1] = oj[q]; ox[q + 1] = ox[q]; --q; }
    oj[q + 1] = key;
    ox[q + 1] = v;
    ++u;
  }
}

}  // namespace b200amg
