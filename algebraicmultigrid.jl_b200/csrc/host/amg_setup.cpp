// Host-side hierarchy construction (setup phase) for the B200 AMG engine.
//
// The north star keeps setup on the host: "The hierarchy (setup in src/classical.jl /
// src/aggregation.jl) is built once on the host from the reference path and uploaded".
// In the reference that host code is Julia; Julia does not exist in this image, so the
// host side of this repo carries its own implementation of the same algorithms, with the
// same semantics (file:line citations below are relative to /root/reference), so that
// the hierarchies it uploads are the ones the reference would have built. It is pinned
// against the reference's own golden tests in tests/test_setup_goldens.py.
//
// Conventions: compressed-sparse-column, 0-based int32 indices, fp64 values, row indices
// sorted ascending inside every column (what Julia's SparseMatrixCSC holds, minus the
// 1-based offset). All entry points are extern "C", operate on caller-provided buffers
// and return 0 on success / a negative status.  None of this code runs in the solve phase.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <limits>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int32_t idx_t;

static inline double wall_s() {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  return (double)std::clock() / CLOCKS_PER_SEC;
#endif
}

extern "C" {

// ---------------------------------------------------------------------------------------
// gallery: N-D Poisson, (2N+1)-point stencil, first index fastest.  src/gallery.jl:1-63
// Column j holds rows j-s_{N-1}, .., j-s_0, j, j+s_0, .., j+s_{N-1} (those inside the box),
// which is already ascending.  Pass colptr == nullptr to get nnz only.
// ---------------------------------------------------------------------------------------
int64_t amgsetup_poisson_nnz(int ndim, const int64_t* dims) {
  int64_t n = 1;
  for (int d = 0; d < ndim; ++d) n *= dims[d];
  int64_t nnz = n;
  for (int d = 0; d < ndim; ++d) nnz += 2 * (n / dims[d]) * (dims[d] - 1);
  return nnz;
}

int amgsetup_poisson(int ndim, const int64_t* dims, idx_t* colptr, idx_t* rowval, double* nzval) {
  if (ndim < 1 || ndim > 8) return -1;
  int64_t n = 1;
  int64_t stride[8];
  for (int d = 0; d < ndim; ++d) { stride[d] = n; n *= dims[d]; }
  if (amgsetup_poisson_nnz(ndim, dims) > std::numeric_limits<idx_t>::max()) return -2;
  // pass 1: column counts (parallel), pass 2: fill
  colptr[0] = 0;
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; ++j) {
    int cnt = 1;
    int64_t rem = j;
    for (int d = 0; d < ndim; ++d) {
      int64_t c = rem % dims[d]; rem /= dims[d];
      cnt += (c > 0) + (c + 1 < dims[d]);
    }
    colptr[j + 1] = cnt;
  }
  for (int64_t j = 0; j < n; ++j) colptr[j + 1] += colptr[j];
  const double center = 2.0 * ndim;
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; ++j) {
    int64_t coord[8];
    int64_t rem = j;
    for (int d = 0; d < ndim; ++d) { coord[d] = rem % dims[d]; rem /= dims[d]; }
    idx_t p = colptr[j];
    for (int d = ndim - 1; d >= 0; --d)
      if (coord[d] > 0) { rowval[p] = (idx_t)(j - stride[d]); nzval[p] = -1.0; ++p; }
    rowval[p] = (idx_t)j; nzval[p] = center; ++p;
    for (int d = 0; d < ndim; ++d)
      if (coord[d] + 1 < dims[d]) { rowval[p] = (idx_t)(j + stride[d]); nzval[p] = -1.0; ++p; }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// transpose (copy(A')) for real matrices; output columns sorted.  m x n in, n x m out.
// ---------------------------------------------------------------------------------------
// Built on all cores: entries reach their output column through an atomic cursor (any order), then every
// output column is sorted by index (columns are short) — the result is the sequential counting-sort
// transpose, bit for bit.
int amgsetup_transpose(int64_t m, int64_t n, const idx_t* colptr, const idx_t* rowval,
                       const double* nzval, idx_t* tcolptr, idx_t* trowval, double* tnzval) {
  const int64_t nnz = colptr[n];
  std::fill(tcolptr, tcolptr + m + 1, 0);
  idx_t* cnt = tcolptr + 1;
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nnz; ++k) __atomic_fetch_add(&cnt[rowval[k]], 1, __ATOMIC_RELAXED);
  for (int64_t i = 0; i < m; ++i) tcolptr[i + 1] += tcolptr[i];
  std::vector<idx_t> next(tcolptr, tcolptr + m);
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; ++j)
    for (idx_t k = colptr[j]; k < colptr[j + 1]; ++k) {
      const idx_t q = __atomic_fetch_add(&next[rowval[k]], 1, __ATOMIC_RELAXED);
      trowval[q] = (idx_t)j;
      if (nzval) tnzval[q] = nzval[k];
    }
#pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t i = 0; i < m; ++i) {
    const idx_t b = tcolptr[i], e = tcolptr[i + 1];
    for (idx_t k = b + 1; k < e; ++k) {
      const idx_t ci = trowval[k];
      const double cv = nzval ? tnzval[k] : 0.0;
      idx_t j = k - 1;
      while (j >= b && trowval[j] > ci) {
        trowval[j + 1] = trowval[j];
        if (nzval) tnzval[j + 1] = tnzval[j];
        --j;
      }
      trowval[j + 1] = ci;
      if (nzval) tnzval[j + 1] = cv;
    }
  }
  return 0;
}

// 1 if the CSC arrays describe a matrix that is bit-for-bit equal to its transpose.
int amgsetup_is_bitsymmetric(int64_t n, const idx_t* colptr, const idx_t* rowval, const double* nzval) {
  const int64_t nnz = colptr[n];
  std::vector<idx_t> tp(n + 1), tr(nnz);
  std::vector<double> tv(nnz);
  amgsetup_transpose(n, n, colptr, rowval, nzval, tp.data(), tr.data(), tv.data());
  if (std::memcmp(tp.data(), colptr, sizeof(idx_t) * (n + 1))) return 0;
  if (std::memcmp(tr.data(), rowval, sizeof(idx_t) * nnz)) return 0;
  for (int64_t k = 0; k < nnz; ++k)
    if (!(tv[k] == nzval[k])) return 0;
  return 1;
}

// ---------------------------------------------------------------------------------------
// Classical strength of connection.  src/strength.jl:7-37 (+ helpers :39-70)
// In: At (n x n).  Out: T (same or smaller pattern).  Returns nnz(T) (>= 0) or < 0.
// Caller allocates tcolptr[n+1], trowval/tnzval[nnz(At)].  S = transpose(T) is formed by
// the caller with amgsetup_transpose.
// ---------------------------------------------------------------------------------------
int64_t amgsetup_classical_strength(int64_t n, const idx_t* colptr, const idx_t* rowval,
                                    const double* nzval, double theta,
                                    idx_t* tcolptr, idx_t* trowval, double* tnzval) {
  // pass 1 (all cores): how many entries of column i survive  (:21-32)
  tcolptr[0] = 0;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double mx = 0.0;   // find_max_off_diag  (strength.jl:39-49)
    for (idx_t j = colptr[i]; j < colptr[i + 1]; ++j)
      if (rowval[j] != i) mx = std::max(mx, std::fabs(nzval[j]));
    const double threshold = theta * mx;
    idx_t kept = 0;
    for (idx_t j = colptr[i]; j < colptr[i + 1]; ++j) {
      double v = nzval[j];
      if (rowval[j] != i) v = (std::fabs(v) >= threshold) ? std::fabs(v) : 0.0;
      if (v != 0.0) ++kept;
    }
    tcolptr[i + 1] = kept;
  }
  for (int64_t i = 0; i < n; ++i) tcolptr[i + 1] += tcolptr[i];
  // pass 2 (all cores): fill and scale
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double mx = 0.0;
    for (idx_t j = colptr[i]; j < colptr[i + 1]; ++j)
      if (rowval[j] != i) mx = std::max(mx, std::fabs(nzval[j]));
    const double threshold = theta * mx;
    const idx_t q0 = tcolptr[i];
    idx_t q = q0;
    for (idx_t j = colptr[i]; j < colptr[i + 1]; ++j) {
      double v = nzval[j];
      if (rowval[j] != i) v = (std::fabs(v) >= threshold) ? std::fabs(v) : 0.0;   // :21-27
      if (v != 0.0) { trowval[q] = rowval[j]; tnzval[q] = v; ++q; }                // dropzeros! :32
    }
    // scale_cols_by_largest_entry!  (find_max starts from zero)  :51-70
    double m2 = 0.0;
    for (idx_t j = q0; j < q; ++j) m2 = std::max(m2, tnzval[j]);
    for (idx_t j = q0; j < q; ++j) tnzval[j] /= m2;
  }
  return tcolptr[n];
}

// ---------------------------------------------------------------------------------------
// Symmetric strength of connection.  src/strength.jl:77-122.  bsr_flag && theta == 0 is
// handled by the caller (pattern of A filled with ones).  Returns nnz(S).
// ---------------------------------------------------------------------------------------
int64_t amgsetup_symmetric_strength(int64_t n, const idx_t* colptr, const idx_t* rowval,
                                    const double* nzval, double theta,
                                    idx_t* scolptr, idx_t* srowval, double* snzval) {
  std::vector<double> diags(n);
  for (int64_t i = 0; i < n; ++i) {
    double d = 0.0;
    for (idx_t j = colptr[i]; j < colptr[i + 1]; ++j)
      if (rowval[j] == i) d += nzval[j];
    diags[i] = std::fabs(d);
  }
  idx_t q = 0;
  scolptr[0] = 0;
  for (int64_t i = 0; i < n; ++i) {
    const double eps_Aii = theta * theta * diags[i];
    const idx_t q0 = q;
    for (idx_t j = colptr[i]; j < colptr[i + 1]; ++j) {
      const idx_t row = rowval[j];
      double v = nzval[j];
      if (row != i && v * v < eps_Aii * diags[row]) v = 0.0;
      if (v != 0.0) { srowval[q] = row; snzval[q] = std::fabs(v); ++q; }
    }
    double m2 = 0.0;
    for (idx_t j = q0; j < q; ++j) m2 = std::max(m2, snzval[j]);
    for (idx_t j = q0; j < q; ++j) snzval[j] /= m2;
    scolptr[i + 1] = q;
  }
  return q;
}

// remove_diag! (src/splitting.jl:8-18): zero the diagonal then dropzeros! (which also
// removes any other stored zero).  Works in place; returns new nnz.
int64_t amgsetup_remove_diag(int64_t n, idx_t* colptr, idx_t* rowval, double* nzval) {
  idx_t q = 0, start = 0;
  for (int64_t i = 0; i < n; ++i) {
    const idx_t end = colptr[i + 1];
    for (idx_t j = start; j < end; ++j)
      if (rowval[j] != i && nzval[j] != 0.0) { rowval[q] = rowval[j]; nzval[q] = nzval[j]; ++q; }
    start = end;
    colptr[i + 1] = q;
  }
  return q;
}

// The same filter out of place and on all cores (count, prefix sum, fill): out_colptr[n+1], out_rowval / out_nzval sized
// nnz(in).  out_nzval may be null (pattern only: what the C/F splitting reads).  Returns the new nnz.
int64_t amgsetup_remove_diag_copy(int64_t n, const idx_t* colptr, const idx_t* rowval, const double* nzval,
                                  idx_t* out_colptr, idx_t* out_rowval, double* out_nzval) {
  out_colptr[0] = 0;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    idx_t kept = 0;
    for (idx_t j = colptr[i]; j < colptr[i + 1]; ++j) kept += (rowval[j] != i && nzval[j] != 0.0);
    out_colptr[i + 1] = kept;
  }
  for (int64_t i = 0; i < n; ++i) out_colptr[i + 1] += out_colptr[i];
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    idx_t q = out_colptr[i];
    for (idx_t j = colptr[i]; j < colptr[i + 1]; ++j)
      if (rowval[j] != i && nzval[j] != 0.0) {
        out_rowval[q] = rowval[j];
        if (out_nzval) out_nzval[q] = nzval[j];
        ++q;
      }
  }
  return out_colptr[n];
}

// ---------------------------------------------------------------------------------------
// Ruge-Stuben first-pass C/F splitting.  src/splitting.jl:25-159.
// S: strength with diagonal removed; T = transpose(S).  Only the patterns are used.
// splitting[i] in {0 = F, 1 = C}.  Index arithmetic is kept 1-based internally so the
// bucket bookkeeping is the reference's statement for statement.
// ---------------------------------------------------------------------------------------
int amgsetup_rs_cf_splitting(int64_t n_, const idx_t* Sp, const idx_t* Sj, const idx_t* Tp,
                             const idx_t* Tj, idx_t* splitting) {
  const int64_t n = n_;
  enum : uint8_t { F_NODE = 0, C_NODE = 1, U_NODE = 2 };
  // The pass is sequential and bound by random accesses to n-sized arrays, so the bookkeeping is kept small: 32-bit
  // (n < 2^31 - 2 is guaranteed by the int32 index type of the matrices), lambda and the bucket position of a node side by
  // side in one 8-byte record (one cache line touched per node instead of two), and the C / F / U state in a byte array
  // (the innermost test "is this neighbour still undecided" then runs mostly out of cache) that is widened into
  // `splitting` at the end.  Statement order and index arithmetic (1-based) are the reference's.
  using bk_t = int32_t;
  struct NodeRec { bk_t lambda, pos; };
  std::vector<NodeRec> node(n + 1, NodeRec{0, 0});
  std::vector<bk_t> interval_ptr(n + 2, 0), interval_count(n + 2, 0), index_to_node(n + 1, 0);
  std::vector<uint8_t> state(n + 1, U_NODE);
  // 1-based: node i in 1..n ; arrays indexed 1..n(+1)
  for (int64_t i = 1; i <= n; ++i) {
    node[i].lambda = (bk_t)(Sp[i] - Sp[i - 1]);
    interval_count[node[i].lambda + 1] += 1;
  }
  // accumulate!(+, interval_ptr[2:end], interval_count[1:end-1])
  {
    int64_t acc = 0;
    for (int64_t k = 1; k <= n; ++k) { acc += interval_count[k]; interval_ptr[k + 1] = (bk_t)acc; }
  }
  std::fill(interval_count.begin(), interval_count.end(), 0);
  for (int64_t i = 1; i <= n; ++i) {
    const int64_t lambda_i = node[i].lambda + 1;
    interval_count[lambda_i] += 1;
    const int64_t index = interval_ptr[lambda_i] + interval_count[lambda_i];
    index_to_node[index] = (bk_t)i;
    node[i].pos = (bk_t)index;
  }
  for (int64_t i = 1; i <= n; ++i) state[i] = (node[i].lambda == 0) ? F_NODE : U_NODE;

  // Software prefetch along the bucket order (hints only: a node met at distance D may still be moved before its turn, which
  // costs nothing but the wasted line): records of the node 12 steps ahead, its neighbour lists 8 ahead, and for a node
  // that is still undecided 4 steps ahead the records of the neighbours it would turn into F points.
  for (int64_t top_index = n; top_index >= 1; --top_index) {
    if (top_index > 12) {
      const int64_t i2 = index_to_node[top_index - 12];
      __builtin_prefetch(&node[i2]);
      __builtin_prefetch(&state[i2]);
      __builtin_prefetch(&Sp[i2 - 1]);
      __builtin_prefetch(&Tp[i2 - 1]);
      const int64_t i3 = index_to_node[top_index - 8];
      if (state[i3] != F_NODE) {
        __builtin_prefetch(&Sj[Sp[i3 - 1]]);
        __builtin_prefetch(&Tj[Tp[i3 - 1]]);
      }
      const int64_t i4 = index_to_node[top_index - 4];
      if (state[i4] != F_NODE) {
        for (idx_t j = Sp[i4 - 1]; j < Sp[i4]; ++j) {
          const int64_t row = (int64_t)Sj[j] + 1;
          __builtin_prefetch(&state[row]);
          __builtin_prefetch(&Tp[row - 1]);
        }
        for (idx_t j = Tp[i4 - 1]; j < Tp[i4]; ++j) __builtin_prefetch(&node[(int64_t)Tj[j] + 1]);
      }
      const int64_t i5 = index_to_node[top_index - 2];
      if (state[i5] != F_NODE) {
        for (idx_t j = Sp[i5 - 1]; j < Sp[i5]; ++j) {
          const int64_t row = (int64_t)Sj[j] + 1;
          if (state[row] == U_NODE)
            for (idx_t k = Tp[row - 1]; k < Tp[row]; k += 16) __builtin_prefetch(&Tj[k]);
        }
      }
      const int64_t i6 = index_to_node[top_index - 1];
      if (state[i6] != F_NODE) {
        for (idx_t j = Sp[i6 - 1]; j < Sp[i6]; ++j) {
          const int64_t row = (int64_t)Sj[j] + 1;
          if (state[row] == U_NODE)
            for (idx_t k = Tp[row - 1]; k < Tp[row]; ++k) __builtin_prefetch(&node[(int64_t)Tj[k] + 1]);
        }
      }
    }
    const int64_t i = index_to_node[top_index];
    const int64_t lambda_i = node[i].lambda + 1;
    interval_count[lambda_i] -= 1;
    if (state[i] == F_NODE) continue;
    state[i] = C_NODE;
    for (idx_t j = Sp[i - 1]; j < Sp[i]; ++j) {
      const int64_t row = (int64_t)Sj[j] + 1;
      if (state[row] == U_NODE) {
        state[row] = F_NODE;
        for (idx_t k = Tp[row - 1]; k < Tp[row]; ++k) {
          const int64_t rowk = (int64_t)Tj[k] + 1;
          if (state[rowk] == U_NODE) {
            NodeRec& nk = node[rowk];
            if (nk.lambda >= n - 1) continue;
            const int64_t lambda_k = nk.lambda + 1;
            const int64_t old_pos = nk.pos;
            const int64_t new_pos = interval_ptr[lambda_k] + interval_count[lambda_k];
            const int64_t swap_node = index_to_node[new_pos];
            index_to_node[old_pos] = (bk_t)swap_node;
            index_to_node[new_pos] = (bk_t)rowk;
            nk.pos = (bk_t)new_pos;
            node[swap_node].pos = (bk_t)old_pos;
            nk.lambda += 1;
            interval_count[lambda_k] -= 1;
            interval_count[lambda_k + 1] += 1;
            interval_ptr[lambda_k + 1] = (bk_t)(new_pos - 1);
          }
        }
      }
    }
    for (idx_t j = Tp[i - 1]; j < Tp[i]; ++j) {
      const int64_t row = (int64_t)Tj[j] + 1;
      if (state[row] == U_NODE) {
        NodeRec& nr = node[row];
        if (nr.lambda == 0) continue;
        const int64_t lambda_j = nr.lambda + 1;
        const int64_t old_pos = nr.pos;
        const int64_t new_pos = interval_ptr[lambda_j] + 1;
        const int64_t swap_node = index_to_node[new_pos];
        index_to_node[old_pos] = (bk_t)swap_node;
        index_to_node[new_pos] = (bk_t)row;
        nr.pos = (bk_t)new_pos;
        node[swap_node].pos = (bk_t)old_pos;
        nr.lambda -= 1;
        interval_count[lambda_j] -= 1;
        interval_count[lambda_j - 1] += 1;
        interval_ptr[lambda_j] += 1;
      }
    }
  }
  for (int64_t i = 1; i <= n; ++i) splitting[i - 1] = (idx_t)state[i];
  return 0;
}

// ---------------------------------------------------------------------------------------
// Direct interpolation.  src/classical.jl:57-189.
// At: operator (column i == row i under the Hermitian assumption).  T: strength pattern
// (values ignored: classical.jl:58-60 overwrites them with At's values on T's pattern, and
// T's pattern is a subset of At's).  Output R (nc x n) in CSC: Rp[n+1], Rj, Rx sized by
// pass 1.  Two calls: first with Rj == nullptr fills Rp and returns nnz; second fills.
// *nc_out receives maximum(Rj)+1 (0 if empty).
// ---------------------------------------------------------------------------------------
int64_t amgsetup_direct_interpolation(int64_t n, const idx_t* Ap, const idx_t* Aj, const double* Ax,
                                      const idx_t* Tp, const idx_t* Tj, const idx_t* splitting,
                                      idx_t* Rp, idx_t* Rj, double* Rx, int64_t* nc_out) {
  enum { C_NODE = 1 };
  // pass 1  (classical.jl:71-89)
  {
    idx_t cnt = 0;
    Rp[0] = 0;
    for (int64_t i = 0; i < n; ++i) {
      if (splitting[i] == C_NODE) cnt += 1;
      else
        for (idx_t j = Tp[i]; j < Tp[i + 1]; ++j)
          if (splitting[Tj[j]] == C_NODE) cnt += 1;
      Rp[i + 1] = cnt;
    }
  }
  if (!Rj) return Rp[n];
  const double eps = std::numeric_limits<double>::epsilon();
  // value of At at (row, col i) for a T entry: T's pattern is a subset of At's, both sorted
  int bad_pattern = 0;   // set (never cleared) by any thread that finds a T entry outside At's pattern
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    bool bad_row = false;
    if (splitting[i] == C_NODE) {
      Rj[Rp[i]] = (idx_t)i;
      Rx[Rp[i]] = 1.0;
      continue;
    }
    double sum_strong_pos = 0, sum_strong_neg = 0;
    {
      idx_t a = Ap[i];
      for (idx_t j = Tp[i]; j < Tp[i + 1]; ++j) {
        const idx_t row = Tj[j];
        while (a < Ap[i + 1] && Aj[a] != row) ++a;
        if (a >= Ap[i + 1]) { bad_row = true; break; }  // T's pattern must be a subset of At's
        const double sval = Ax[a];
        if (splitting[row] == C_NODE) {
          if (sval < 0) sum_strong_neg += sval; else sum_strong_pos += sval;
        }
      }
    }
    if (bad_row) {
#pragma omp atomic write
      bad_pattern = 1;
      continue;
    }
    double sum_all_pos = 0, sum_all_neg = 0, diag = 0;
    for (idx_t j = Ap[i]; j < Ap[i + 1]; ++j) {
      const double aval = Ax[j];
      if (Aj[j] == i) diag += aval;
      else if (aval < 0) sum_all_neg += aval;
      else sum_all_pos += aval;
    }
    double alpha, beta;
    if (sum_strong_pos == 0) { beta = 0; if (diag >= 0) diag += sum_all_pos; }
    else beta = sum_all_pos / sum_strong_pos;
    if (sum_strong_neg == 0) { alpha = 0; if (diag < 0) diag += sum_all_neg; }
    else alpha = sum_all_neg / sum_strong_neg;
    double neg_coeff, pos_coeff;
    if (std::fabs(diag) <= eps) { neg_coeff = 0; pos_coeff = 0; }
    else { neg_coeff = alpha / diag; pos_coeff = beta / diag; }
    idx_t q = Rp[i];
    idx_t a = Ap[i];
    for (idx_t j = Tp[i]; j < Tp[i + 1]; ++j) {
      const idx_t row = Tj[j];
      while (Aj[a] != row) ++a;
      const double sval = Ax[a];
      if (splitting[row] == C_NODE) {
        Rj[q] = row;
        Rx[q] = (sval < 0) ? std::fabs(neg_coeff * sval) : std::fabs(pos_coeff * sval);
        ++q;
      }
    }
  }
  if (bad_pattern) return -3;
  // coarse numbering = exclusive prefix sum of splitting  (classical.jl:180-186)
  std::vector<idx_t> map(n);
  idx_t sum = 0;
  for (int64_t i = 0; i < n; ++i) { map[i] = sum; sum += splitting[i]; }
  idx_t mx = -1;
  const int64_t nnz = Rp[n];
#pragma omp parallel for schedule(static) reduction(max : mx)
  for (int64_t k = 0; k < nnz; ++k) { Rj[k] = map[Rj[k]]; mx = std::max(mx, Rj[k]); }
  *nc_out = (int64_t)mx + 1;
  return nnz;
}

// ---------------------------------------------------------------------------------------
// Sparse x sparse product C = A*B (CSC, Gustavson by columns).  Structural zeros produced
// by cancellation are KEPT (as Julia's SparseArrays spmatmul does; the reference's nnz
// goldens at test/runtests.jl:80-102 rely on that); rows sorted per column; per-entry
// accumulation order = ascending k over B's column, as in the stdlib loop.
// A: m x k, B: k x n.  Two-phase API: amgsetup_spgemm_begin computes into an internal
// buffer and returns nnz; amgsetup_spgemm_fetch copies out and frees.
// ---------------------------------------------------------------------------------------
// Work distribution: the columns of C are cut into chunks of equal ESTIMATED work (products per column), the chunks are
// handed out dynamically; a thread computes a chunk into its own scratch (kept between chunks and calls: no page faults, no
// growth in the loop), then copies it into an exactly-sized block.  amgsetup_spgemm_fetch copies the blocks to their final
// place on all cores.  The result does not depend on the team size or on which thread computed which chunk.
struct SpgemmChunk {
  int64_t c0 = 0, c1 = 0, nnz = 0;
  idx_t* rows = nullptr;
  double* vals = nullptr;
};
struct SpgemmResult {
  std::vector<SpgemmChunk> chunks;
  std::vector<idx_t> colcount;            // per column
  int64_t n = 0, nnz = 0;
  ~SpgemmResult() {
    for (auto& c : chunks) { std::free(c.rows); std::free(c.vals); }
  }
};
static SpgemmResult* g_spgemm = nullptr;
static thread_local std::vector<double> tl_spgemm_acc;
static thread_local std::vector<int32_t> tl_spgemm_mark;
static thread_local int32_t tl_spgemm_stamp = 0;
static thread_local std::vector<idx_t> tl_spgemm_rows;     // chunk scratch
static thread_local std::vector<double> tl_spgemm_vals;

// gives the per-thread accumulators back (call when a hierarchy is finished; the OpenMP pool threads persist)
void amgsetup_spgemm_release(void) {
#pragma omp parallel
  {
    std::vector<double>().swap(tl_spgemm_acc);
    std::vector<int32_t>().swap(tl_spgemm_mark);
    std::vector<idx_t>().swap(tl_spgemm_rows);
    std::vector<double>().swap(tl_spgemm_vals);
    tl_spgemm_stamp = 0;
  }
}

int64_t amgsetup_spgemm_begin(int64_t m, int64_t k, int64_t n, const idx_t* Ap, const idx_t* Aj,
                              const double* Ax, const idx_t* Bp, const idx_t* Bj, const double* Bx) {
  (void)k;
  delete g_spgemm;
  g_spgemm = new SpgemmResult();
  SpgemmResult& R = *g_spgemm;
  R.n = n;
  R.colcount.assign(n, 0);
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  // products per column (an upper bound of the column's entries), as a running sum
  std::vector<int64_t> work(n + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; ++j) {
    int64_t w = 0;
    for (idx_t bp = Bp[j]; bp < Bp[j + 1]; ++bp) w += Ap[Bj[bp] + 1] - Ap[Bj[bp]];
    work[j + 1] = w;
  }
  for (int64_t j = 0; j < n; ++j) work[j + 1] += work[j];
  const int64_t total = work[n];
  // ~32 chunks per thread, at least 256 products each
  const int64_t per_chunk = std::max<int64_t>(256, total / std::max<int64_t>(1, (int64_t)nthreads * 32) + 1);
  {
    int64_t c0 = 0;
    while (c0 < n) {
      const int64_t target = work[c0] + per_chunk;
      int64_t c1 = std::upper_bound(work.begin() + c0 + 1, work.begin() + n + 1, target) - work.begin();
      c1 = std::min<int64_t>(std::max<int64_t>(c1 - 1, c0 + 1), n);
      SpgemmChunk ch;
      ch.c0 = c0;
      ch.c1 = c1;
      R.chunks.push_back(ch);
      c0 = c1;
    }
  }
  const int64_t nchunks = (int64_t)R.chunks.size();
  int failed = 0;
  const bool verbose = std::getenv("B200AMG_VERBOSE_SETUP") != nullptr;   // stage timers on stderr
  const double t_begin = verbose ? wall_s() : 0.0;
  double tc_sum = 0, tm_sum = 0;
#pragma omp parallel num_threads(nthreads) reduction(+:tc_sum, tm_sum)
  {
    // dense accumulator + marker per thread, kept between calls (a setup multiplies 2 x levels times): the marker holds
    // a per-thread column stamp so nothing has to be cleared; only growth allocates
    std::vector<double>& acc = tl_spgemm_acc;
    std::vector<int32_t>& mark = tl_spgemm_mark;
    int32_t& stamp = tl_spgemm_stamp;
    if ((int64_t)acc.size() < m) { acc.resize(m); mark.assign(m, 0); stamp = 0; }
    std::vector<idx_t>& out_r = tl_spgemm_rows;
    std::vector<double>& out_v = tl_spgemm_vals;
#pragma omp for schedule(dynamic, 1)
    for (int64_t c = 0; c < nchunks; ++c) {
      SpgemmChunk& ch = R.chunks[c];
      const int64_t bound = work[ch.c1] - work[ch.c0];
      if ((int64_t)out_r.size() < bound) { out_r.resize(bound); out_v.resize(bound); }
      idx_t* orow = out_r.data();
      double* oval = out_v.data();
      int64_t cnt = 0;
      const double tq0 = verbose ? wall_s() : 0.0;
      for (int64_t j = ch.c0; j < ch.c1; ++j) {
        if (stamp == std::numeric_limits<int32_t>::max()) { std::fill(mark.begin(), mark.end(), 0); stamp = 0; }
        ++stamp;
        const int64_t first = cnt;
        for (idx_t bp = Bp[j]; bp < Bp[j + 1]; ++bp) {
          const idx_t kk = Bj[bp];
          const double bv = Bx[bp];
          for (idx_t ap = Ap[kk]; ap < Ap[kk + 1]; ++ap) {
            const idx_t r = Aj[ap];
            const double prod = Ax[ap] * bv;
            if (mark[r] != stamp) { mark[r] = stamp; acc[r] = prod; orow[cnt++] = r; }
            else acc[r] += prod;
          }
        }
        std::sort(orow + first, orow + cnt);
        for (int64_t q = first; q < cnt; ++q) oval[q] = acc[orow[q]];
        R.colcount[j] = (idx_t)(cnt - first);
      }
      ch.nnz = cnt;
      const double tq1 = verbose ? wall_s() : 0.0;
      tc_sum += tq1 - tq0;
      if (cnt > 0) {
        ch.rows = (idx_t*)std::malloc(sizeof(idx_t) * cnt);
        ch.vals = (double*)std::malloc(sizeof(double) * cnt);
        if (!ch.rows || !ch.vals) {
#pragma omp atomic write
          failed = 1;
        } else {
          std::memcpy(ch.rows, orow, sizeof(idx_t) * cnt);
          std::memcpy(ch.vals, oval, sizeof(double) * cnt);
        }
      }
      if (verbose) tm_sum += wall_s() - tq1;
    }
  }
  if (verbose) std::fprintf(stderr, "[amgsetup] spgemm %lld chunks: wall %.3f s, thread-seconds compute %.3f, block copy %.3f\n", (long long)nchunks, wall_s() - t_begin, tc_sum, tm_sum);
  if (failed) { delete g_spgemm; g_spgemm = nullptr; return -3; }
  int64_t nnz = 0;
  for (auto& c : R.chunks) nnz += c.nnz;
  R.nnz = nnz;
  if (nnz > std::numeric_limits<idx_t>::max()) { delete g_spgemm; g_spgemm = nullptr; return -2; }
  return nnz;
}

int amgsetup_spgemm_fetch(idx_t* Cp, idx_t* Cj, double* Cx) {
  if (!g_spgemm) return -1;
  SpgemmResult& R = *g_spgemm;
  Cp[0] = 0;
  for (int64_t j = 0; j < R.n; ++j) Cp[j + 1] = Cp[j] + R.colcount[j];
  const int64_t nchunks = (int64_t)R.chunks.size();
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t c = 0; c < nchunks; ++c) {
    SpgemmChunk& ch = R.chunks[c];
    if (ch.nnz > 0) {
      const int64_t off = Cp[ch.c0];
      std::memcpy(Cj + off, ch.rows, sizeof(idx_t) * ch.nnz);
      std::memcpy(Cx + off, ch.vals, sizeof(double) * ch.nnz);
    }
    std::free(ch.rows);
    std::free(ch.vals);
    ch.rows = nullptr;
    ch.vals = nullptr;
  }
  delete g_spgemm;
  g_spgemm = nullptr;
  return 0;
}

// ---------------------------------------------------------------------------------------
// Standard (Vanek) aggregation.  src/aggregate.jl:12-134.
// In: S (n x n, values used for "strongest neighbour" in pass 2).
// Out: x[i] = aggregate id (0-based) or -1 for isolated nodes.  Returns number of aggregates.
// ---------------------------------------------------------------------------------------
int64_t amgsetup_standard_aggregation(int64_t n, const idx_t* Sp, const idx_t* Sj, const double* Sx,
                                      int64_t* x) {
  std::fill(x, x + n, (int64_t)0);
  int64_t next_aggregate = 1;
  // Pass 1  (:19-51)
  for (int64_t i = 0; i < n; ++i) {
    if (x[i] != 0) continue;
    bool has_agg_neighbors = false, has_neighbors = false;
    for (idx_t j = Sp[i]; j < Sp[i + 1]; ++j) {
      const idx_t row = Sj[j];
      if (row != i) {
        has_neighbors = true;
        if (x[row] != 0) { has_agg_neighbors = true; break; }
      }
    }
    if (!has_neighbors) x[i] = -n;
    else if (!has_agg_neighbors) {
      x[i] = next_aggregate;
      for (idx_t j = Sp[i]; j < Sp[i + 1]; ++j)
        if (Sj[j] != i) x[Sj[j]] = next_aggregate;
      next_aggregate += 1;
    }
  }
  // Pass 2  (:54-74)
  for (int64_t i = 0; i < n; ++i) {
    if (x[i] != 0) continue;
    double s_best = 0.0;
    int64_t x_best = 0;
    for (idx_t j = Sp[i]; j < Sp[i + 1]; ++j) {
      const int64_t x_row = x[Sj[j]];
      const double s_candidate = Sx[j];
      if (x_row > 0 && s_candidate > s_best) { s_best = s_candidate; x_best = x_row; }
    }
    if (x_best > 0) x[i] = -x_best;
  }
  std::vector<char> unagg(n);
  for (int64_t i = 0; i < n; ++i) unagg[i] = (x[i] == 0);
  // shift to 0-based  (:80-94).  NB: when n == 1 an isolated node has x = -n = -1 which the
  // reference's `xi == -n` test catches before the generic negative branch; keep that order.
  next_aggregate -= 1;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t xi = x[i];
    if (xi > 0) x[i] = xi - 1;
    else if (xi == -n) x[i] = -1;
    else if (xi < 0) x[i] = -xi - 1;
  }
  // Pass 3  (:99-113)
  for (int64_t i = 0; i < n; ++i) {
    if (!unagg[i]) continue;
    x[i] = next_aggregate;
    for (idx_t j = Sp[i]; j < Sp[i + 1]; ++j) {
      const idx_t row = Sj[j];
      if (unagg[row]) { x[row] = next_aggregate; unagg[row] = 0; }
    }
    unagg[i] = 0;
    next_aggregate += 1;
  }
  return next_aggregate;
}

// ---------------------------------------------------------------------------------------
// fit_candidates, vector B.  src/aggregation.jl:161-193 (+ norm_col :232-240).
// A = copy(AggOp') (n_fine x n_coarse CSC, one column per aggregate) given by pattern; fills
// Tx (the tentative prolongator's values) and Rc (coarse candidate = column norms).
// ---------------------------------------------------------------------------------------
int amgsetup_fit_candidates_vec(int64_t n_coarse, const idx_t* Ap, const idx_t* Aj, const double* B,
                                double tol, double* Tx, double* Rc) {
  for (int64_t i = 0; i < n_coarse; ++i) {
    double s = 0.0;
    for (idx_t j = Ap[i]; j < Ap[i + 1]; ++j) { const double v = B[Aj[j]]; s += v * v; }
    const double norm_i = std::sqrt(s);
    const double threshold_i = tol * norm_i;
    double scale;
    if (norm_i > threshold_i) { scale = 1 / norm_i; Rc[i] = norm_i; }
    else { scale = 0; Rc[i] = 0; }
    for (idx_t j = Ap[i]; j < Ap[i + 1]; ++j) Tx[j] = B[Aj[j]] * scale;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// LocalWeighting of the Jacobi prolongation smoother.  src/aggregation.jl:26-59.
// Out: W = omega * D^-1 * A with D = row sums of |A| (same pattern as A).
// ---------------------------------------------------------------------------------------
int amgsetup_local_weight(int64_t n, const idx_t* Ap, const idx_t* Aj, const double* Ax, double omega,
                          double* Wx) {
  std::vector<double> D(n, 0.0);
  for (int64_t i = 0; i < n; ++i)
    for (idx_t j = Ap[i]; j < Ap[i + 1]; ++j) D[Aj[j]] += std::fabs(Ax[j]);
  for (int64_t i = 0; i < n; ++i)
    if (D[i] != 0) D[i] = 1 / D[i];
  const int64_t nnz = Ap[n];
  for (int64_t i = 0; i < n; ++i)
    for (idx_t j = Ap[i]; j < Ap[i + 1]; ++j) Wx[j] = Ax[j] * D[Aj[j]];
  for (int64_t k = 0; k < nnz; ++k) Wx[k] *= omega;
  return 0;
}

// C = A - B for same-shape CSC matrices (union pattern, exact zeros not stored — the
// stdlib's zero-preserving map drops them).  Returns nnz; Cj/Cx sized nnz(A)+nnz(B).
int64_t amgsetup_sub(int64_t n, const idx_t* Ap, const idx_t* Aj, const double* Ax, const idx_t* Bp,
                     const idx_t* Bj, const double* Bx, idx_t* Cp, idx_t* Cj, double* Cx) {
  idx_t q = 0;
  Cp[0] = 0;
  for (int64_t j = 0; j < n; ++j) {
    idx_t a = Ap[j], b = Bp[j];
    const idx_t ae = Ap[j + 1], be = Bp[j + 1];
    while (a < ae || b < be) {
      idx_t r;
      double v;
      if (b >= be || (a < ae && Aj[a] < Bj[b])) { r = Aj[a]; v = Ax[a]; ++a; }
      else if (a >= ae || Bj[b] < Aj[a]) { r = Bj[b]; v = -Bx[b]; ++b; }
      else { r = Aj[a]; v = Ax[a] - Bx[b]; ++a; ++b; }
      if (v != 0.0) { Cj[q] = r; Cx[q] = v; ++q; }
    }
    Cp[j + 1] = q;
  }
  return q;
}

// ---------------------------------------------------------------------------------------
// Setup-time relaxation used by smoothed aggregation's improve_candidates
// (src/aggregation.jl:135-136 -> src/smoother.jl:34-38,61-90): symmetric Gauss-Seidel on
// A*B = 0, column-as-row.  Setup-only helper: the solve phase never calls this (its
// sweeps are CUDA kernels behind include/b200amg.h).
// ---------------------------------------------------------------------------------------
int amgsetup_gs_sweeps(int64_t n, const idx_t* Ap, const idx_t* Aj, const double* Ax, const double* b,
                       double* x, int64_t ncols, int iters, int forward, int backward) {
  for (int it = 0; it < iters; ++it) {
    for (int dir = 0; dir < 2; ++dir) {
      if (dir == 0 && !forward) continue;
      if (dir == 1 && !backward) continue;
      for (int64_t c = 0; c < ncols; ++c) {
        double* xc = x + c * n;
        const double* bc = b + c * n;
        for (int64_t s = 0; s < n; ++s) {
          const int64_t i = dir == 0 ? s : n - 1 - s;
          double rsum = 0, d = 0;
          for (idx_t j = Ap[i]; j < Ap[i + 1]; ++j) {
            const idx_t row = Aj[j];
            const double val = Ax[j];
            if (row == i) d = val; else rsum += val * xc[row];
          }
          if (d != 0) xc[i] = (bc[i] - rsum) / d;
        }
      }
    }
  }
  return 0;
}

// y = A*x  (CSC scatter) — setup-side helper for building right-hand sides b = A*ones.
int amgsetup_csc_matvec(int64_t m, int64_t n, const idx_t* Ap, const idx_t* Aj, const double* Ax,
                        const double* x, double* y) {
  std::fill(y, y + m, 0.0);
  for (int64_t j = 0; j < n; ++j) {
    const double xj = x[j];
    for (idx_t k = Ap[j]; k < Ap[j + 1]; ++k) y[Aj[k]] += Ax[k] * xj;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// r = b - A x on ALL host cores, walking the columns of a (numerically symmetric) CSC matrix as rows.  NOT reference
// behaviour (the reference's solve phase is single-threaded): bench.py reports it as a courtesy upper bound of what the
// host's memory system can do on the headline kernel.  Returns the best wall-clock seconds of `reps` passes.
// ---------------------------------------------------------------------------------------
double amgsetup_residual_allcores(int64_t n, const idx_t* ptr, const idx_t* idx, const double* val, const double* x,
                                  const double* b, double* r, int reps) {
  double best = 1e300;
  for (int rep = 0; rep < (reps > 0 ? reps : 1); ++rep) {
    const double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      double s = 0.0;
      for (idx_t k = ptr[i]; k < ptr[i + 1]; ++k) s += val[k] * x[idx[k]];
      r[i] = b[i] - s;
    }
    best = std::min(best, omp_get_wtime() - t0);
  }
  return best;
}

}  // extern "C"
