// Host-side sparse staging shared by the engine translation units: int32, 0-based, "by rows" (compressed along the
// first index), plus the renumbering helpers.  Pure C++ (OpenMP), no CUDA.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <utility>
#include <vector>

namespace b200amg {

// ------------------------------------------------------------------------------------------
// host-side sparse staging (int32, 0-based, "by rows" = compressed along the first index)
// ------------------------------------------------------------------------------------------
// The staging arrays of a 256^3 hierarchy are 0.5-1.3 GB each and every one of them is completely overwritten by a parallel
// loop right after it is sized: std::vector's value-initialisation would first zero (and page-fault) them on ONE thread.
// This allocator default-initialises instead, so the first touch happens in the parallel fill.
template <class T>
struct NoInitAlloc : std::allocator<T> {
  template <class U> struct rebind { using other = NoInitAlloc<U>; };
  NoInitAlloc() = default;
  template <class U> NoInitAlloc(const NoInitAlloc<U>&) {}
  template <class U, class... Args>
  void construct(U* p, Args&&... args) {
    if constexpr (sizeof...(Args) == 0) ::new ((void*)p) U;
    else ::new ((void*)p) U(std::forward<Args>(args)...);
  }
};
template <class T> using hvec = std::vector<T, NoInitAlloc<T>>;

struct HostCsr {
  int64_t nrows = 0, ncols = 0;
  hvec<int> ptr, idx;      // (resize() leaves new elements uninitialised: every producer overwrites them all)
  hvec<double> val;
  int64_t nnz() const { return ptr.empty() ? 0 : ptr.back(); }
};

// Entries land in their output row by an atomic cursor (any order), then every output row is sorted by index: the
// result is the sequential counting-sort transpose (ascending input row inside an output row), built on all cores.
static inline HostCsr transpose(const HostCsr& a) {
  HostCsr t;
  t.nrows = a.ncols;
  t.ncols = a.nrows;
  const int64_t nnz = a.nnz();
  t.ptr.assign(t.nrows + 1, 0);
  t.idx.resize(nnz);
  t.val.resize(nnz);
  int* cnt = t.ptr.data() + 1;
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nnz; ++k) __atomic_fetch_add(&cnt[a.idx[k]], 1, __ATOMIC_RELAXED);
  for (int64_t i = 0; i < t.nrows; ++i) t.ptr[i + 1] += t.ptr[i];
  std::vector<int> next(t.ptr.begin(), t.ptr.end() - 1);
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < a.nrows; ++r)
    for (int k = a.ptr[r]; k < a.ptr[r + 1]; ++k) {
      const int q = __atomic_fetch_add(&next[a.idx[k]], 1, __ATOMIC_RELAXED);
      t.idx[q] = (int)r;
      t.val[q] = a.val[k];
    }
#pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t i = 0; i < t.nrows; ++i) {   // insertion sort: rows are short and nearly sorted
    const int b = t.ptr[i], e = t.ptr[i + 1];
    for (int k = b + 1; k < e; ++k) {
      const int ci = t.idx[k];
      const double cv = t.val[k];
      int j = k - 1;
      while (j >= b && t.idx[j] > ci) { t.idx[j + 1] = t.idx[j]; t.val[j + 1] = t.val[j]; --j; }
      t.idx[j + 1] = ci;
      t.val[j + 1] = cv;
    }
  }
  return t;
}

static inline bool bit_equal(const HostCsr& a, const HostCsr& b) {
  return a.nrows == b.nrows && a.ncols == b.ncols && a.ptr == b.ptr && a.idx == b.idx &&
         std::memcmp(a.val.data(), b.val.data(), sizeof(double) * a.val.size()) == 0;
}

// 2: a equals its transpose bit for bit, 1: only the pattern is symmetric, 0: neither.  Every entry (i, j) looks its
// mirror (j, i) up by binary search in row j (columns are sorted), rows in parallel: no transpose is materialised.
static inline int symmetry_kind(const HostCsr& a) {
  if (a.nrows != a.ncols) return 0;
  int kind = 2;
#pragma omp parallel for schedule(dynamic, 4096) reduction(min : kind)
  for (int64_t i = 0; i < a.nrows; ++i) {
    if (kind == 0) continue;
    for (int k = a.ptr[i]; k < a.ptr[i + 1]; ++k) {
      const int j = a.idx[k];
      const int* lo = a.idx.data() + a.ptr[j];
      const int* hi = a.idx.data() + a.ptr[j + 1];
      const int* it = std::lower_bound(lo, hi, (int)i);
      if (it == hi || *it != (int)i) { kind = 0; break; }
      if (std::memcmp(&a.val[it - a.idx.data()], &a.val[k], sizeof(double)) != 0) kind = std::min(kind, 1);
    }
  }
  return kind;
}

// Wavefront (level) number of every row for an in-order FORWARD sweep over rows 0..n-1 of `a`, honouring
// both true dependencies (a_ij, j earlier) and anti-dependencies (a_ji): the dependency graph is the
// symmetrised pattern, which is why `at` (the transpose pattern) is needed.  The backward sweep walks
// the same wavefronts in reverse order (level strictly increases along every edge, so the reversed
// numbering is a valid schedule for the descending-index sweep).
static inline std::vector<int> wavefront_levels(const HostCsr& a, const HostCsr& at, int* nlev_out) {
  const int64_t n = a.nrows;
  std::vector<int> level(n, 0);
  int nlev = 0;
  for (int64_t i = 0; i < n; ++i) {
    int lv = 0;
    for (int k = a.ptr[i]; k < a.ptr[i + 1]; ++k) {
      const int j = a.idx[k];
      if (j < i) lv = std::max(lv, level[j] + 1);
    }
    if (&at != &a)
      for (int k = at.ptr[i]; k < at.ptr[i + 1]; ++k) {
        const int j = at.idx[k];
        if (j < i) lv = std::max(lv, level[j] + 1);
      }
    level[i] = lv;
    nlev = std::max(nlev, lv + 1);
  }
  *nlev_out = nlev;
  return level;
}

// MULTICOLOUR ordering (NOT the reference's: an explicit opt-in, B200AMG_GS_MULTICOLOR=1): greedy colouring of the symmetrised
// pattern in index order (smallest colour no neighbour holds).  Used in place of the wavefront number, it makes the sweep
// relax colour after colour — a Gauss-Seidel sweep in a DIFFERENT row order (a handful of wide, independent "wavefronts":
// bandwidth-bound) whose iterates differ from gs! (src/smoother.jl:73-90) while the fixed point is the same.
static inline std::vector<int> greedy_colours(const HostCsr& a, const HostCsr& at, int* ncol_out) {
  const int64_t n = a.nrows;
  std::vector<int> colour((size_t)n, -1);
  std::vector<int64_t> seen;   // seen[c] == i: colour c is taken by a neighbour of row i
  int ncol = 0;
  for (int64_t i = 0; i < n; ++i) {
    auto mark = [&](const HostCsr& m) {
      for (int k = m.ptr[i]; k < m.ptr[i + 1]; ++k) {
        const int j = m.idx[k];
        if (j != i && colour[(size_t)j] >= 0) seen[(size_t)colour[(size_t)j]] = i;
      }
    };
    mark(a);
    if (&at != &a) mark(at);
    int c = 0;
    while (c < ncol && seen[(size_t)c] == i) ++c;
    if (c == ncol) { ++ncol; seen.push_back(-1); }
    colour[(size_t)i] = c;
  }
  *ncol_out = ncol;
  return colour;
}

// A renumbering of one level: new index p holds old row old_of_new[p]; empty vectors = identity.
struct HostPerm {
  std::vector<int> new_of_old, old_of_new;
  bool identity() const { return new_of_old.empty(); }
};
// rows AND columns renumbered; the order of the entries inside a row is kept (reference accumulation order)
static inline HostCsr permute_sym(const HostCsr& m, const HostPerm& p) {
  HostCsr out;
  out.nrows = m.nrows; out.ncols = m.ncols;
  out.ptr.resize(m.nrows + 1);
  out.idx.resize(m.idx.size());
  out.val.resize(m.val.size());
  out.ptr[0] = 0;
  for (int64_t q = 0; q < m.nrows; ++q) {
    const int r = p.old_of_new[q];
    out.ptr[q + 1] = out.ptr[q] + (m.ptr[r + 1] - m.ptr[r]);
  }
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < m.nrows; ++q) {
    const int r = p.old_of_new[q];
    int o = out.ptr[q];
    for (int k = m.ptr[r]; k < m.ptr[r + 1]; ++k, ++o) {
      out.idx[o] = p.new_of_old[m.idx[k]];
      out.val[o] = m.val[k];
    }
  }
  return out;
}
static inline HostCsr permute_rows(const HostCsr& m, const HostPerm& p) {
  if (p.identity()) return m;
  HostCsr out;
  out.nrows = m.nrows; out.ncols = m.ncols;
  out.ptr.resize(m.nrows + 1);
  out.idx.resize(m.idx.size());
  out.val.resize(m.val.size());
  out.ptr[0] = 0;
  for (int64_t q = 0; q < m.nrows; ++q) {
    const int r = p.old_of_new[q];
    out.ptr[q + 1] = out.ptr[q] + (m.ptr[r + 1] - m.ptr[r]);
  }
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < m.nrows; ++q) {
    const int r = p.old_of_new[q];
    std::copy(m.idx.begin() + m.ptr[r], m.idx.begin() + m.ptr[r + 1], out.idx.begin() + out.ptr[q]);
    std::copy(m.val.begin() + m.ptr[r], m.val.begin() + m.ptr[r + 1], out.val.begin() + out.ptr[q]);
  }
  return out;
}
static inline void map_cols(HostCsr& m, const HostPerm& p) {
  if (p.identity()) return;
  const int64_t nnz = (int64_t)m.idx.size();
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nnz; ++k) m.idx[k] = p.new_of_old[m.idx[k]];
}

}  // namespace b200amg
