// Row partition of the FINE level across ranks (one process per GPU) — host-side plan.
//
// The reference has no distributed path; this is the B200 design for BASELINE config 4: the finest
// level's A, P, R and vectors are split by contiguous 1-D row blocks (balanced by nnz), coarser
// levels stay on rank 0.  Every rank holds the full host hierarchy (setup is deterministic), so each
// rank derives every list below locally — no setup-time communication.
//
//   rows owned by rank g          [row_split[g], row_split[g+1])
//   coarse rows restricted by g   [coarse_split[g], coarse_split[g+1])   (rows of R whose first fine column g owns)
//   halo                          sorted global columns outside the block that the owned rows of A, A' (the
//                                 walked matrix of the "fast" smoothers) and R reference; segment per owner
//   send list for rank q          the owned entries q's halo needs (what pack gathers before the exchange)
//   coarse_x window               [cx_lo, cx_hi): the coarse entries the owned rows of P reference
//
// Local vectors are laid out [owned | halo]; local column ids are remapped accordingly, the order of
// the entries inside a row is unchanged (so accumulation order == single-GPU order == reference).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace b200amg {

struct PartCsr {   // the same staging type the engine uses (int32, 0-based, compressed by rows)
  int64_t nrows = 0, ncols = 0;
  std::vector<int> ptr, idx;
  std::vector<double> val;
};

struct PartPlan {
  int rank = 0, world = 1;
  std::vector<int64_t> row_split, coarse_split;
  int64_t nloc = 0, nhalo = 0, ncloc = 0;
  std::vector<int> halo_cols;          // nhalo global column ids, ascending
  std::vector<int> recv_off;           // world + 1: halo segment of each owner rank
  std::vector<int> send_idx;           // local owned indices, grouped by destination rank
  std::vector<int> send_off;           // world + 1
  int64_t cx_lo = 0, cx_hi = 0;        // my coarse_x window
  std::vector<int64_t> cx_lo_all, cx_hi_all;   // every rank's window (rank 0 sends them)
};

inline int part_owner(const std::vector<int64_t>& split, int64_t i) {
  return (int)(std::upper_bound(split.begin(), split.end(), i) - split.begin()) - 1;
}

// columns outside [lo, hi) referenced by rows [r0, r1) of m
inline void part_collect_external(const int* ptr, const int* idx, int64_t r0, int64_t r1, int64_t lo, int64_t hi,
                                  std::vector<int>& out) {
  for (int64_t k = ptr[r0]; k < ptr[r1]; ++k) {
    const int c = idx[k];
    if (c < lo || c >= hi) out.push_back(c);
  }
}
inline void part_sort_unique(std::vector<int>& v) {
  std::sort(v.begin(), v.end());
  v.erase(std::unique(v.begin(), v.end()), v.end());
}

// A, At: n x n by rows (At may alias A); R: nc x n by rows; P: n x nc by rows.
// given_split: row blocks to use (a level below a partitioned one inherits its parent's coarse_split) or nullptr.
// parentP / parent_split: the prolongation of the level above and its row blocks — the entries of THIS level's x that
// the parent's owned rows of P reference must be in the halo too (nullptr for the finest level).
template <class Csr>
PartPlan make_part_plan(int rank, int world, const Csr& A, const Csr* At, const Csr& R, const Csr& P,
                        const std::vector<int64_t>* given_split = nullptr, const Csr* parentP = nullptr,
                        const std::vector<int64_t>* parent_split = nullptr) {
  PartPlan pl;
  pl.rank = rank;
  pl.world = world;
  const int64_t n = A.nrows, nc = R.nrows;
  // ---- row blocks: inherited, or balanced by nnz(A) ----
  pl.row_split.assign(world + 1, n);
  pl.row_split[0] = 0;
  if (given_split) {
    pl.row_split = *given_split;
  } else {
    const double total = (double)A.ptr[n];
    int64_t r = 0;
    for (int g = 1; g < world; ++g) {
      const double target = total * g / world;
      while (r < n && (double)A.ptr[r] < target) ++r;
      pl.row_split[g] = r;
    }
  }
  // ---- coarse rows: owner of the first (smallest) fine column of the R row; made monotone ----
  pl.coarse_split.assign(world + 1, nc);
  pl.coarse_split[0] = 0;
  {
    int64_t i = 0;
    for (int g = 1; g < world; ++g) {
      while (i < nc && (R.ptr[i + 1] == R.ptr[i] || R.idx[R.ptr[i]] < pl.row_split[g])) ++i;
      pl.coarse_split[g] = i;
    }
  }
  // ---- halo lists of every rank (needed to know what to send) ----
  std::vector<std::vector<int>> halos(world);
  for (int g = 0; g < world; ++g) {
    const int64_t lo = pl.row_split[g], hi = pl.row_split[g + 1];
    std::vector<int>& hcols = halos[g];
    part_collect_external(A.ptr.data(), A.idx.data(), lo, hi, lo, hi, hcols);
    if (At && At != &A) part_collect_external(At->ptr.data(), At->idx.data(), lo, hi, lo, hi, hcols);
    part_collect_external(R.ptr.data(), R.idx.data(), pl.coarse_split[g], pl.coarse_split[g + 1], lo, hi, hcols);
    if (parentP && parent_split)
      part_collect_external(parentP->ptr.data(), parentP->idx.data(), (*parent_split)[g], (*parent_split)[g + 1], lo, hi, hcols);
    part_sort_unique(hcols);
  }
  const int64_t lo = pl.row_split[rank], hi = pl.row_split[rank + 1];
  pl.nloc = hi - lo;
  pl.ncloc = pl.coarse_split[rank + 1] - pl.coarse_split[rank];
  pl.halo_cols = halos[rank];
  pl.nhalo = (int64_t)pl.halo_cols.size();
  pl.recv_off.assign(world + 1, 0);
  for (int c : pl.halo_cols) pl.recv_off[part_owner(pl.row_split, c) + 1]++;
  for (int g = 0; g < world; ++g) pl.recv_off[g + 1] += pl.recv_off[g];
  pl.send_off.assign(world + 1, 0);
  for (int q = 0; q < world; ++q) {
    if (q != rank)
      for (int c : halos[q])
        if (c >= lo && c < hi) pl.send_idx.push_back((int)(c - lo));
    pl.send_off[q + 1] = (int)pl.send_idx.size();
  }
  // ---- coarse_x windows ----
  pl.cx_lo_all.assign(world, 0);
  pl.cx_hi_all.assign(world, 0);
  for (int g = 0; g < world; ++g) {
    int cmin = INT32_MAX, cmax = -1;
    for (int64_t k = P.ptr[pl.row_split[g]]; k < P.ptr[pl.row_split[g + 1]]; ++k) {
      cmin = std::min(cmin, P.idx[k]);
      cmax = std::max(cmax, P.idx[k]);
    }
    if (cmax < 0) cmin = 0;
    pl.cx_lo_all[g] = cmin;
    pl.cx_hi_all[g] = cmax + 1;
  }
  pl.cx_lo = pl.cx_lo_all[rank];
  pl.cx_hi = pl.cx_hi_all[rank];
  return pl;
}

// rows [r0, r1) of m with columns remapped: owned [lo, hi) -> c - lo, others -> nloc + position in halo_cols
template <class Csr>
Csr part_local_block(const Csr& m, int64_t r0, int64_t r1, int64_t lo, int64_t hi, const std::vector<int>& halo_cols) {
  Csr out;
  out.nrows = r1 - r0;
  out.ncols = (hi - lo) + (int64_t)halo_cols.size();
  out.ptr.resize(out.nrows + 1);
  const int base = m.ptr[r0];
  for (int64_t r = r0; r <= r1; ++r) out.ptr[r - r0] = m.ptr[r] - base;
  const int64_t nnz = m.ptr[r1] - base;
  out.idx.resize(nnz);
  out.val.assign(m.val.begin() + base, m.val.begin() + base + nnz);
  const int nloc = (int)(hi - lo);
  for (int64_t k = 0; k < nnz; ++k) {
    const int c = m.idx[base + k];
    if (c >= lo && c < hi) out.idx[k] = (int)(c - lo);
    else out.idx[k] = nloc + (int)(std::lower_bound(halo_cols.begin(), halo_cols.end(), c) - halo_cols.begin());
  }
  return out;
}

// rows [r0, r1) of m with columns shifted by -shift (the coarse_x window of P)
template <class Csr>
Csr part_shifted_block(const Csr& m, int64_t r0, int64_t r1, int64_t shift, int64_t ncols) {
  Csr out;
  out.nrows = r1 - r0;
  out.ncols = ncols;
  out.ptr.resize(out.nrows + 1);
  const int base = m.ptr[r0];
  for (int64_t r = r0; r <= r1; ++r) out.ptr[r - r0] = m.ptr[r] - base;
  const int64_t nnz = m.ptr[r1] - base;
  out.idx.resize(nnz);
  out.val.assign(m.val.begin() + base, m.val.begin() + base + nnz);
  for (int64_t k = 0; k < nnz; ++k) out.idx[k] = (int)(m.idx[base + k] - shift);
  return out;
}

}  // namespace b200amg
