// Plan of the BLOCKED exact-order Gauss-Seidel / SOR sweep (block_gs.cuh) — host side, pure C++.
//
// The reference's sweep (gs! src/smoother.jl:73-90, sor_step! :205-221) relaxes rows in index order; row i needs the
// NEW value of every neighbour j < i and the OLD value of every neighbour j > i.  The dependency DAG of a 3-D problem
// is deep (766 wavefronts for the 7-point 256^3 level, 1700 on coarse RS levels), and a hand-off between SMs costs
// 1-3 us through L2 — so round 1's sweeps (one hand-off per wavefront) ran at 7 % of the HBM roofline.
//
// Here the rows are grouped into TILES, one CTA relaxes a whole tile, and the tile is walked in the order of its LOCAL
// level schedule (dependencies inside the tile only): consecutive local steps hand over through shared memory and a
// CTA barrier (~0.2 us).  Only edges that cross tiles go through L2, they are waited for per STAGE (a few steps) by
// dedicated warps that run ahead of the arithmetic, and the tiles are shaped so that the producer of such an edge is
// many local steps ahead of its consumer:
//
//   * Tiles from two MONOTONE coordinates.  With theta = the gap in the histogram of index distances i - j that separates
//     the "next plane" neighbours of a lexicographically numbered 3-D problem from the in-plane ones,
//         K(i) = max_j<i ( K(j) + [i - j >  theta] )            "plane"
//         J(i) = max_j<i ( J(j) + (i - j <= theta ? i - j : 0) ) "offset inside the plane" (skews by itself when a
//                                                                 stencil couples (j+1, k-1), as RS coarse levels do)
//     Both never decrease along a dependency, so blocks  tile = (K / b, J / a)  numbered lexicographically form an
//     ACYCLIC tile graph whatever the matrix is: a tile only ever waits for lower-numbered tiles.  Matrices without such
//     a gap (2-D problems, unstructured meshes, small coarse levels) get contiguous index ranges (also acyclic).
//   * a, b are sized so that a local step holds ~X rows, X = what one CTA can stream in one step time.
//   * Rows are renumbered (tile, local step, old index); the entries inside a row keep the reference's order.
//
// Everything here is deterministic and depends on the matrix pattern only.
#pragma once
#include <algorithm>
#include <cmath>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "host_csr.h"

namespace b200amg {

// sizes of gs_block_kernel's shared-memory stages (block_gs.cuh) that the plan is built for
constexpr int kBgStageNnz = 1024;
constexpr int kBgStageRows = 256;
constexpr int kBgWindow = 2048;
constexpr int kBgDepth = 5;


struct BI4 { int x, y, z, w; };   // layout of CUDA's int4 / int2 (uploaded as such)
struct BI2 { int x, y; };

struct BlockPlanParams {
  int stage_nnz = 1024;     // entries one pipeline stage holds
  int stage_rows = 256;     // rows one pipeline stage holds (>= the widest step)
  int window = 4096;        // x values of the tile kept in shared memory (power of two)
  int depth = 6;            // pipeline stages in flight
  int max_lanes = 32;
  double step_us = 0.22;    // one local step (barrier + dependent arithmetic)
  double cta_gbs = 55.0;    // what one CTA streams while every SM is busy
  int cap_step_to_stage = 1;
  int force_tile_rows = 0;  // > 0: contiguous tiles of this many rows (diagnostics)
  int force_a = 0, force_b = 0;   // > 0: block sizes of the monotone coordinates
  int verbose = 0;
};

struct BlockPlan {
  bool ok = false;
  std::string why;              // why not, when !ok
  int64_t n = 0, nnz = 0;
  int ntiles = 0, nstages = 0, nsteps = 0, lanes = 1;
  int window = 0, window_eff = 0, depth = 0, stage_nnz = 0, stage_rows = 0;
  HostPerm perm;                // new index p holds old row old_of_new[p]
  std::vector<BI4> tile;        // {first stage, end stage, first row, end row}           (new numbering)
  std::vector<BI4> stage_meta;  // {first row, end row, first nnz, end nnz}
  std::vector<BI4> stage_aux;   // {first step (index into steps), steps, first fwd requirement, fwd requirements}
  std::vector<BI2> stage_auxb;  // {first bwd requirement, bwd requirements}
  std::vector<int> steps;       // end row (exclusive) of every step, stages concatenated
  std::vector<BI2> req_fwd;     // {tile, stages of it that must be complete (counted in forward order)}
  std::vector<BI2> req_bwd;     // {tile, stages of it that must be complete (counted from its LAST stage)}
  std::vector<int> order_fwd;   // ticket -> tile: a topological order of the tile graph that hands tiles out by the wavefront
  std::vector<int> order_bwd;   // they start with (there are more tiles than SMs: the resident ones must be the EARLIEST ones)
  // statistics
  double theta = 0, mean_step_rows = 0, target_step_rows = 0;
  int k_extent = 0, j_extent = 0, block_a = 0, block_b = 0, max_tile_steps = 0, global_wavefronts = 0;
  int64_t max_tile_rows = 0;
};

namespace blockplan_detail {

// the highest run of >= 3 empty half-octave bins of the distance histogram that has populated bins on both sides
static inline double valley_threshold(const HostCsr& w) {
  const int64_t n = w.nrows;
  std::vector<int64_t> h(64, 0);
  int64_t total = 0;
#pragma omp parallel
  {
    std::vector<int64_t> hl(64, 0);
#pragma omp for schedule(static) nowait
    for (int64_t i = 0; i < n; ++i)
      for (int k = w.ptr[i]; k < w.ptr[i + 1]; ++k) {
        const int64_t d = i - w.idx[k];
        if (d > 0) hl[(int)std::floor(2.0 * std::log2((double)d))]++;
      }
#pragma omp critical
    for (int q = 0; q < 64; ++q) h[q] += hl[q];
  }
  for (int q = 0; q < 64; ++q) total += h[q];
  if (total == 0) return 0.0;
  auto sig = [&](int q) { return (double)h[q] > 0.002 * (double)total; };
  int top = 63;
  while (top >= 0 && !sig(top)) --top;
  int b = top;
  while (b > 0) {
    if (!sig(b)) {
      const int e = b;
      while (b >= 0 && !sig(b)) --b;
      if (b >= 0 && e - b >= 3) return std::pow(2.0, (double)(e + 1 + b + 1) / 4.0);   // geometric middle of the gap
    } else {
      --b;
    }
  }
  return 0.0;
}

static inline int global_wavefronts(const HostCsr& w, std::vector<int>& lv) {
  lv.assign((size_t)w.nrows, 0);
  int nlev = 0;
  for (int64_t i = 0; i < w.nrows; ++i) {
    int l = 0;
    for (int k = w.ptr[i]; k < w.ptr[i + 1]; ++k) {
      const int j = w.idx[k];
      if (j < i) l = std::max(l, lv[j] + 1);
    }
    lv[i] = l;
    nlev = std::max(nlev, l + 1);
  }
  return nlev;
}

}  // namespace blockplan_detail

// w: the matrix the sweep walks, ORIGINAL numbering, structurally symmetric pattern, sorted columns.
static inline BlockPlan build_block_plan(const HostCsr& w, const BlockPlanParams& prm) {
  using namespace blockplan_detail;
  BlockPlan P;
  const int64_t n = w.nrows;
  P.n = n;
  P.nnz = w.nnz();
  P.window = prm.window;
  P.stage_nnz = prm.stage_nnz;
  P.stage_rows = prm.stage_rows;
  P.depth = prm.depth;
  P.window_eff = prm.window - prm.stage_rows;
  if (n == 0) { P.why = "empty level"; return P; }
  if ((int64_t)prm.depth * prm.stage_rows > P.window_eff) { P.why = "window too small for the pipeline depth"; return P; }
  const double mean = (double)P.nnz / (double)n;
  int maxlen = 0;
  for (int64_t i = 0; i < n; ++i) maxlen = std::max(maxlen, w.ptr[i + 1] - w.ptr[i]);
  if (maxlen > prm.stage_nnz) { P.why = "a row is longer than a pipeline stage"; return P; }
  P.lanes = 1;
  while (P.lanes < prm.max_lanes && 7.0 * P.lanes < mean) P.lanes *= 2;   // <= ~7 entries per lane: one burst of 8
  std::vector<int> glev;   // wavefront (level-schedule) number of every row in the whole level
  const int D = global_wavefronts(w, glev);
  P.global_wavefronts = D;
  const double bytes_row = 12.0 * mean + 28.0;
  // rows per step one CTA can stream in one step time — but a step should fit ONE pipeline stage (a split step costs a
  // second hand-off)
  const double xcap = prm.cap_step_to_stage ? std::min(0.75 * prm.stage_nnz / std::max(mean, 1.0), 0.75 * prm.stage_rows) : 1e9;
  const double X = std::max(4.0, std::min(xcap, prm.step_us * prm.cta_gbs * 1e3 / bytes_row));
  P.target_step_rows = X;

  // ---- tiles -----------------------------------------------------------------------------------
  std::vector<int> tile_of((size_t)n, 0);
  std::vector<int64_t> Kc, Jc;
  double theta = prm.force_tile_rows > 0 ? 0.0 : valley_threshold(w);
  const bool one_tile = prm.force_tile_rows == 0 && prm.force_a == 0 && (double)n / (double)D <= 1.25 * X;
  if (one_tile) theta = 0.0;
  P.theta = theta;
  if (theta > 0.0) {
    Kc.assign((size_t)n, 0);
    Jc.assign((size_t)n, 0);
    const int64_t th = (int64_t)theta;
    for (int64_t i = 0; i < n; ++i) {
      int64_t k = 0, jj = 0;
      for (int q = w.ptr[i]; q < w.ptr[i + 1]; ++q) {
        const int64_t j = w.idx[q];
        if (j >= i) break;   // columns are sorted
        const int64_t d = i - j;
        if (d > th) { k = std::max(k, Kc[j] + 1); jj = std::max(jj, Jc[j]); }
        else { k = std::max(k, Kc[j]); jj = std::max(jj, Jc[j] + d); }
      }
      Kc[i] = k;
      Jc[i] = jj;
    }
    int64_t km = 0, jm = 0;
    for (int64_t i = 0; i < n; ++i) { km = std::max(km, Kc[i]); jm = std::max(jm, Jc[i]); }
    P.k_extent = (int)(km + 1);
    P.j_extent = (int)(jm + 1);
    if (P.k_extent < 4) theta = P.theta = 0.0;   // no third dimension to speak of: contiguous ranges
  }
  std::vector<int> step((size_t)n, 0);
  std::vector<int> tile_rows_cnt, tile_nsteps, tmin, tmax;   // per tile: rows, steps, first / last wavefront of the level it holds
  int64_t a = 0, b = 0, trows = 0;
  if (theta > 0.0) {
    b = prm.force_b > 0 ? prm.force_b : std::max<int64_t>(1, (int64_t)std::llround(std::sqrt(X)));
    a = prm.force_a > 0 ? prm.force_a : std::max<int64_t>(64, (int64_t)(X * (double)D / 3.0 / (double)b));
    a = std::min<int64_t>(a, P.j_extent);
  } else {
    trows = prm.force_tile_rows > 0 ? prm.force_tile_rows : (one_tile ? n : std::max<int64_t>(256, (int64_t)(X * (double)D / 2.0)));
    trows = std::min<int64_t>(trows, n);
  }
  for (int iter = 0; iter < 4; ++iter) {
    int ntiles = 0;
    if (theta > 0.0) {
      const int64_t nJ = (P.j_extent + a - 1) / a;
      const int64_t nK = (P.k_extent + b - 1) / b;
      std::vector<int> remap((size_t)(nJ * nK), -1);
      std::vector<char> used((size_t)(nJ * nK), 0);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) used[(size_t)((Kc[i] / b) * nJ + Jc[i] / a)] = 1;
      for (size_t q = 0; q < used.size(); ++q)
        if (used[q]) remap[q] = ntiles++;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) tile_of[i] = remap[(size_t)((Kc[i] / b) * nJ + Jc[i] / a)];
    } else {
      ntiles = (int)((n + trows - 1) / trows);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) tile_of[i] = (int)(i / trows);
    }
    // order inside a tile: by the row's wavefront number in the WHOLE level, counted from the tile's first wavefront.  Any
    // schedule in which every dependency points to an earlier step would do inside the tile; this one also lines the tiles
    // up in time: a row never sits earlier in its tile's walk than the rows of other tiles it depends on sit in theirs
    // (an as-soon-as-possible local schedule would put a tile's boundary rows first and make them wait for the END of
    // the neighbouring tile).
    tile_nsteps.assign((size_t)ntiles, 0);
    tile_rows_cnt.assign((size_t)ntiles, 0);
    tmin.assign((size_t)ntiles, INT_MAX);
    tmax.assign((size_t)ntiles, -1);
    for (int64_t i = 0; i < n; ++i) {
      const int t = tile_of[i];
      tmin[t] = std::min(tmin[t], glev[i]);
      tmax[t] = std::max(tmax[t], glev[i]);
      tile_rows_cnt[t]++;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) step[i] = glev[i] - tmin[tile_of[i]];
    for (int t = 0; t < ntiles; ++t) tile_nsteps[t] = tmax[t] - tmin[t] + 1;
    int64_t total_steps = 0;
    for (int t = 0; t < ntiles; ++t) total_steps += tile_nsteps[t];
    P.ntiles = ntiles;
    P.mean_step_rows = (double)n / (double)total_steps;
    if (prm.verbose)
      fprintf(stderr, "[b200amg] block plan iter %d: n=%lld D=%d theta=%.0f K=%d J=%d a=%lld b=%lld tile_rows=%lld tiles=%d rows/step %.1f (target %.1f)\n",
              iter, (long long)n, D, theta, P.k_extent, P.j_extent, (long long)a, (long long)b, (long long)trows, ntiles,
              P.mean_step_rows, X);
    if (prm.force_tile_rows > 0 || prm.force_a > 0 || one_tile || ntiles == 1 || iter == 3) break;
    const double ratio = X / P.mean_step_rows;
    if (ratio < 1.3 && ratio > 0.77) break;
    const double f = std::min(4.0, std::max(0.25, ratio));
    if (theta > 0.0) {
      const int64_t a_new = std::min<int64_t>(P.j_extent, std::max<int64_t>(64, (int64_t)((double)a * f)));
      if (a_new == a) {
        const int64_t b_new = std::max<int64_t>(1, (int64_t)std::llround((double)b * f));
        if (b_new == b) break;
        b = b_new;
      } else {
        a = a_new;
      }
    } else {
      const int64_t t_new = std::min<int64_t>(n, std::max<int64_t>(256, (int64_t)((double)trows * f)));
      if (t_new == trows) break;
      trows = t_new;
    }
  }
  P.block_a = (int)a;
  P.block_b = (int)b;
  const int ntiles = P.ntiles;

  // ---- numbering: (tile, local step, old index) ---------------------------------------------------
  std::vector<int64_t> tile_step_ptr((size_t)ntiles + 1, 0);   // offset of the tile's first step among all steps
  for (int t = 0; t < ntiles; ++t) tile_step_ptr[t + 1] = tile_step_ptr[t] + tile_nsteps[t];
  const int64_t nsteps_raw = tile_step_ptr[ntiles];
  std::vector<int> step_ptr((size_t)nsteps_raw + 1, 0);         // rows per (tile, step) -> offsets
  for (int64_t i = 0; i < n; ++i) step_ptr[(size_t)(tile_step_ptr[tile_of[i]] + step[i]) + 1]++;
  for (int64_t s = 0; s < nsteps_raw; ++s) step_ptr[s + 1] += step_ptr[s];
  P.perm.old_of_new.resize((size_t)n);
  P.perm.new_of_old.resize((size_t)n);
  {
    std::vector<int> next(step_ptr.begin(), step_ptr.end() - 1);
    for (int64_t i = 0; i < n; ++i) {   // ascending old index inside a step
      const int q = next[(size_t)(tile_step_ptr[tile_of[i]] + step[i])]++;
      P.perm.old_of_new[q] = (int)i;
      P.perm.new_of_old[i] = q;
    }
  }
  // row lengths in the new numbering -> nnz offsets
  std::vector<int> nptr((size_t)n + 1, 0);
  for (int64_t q = 0; q < n; ++q) {
    const int r = P.perm.old_of_new[q];
    nptr[q + 1] = nptr[q] + (w.ptr[r + 1] - w.ptr[r]);
  }

  // ---- stages: consecutive steps of a tile, <= stage_nnz entries and <= stage_rows rows; oversize steps are split ----
  P.tile.resize((size_t)ntiles);
  std::vector<int> stage_of_row((size_t)n, 0);   // GLOBAL stage id of every (new) row
  P.max_tile_steps = 0;
  P.max_tile_rows = 0;
  for (int t = 0; t < ntiles; ++t) {
    const int stage_begin = (int)P.stage_meta.size();
    const int row_begin = step_ptr[(size_t)tile_step_ptr[t]];
    const int row_end = step_ptr[(size_t)tile_step_ptr[t + 1]];
    int cur_row0 = row_begin, cur_steps = 0, cur_step_begin = (int)P.steps.size();
    auto close_stage = [&](int row1) {
      if (cur_steps == 0) return;
      P.stage_meta.push_back(BI4{cur_row0, row1, nptr[cur_row0], nptr[row1]});
      P.stage_aux.push_back(BI4{cur_step_begin, cur_steps, 0, 0});
      cur_row0 = row1;
      cur_steps = 0;
      cur_step_begin = (int)P.steps.size();
    };
    for (int64_t s = tile_step_ptr[t]; s < tile_step_ptr[t + 1]; ++s) {
      int r = step_ptr[(size_t)s];
      const int e = step_ptr[(size_t)s + 1];
      while (r < e) {
        // the longest prefix [r, r2) of the step that still fits the open stage
        int r2 = r;
        while (r2 < e && r2 + 1 - cur_row0 <= prm.stage_rows && nptr[r2 + 1] - nptr[cur_row0] <= prm.stage_nnz) ++r2;
        if (r2 < e && cur_steps > 0) { close_stage(r); continue; }   // does not fit behind earlier steps: start a fresh stage
        if (r2 == r) { P.why = "internal: a row does not fit an empty stage"; return P; }
        P.steps.push_back(r2);
        ++cur_steps;
        r = r2;
        if (r < e) close_stage(r);   // the step was split: the rest goes to the next stage
      }
    }
    close_stage(row_end);
    P.tile[t] = BI4{stage_begin, (int)P.stage_meta.size(), row_begin, row_end};
    P.max_tile_steps = std::max(P.max_tile_steps, tile_nsteps[t]);
    P.max_tile_rows = std::max<int64_t>(P.max_tile_rows, row_end - row_begin);
  }
  P.nstages = (int)P.stage_meta.size();
  P.nsteps = (int)P.steps.size();
  P.stage_auxb.assign((size_t)P.nstages, BI2{0, 0});
#pragma omp parallel for schedule(dynamic, 64)
  for (int g = 0; g < P.nstages; ++g)
    for (int r = P.stage_meta[g].x; r < P.stage_meta[g].y; ++r) stage_of_row[r] = g;
  std::vector<int> tile_of_new((size_t)n);
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < n; ++q) tile_of_new[q] = tile_of[P.perm.old_of_new[q]];

  // ---- cross-tile requirements per stage, both directions (only increases are recorded) ----------------
  std::vector<std::vector<BI2>> rf((size_t)ntiles), rb((size_t)ntiles);   // per tile: {stage << 0 .. } flattened below
  std::vector<std::vector<int>> rf_cnt((size_t)ntiles), rb_cnt((size_t)ntiles);
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < ntiles; ++t) {
    const int s0 = P.tile[t].x, s1 = P.tile[t].y;
    std::vector<BI2> have;   // running maximum per predecessor tile
    auto need = [&](std::vector<BI2>& out, int pt, int cnt) {
      for (BI2& hq : have)
        if (hq.x == pt) {
          if (cnt > hq.y) { hq.y = cnt; out.push_back(BI2{pt, cnt}); }
          return;
        }
      have.push_back(BI2{pt, cnt});
      out.push_back(BI2{pt, cnt});
    };
    auto merge_stage = [](std::vector<BI2>& v, size_t from) {   // one entry per tile inside a stage: keep the largest count
      for (size_t i = from; i < v.size(); ++i)
        for (size_t j = i + 1; j < v.size();)
          if (v[j].x == v[i].x) { v[i].y = std::max(v[i].y, v[j].y); v.erase(v.begin() + (long)j); } else ++j;
    };
    rf_cnt[t].assign((size_t)(s1 - s0), 0);
    rb_cnt[t].assign((size_t)(s1 - s0), 0);
    for (int g = s0; g < s1; ++g) {   // forward: stages in ascending order, neighbours with a lower index in other tiles
      const size_t from = rf[t].size();
      for (int r = P.stage_meta[g].x; r < P.stage_meta[g].y; ++r) {
        const int ro = P.perm.old_of_new[r];
        for (int k = w.ptr[ro]; k < w.ptr[ro + 1]; ++k) {
          const int co = w.idx[k];
          if (co >= ro) break;
          const int c = P.perm.new_of_old[co];
          const int pt = tile_of_new[c];
          if (pt != t) need(rf[t], pt, stage_of_row[c] - P.tile[pt].x + 1);
        }
      }
      merge_stage(rf[t], from);
      rf_cnt[t][(size_t)(g - s0)] = (int)(rf[t].size() - from);
    }
    have.clear();
    for (int g = s1 - 1; g >= s0; --g) {   // backward: stages in descending order, neighbours with a higher index
      const size_t from = rb[t].size();
      for (int r = P.stage_meta[g].x; r < P.stage_meta[g].y; ++r) {
        const int ro = P.perm.old_of_new[r];
        for (int k = w.ptr[ro + 1] - 1; k >= w.ptr[ro]; --k) {
          const int co = w.idx[k];
          if (co <= ro) break;
          const int c = P.perm.new_of_old[co];
          const int pt = tile_of_new[c];
          if (pt != t) need(rb[t], pt, P.tile[pt].y - stage_of_row[c]);
        }
      }
      merge_stage(rb[t], from);
      rb_cnt[t][(size_t)(g - s0)] = (int)(rb[t].size() - from);
    }
  }
  for (int t = 0; t < ntiles; ++t) {
    const int s0 = P.tile[t].x, s1 = P.tile[t].y;
    size_t o = 0;
    for (int g = s0; g < s1; ++g) {
      P.stage_aux[g].z = (int)P.req_fwd.size();
      P.stage_aux[g].w = rf_cnt[t][(size_t)(g - s0)];
      for (int q = 0; q < P.stage_aux[g].w; ++q) P.req_fwd.push_back(rf[t][o++]);
    }
    o = 0;
    for (int g = s1 - 1; g >= s0; --g) {
      P.stage_auxb[g].x = (int)P.req_bwd.size();
      P.stage_auxb[g].y = rb_cnt[t][(size_t)(g - s0)];
      for (int q = 0; q < P.stage_auxb[g].y; ++q) P.req_bwd.push_back(rb[t][o++]);
    }
  }
  // ---- ticket order: Kahn's algorithm with a priority queue keyed by the tile's first (forward) / last (backward) wavefront ----
  {
    auto topo = [&](const std::vector<std::vector<BI2>>& reqs, bool fwd, std::vector<int>& order) {
      std::vector<std::vector<int>> succ((size_t)ntiles);
      std::vector<int> indeg((size_t)ntiles, 0);
      for (int t = 0; t < ntiles; ++t) {
        std::vector<int> preds;
        for (const BI2& r : reqs[t]) preds.push_back(r.x);
        std::sort(preds.begin(), preds.end());
        preds.erase(std::unique(preds.begin(), preds.end()), preds.end());
        for (int pt : preds) { succ[(size_t)pt].push_back(t); indeg[t]++; }
      }
      auto key = [&](int t) { return fwd ? (int64_t)tmin[t] * ntiles + t : (int64_t)(D - tmax[t]) * ntiles + (ntiles - 1 - t); };
      std::vector<std::pair<int64_t, int>> heap;
      auto cmp = [](const std::pair<int64_t, int>& x, const std::pair<int64_t, int>& y) { return x.first > y.first; };
      for (int t = 0; t < ntiles; ++t)
        if (indeg[t] == 0) heap.push_back({key(t), t});
      std::make_heap(heap.begin(), heap.end(), cmp);
      order.clear();
      while (!heap.empty()) {
        std::pop_heap(heap.begin(), heap.end(), cmp);
        const int t = heap.back().second;
        heap.pop_back();
        order.push_back(t);
        for (int u : succ[(size_t)t])
          if (--indeg[u] == 0) { heap.push_back({key(u), u}); std::push_heap(heap.begin(), heap.end(), cmp); }
      }
    };
    topo(rf, true, P.order_fwd);
    topo(rb, false, P.order_bwd);
    if ((int)P.order_fwd.size() != ntiles || (int)P.order_bwd.size() != ntiles) { P.why = "internal: the tile graph has a cycle"; return P; }
  }
  if (prm.verbose >= 2)
    for (int t = 0; t < std::min(ntiles, 24); ++t) {
      const int s0 = P.tile[t].x, s1 = P.tile[t].y;
      fprintf(stderr, "[b200amg]   tile %d: rows [%d, %d) stages %d steps %d | fwd requirements of its first stages:", t, P.tile[t].z, P.tile[t].w,
              s1 - s0, tile_nsteps[t]);
      for (int g = s0; g < std::min(s1, s0 + 3); ++g) {
        fprintf(stderr, " [stage %d rows %d steps %d:", g - s0, P.stage_meta[g].y - P.stage_meta[g].x, P.stage_aux[g].y);
        for (int q = 0; q < P.stage_aux[g].w; ++q) fprintf(stderr, " t%d>=%d", P.req_fwd[(size_t)(P.stage_aux[g].z + q)].x, P.req_fwd[(size_t)(P.stage_aux[g].z + q)].y);
        fprintf(stderr, "]");
      }
      fprintf(stderr, "\n");
    }
  P.ok = true;
  return P;
}

// Per-entry codes of the walked matrix in its NEW numbering (`wp`), one array per sweep direction — everything about an entry
// that does not change from sweep to sweep, so that the kernel's scout warps do not have to work it out again per stage:
//   code >= 0                     FAR: the column index; the value of x is gathered from global memory when the stage is staged
//   code = 0x80000000 | slot      NEAR: relaxed by the same tile at most window_eff rows earlier in the sweep; read from the
//                                 shared-memory window, slot = column & (window - 1)
//   code = 0xC0000000             the diagonal
// dpos[row] = position of the row's diagonal in the value array (-1: none).
constexpr int kBlockCodeNear = (int)0x80000000u, kBlockCodeDiag = (int)0xC0000000u;
static inline void build_block_codes(const BlockPlan& P, const HostCsr& wp, hvec<int>& code_fwd, hvec<int>& code_bwd, hvec<int>& dpos) {
  code_fwd.resize(wp.idx.size());   // (no zero fill: the tiles cover every row, every entry is written below)
  code_bwd.resize(wp.idx.size());
  dpos.resize((size_t)P.n);
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < P.n; ++r) dpos[(size_t)r] = -1;
  const int W = P.window, WE = P.window_eff;
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < P.ntiles; ++t) {
    const int r0 = P.tile[t].z, r1 = P.tile[t].w;
    for (int r = r0; r < r1; ++r)
      for (int k = wp.ptr[r]; k < wp.ptr[r + 1]; ++k) {
        const int c = wp.idx[k];
        if (c == r) {
          code_fwd[k] = code_bwd[k] = kBlockCodeDiag;
          dpos[r] = k;
          continue;
        }
        const bool in_tile = c >= r0 && c < r1;
        code_fwd[k] = (in_tile && c < r && r - c <= WE) ? (kBlockCodeNear | (c & (W - 1))) : c;
        code_bwd[k] = (in_tile && c > r && c - r <= WE) ? (kBlockCodeNear | (c & (W - 1))) : c;
      }
  }
}

// Checks every invariant the kernel relies on, against the matrix in its NEW numbering (`wp` = permute_sym(w, plan.perm)).
// Returns an empty string when the plan is sound.
static inline std::string validate_block_plan(const BlockPlan& P, const HostCsr& wp) {
  char buf[256];
  const int64_t n = P.n;
  if (!P.ok) return "plan not ok: " + P.why;
  std::vector<int> tile_of((size_t)n, -1), stage_of((size_t)n, -1), step_of((size_t)n, -1);
  int prev_end = 0;
  for (int t = 0; t < P.ntiles; ++t) {
    const BI4 T = P.tile[t];
    if (T.z != prev_end || T.w < T.z) return "tile rows are not a partition";
    prev_end = T.w;
    int r_expect = T.z;
    for (int g = T.x; g < T.y; ++g) {
      const BI4 m = P.stage_meta[g];
      const BI4 ax = P.stage_aux[g];
      if (m.x != r_expect || m.y <= m.x) return "stage rows are not a partition of the tile";
      if (m.y - m.x > P.stage_rows || m.w - m.z > P.stage_nnz) return "stage exceeds its capacity";
      if (m.z != wp.ptr[m.x] || m.w != wp.ptr[m.y]) return "stage nnz range does not match the matrix";
      int r = m.x;
      for (int s = 0; s < ax.y; ++s) {
        const int e = P.steps[(size_t)(ax.x + s)];
        if (e <= r || e > m.y) return "step boundaries are not increasing inside the stage";
        for (int q = r; q < e; ++q) { tile_of[q] = t; stage_of[q] = g; step_of[q] = ax.x + s; }
        r = e;
      }
      if (r != m.y) return "steps do not cover the stage";
      r_expect = m.y;
    }
    if (r_expect != T.w) return "stages do not cover the tile";
  }
  if (prev_end != n) return "tiles do not cover the level";
  // requirements, replayed cumulatively per tile
  for (int dir = 0; dir < 2; ++dir) {
    const std::vector<int>& order = dir == 0 ? P.order_fwd : P.order_bwd;
    if ((int)order.size() != P.ntiles) return "ticket order is not a permutation of the tiles";
    std::vector<int> ticket((size_t)P.ntiles, -1);
    for (int q = 0; q < P.ntiles; ++q) {
      if (order[q] < 0 || order[q] >= P.ntiles || ticket[order[q]] >= 0) return "ticket order is not a permutation of the tiles";
      ticket[order[q]] = q;
    }
    for (int t = 0; t < P.ntiles; ++t) {
      const BI4 T = P.tile[t];
      std::vector<BI2> have;
      for (int gi = 0; gi < T.y - T.x; ++gi) {
        const int g = dir == 0 ? T.x + gi : T.y - 1 - gi;
        const int rb = dir == 0 ? P.stage_aux[g].z : P.stage_auxb[g].x;
        const int rc = dir == 0 ? P.stage_aux[g].w : P.stage_auxb[g].y;
        const std::vector<BI2>& req = dir == 0 ? P.req_fwd : P.req_bwd;
        for (int q = 0; q < rc; ++q) {
          const BI2 rq = req[(size_t)(rb + q)];
          if (rq.x < 0 || rq.x >= P.ntiles) return "requirement names a tile that does not exist";
          if (dir == 0 ? rq.x >= t : rq.x <= t) return "requirement on a tile that is not earlier in the sweep (cycle)";
          if (ticket[rq.x] >= ticket[t]) return "requirement on a tile with a later ticket (the sweep could deadlock)";
          if (rq.y < 1 || rq.y > P.tile[rq.x].y - P.tile[rq.x].x) return "requirement count out of range";
          bool f = false;
          for (BI2& hq : have)
            if (hq.x == rq.x) { hq.y = std::max(hq.y, rq.y); f = true; }
          if (!f) have.push_back(rq);
        }
        for (int r = P.stage_meta[g].x; r < P.stage_meta[g].y; ++r)
          for (int k = wp.ptr[r]; k < wp.ptr[r + 1]; ++k) {
            const int c = wp.idx[k];
            if (c == r) continue;
            const bool earlier = dir == 0 ? c < r : c > r;
            if (tile_of[c] == t) {
              // same tile: earlier-ordered neighbours sit in an earlier step; near ones are read from the window, the
              // others from global memory, where they must have arrived before the stage can be staged
              if (earlier && !(dir == 0 ? step_of[c] < step_of[r] : step_of[c] > step_of[r])) {
                snprintf(buf, sizeof buf, "row %d and its neighbour %d share a tile but not the step order", r, c);
                return buf;
              }
              const int dist = dir == 0 ? r - c : c - r;
              if (earlier && dist > P.window_eff) {
                const int gap = dir == 0 ? stage_of[r] - stage_of[c] : stage_of[c] - stage_of[r];
                if (gap < P.depth) return "a far neighbour inside the tile may still be in flight when it is gathered";
              }
            } else if (earlier) {
              const int needc = dir == 0 ? stage_of[c] - P.tile[tile_of[c]].x + 1 : P.tile[tile_of[c]].y - stage_of[c];
              int got = 0;
              for (const BI2& hq : have)
                if (hq.x == tile_of[c]) got = hq.y;
              if (got < needc) {
                snprintf(buf, sizeof buf, "row %d (tile %d) needs %d stages of tile %d, the plan waits for %d", r, t, needc, tile_of[c], got);
                return buf;
              }
            } else {
              // later-ordered neighbour in another tile: that tile must come later in the sweep (it reads OUR new value,
              // we read ITS old value before it can have started)
              if (dir == 0 ? tile_of[c] < t : tile_of[c] > t) return "a later-ordered neighbour lives in an earlier tile";
            }
          }
      }
    }
  }
  return "";
}

// Host emulation of gs_block_kernel (block_gs.cuh) for one sweep, tiles walked one after the other in ticket order: the
// same stage / step / window / far-gather logic, so the plan AND the kernel's addressing rules can be checked against the
// sequential sweep without a GPU.  x, b: NEW numbering.  Separate multiply and add, true division (no contraction).
static inline void emulate_block_sweep(const BlockPlan& P, const HostCsr& wp, std::vector<double>& x, const std::vector<double>& b,
                                       double omega, bool sor, bool backward) {
  const int W = P.window, W_EFF = P.window_eff;
  std::vector<double> win((size_t)W, 0.0), xs((size_t)P.stage_nnz + 8, 0.0);
  for (int tk = 0; tk < P.ntiles; ++tk) {
    const int t = backward ? P.order_bwd[(size_t)tk] : P.order_fwd[(size_t)tk];
    const BI4 TT = P.tile[t];
    std::fill(win.begin(), win.end(), std::nan(""));   // nothing of another tile may be read from the window
    for (int i = 0; i < TT.y - TT.x; ++i) {
      const int g = backward ? TT.y - 1 - i : TT.x + i;
      const BI4 m = P.stage_meta[g];
      const BI4 ax = P.stage_aux[g];
      auto near = [&](int row, int c) {
        return backward ? (c > row && c < TT.w && c - row <= W_EFF) : (c < row && c >= TT.z && row - c <= W_EFF);
      };
      // scouts: everything that is not near is gathered from "global memory" before the stage is relaxed
      for (int row = m.x; row < m.y; ++row)
        for (int k = wp.ptr[row]; k < wp.ptr[row + 1]; ++k) {
          const int c = wp.idx[k];
          if (!near(row, c) && (c != row || sor)) xs[(size_t)(k - m.z)] = x[c];
        }
      // compute: steps in sweep order
      for (int s = 0; s < ax.y; ++s) {
        const int sf = backward ? ax.y - 1 - s : s;
        const int lo = sf == 0 ? m.x : P.steps[(size_t)(ax.x + sf - 1)];
        const int hi = P.steps[(size_t)(ax.x + sf)];
        std::vector<double> newx((size_t)(hi - lo));
        for (int row = lo; row < hi; ++row) {
          volatile double rsum = 0.0;
          double d = 0.0, xold = 0.0;
          for (int k = wp.ptr[row]; k < wp.ptr[row + 1]; ++k) {
            const int c = wp.idx[k];
            if (c == row) { d = wp.val[k]; if (sor) xold = xs[(size_t)(k - m.z)]; continue; }
            const double xv = near(row, c) ? win[(size_t)(c & (W - 1))] : xs[(size_t)(k - m.z)];
            volatile double prod = wp.val[k] * xv;
            rsum = rsum + prod;
          }
          double xn;
          if (d != 0.0) {
            volatile double r = b[row] - rsum;
            if (sor) {
              volatile double t1 = (1.0 - omega) * xold, t2 = omega / d;
              volatile double t3 = t2 * r;
              xn = t1 + t3;
            } else {
              xn = r / d;
            }
          } else {
            xn = x[row];
          }
          newx[(size_t)(row - lo)] = xn;
        }
        for (int row = lo; row < hi; ++row) {   // rows of a step are independent: publish after the whole step
          x[row] = newx[(size_t)(row - lo)];
          win[(size_t)(row & (W - 1))] = newx[(size_t)(row - lo)];
        }
      }
    }
  }
}

}  // namespace b200amg
