// Data layout of the PASS sweep (pass_gs.cuh) — host side, pure C++.  Built on top of the blocked plan (block_plan.h), of which
// it keeps the tiles, the renumbering (tile, local step, old index) and the ticket orders; stages and their requirements are
// replaced by PER-PASS requirements (a stage-granular hand-off between tiles triples the critical path: see simulate_pass_sweep).
//
// A PASS is what one group of threads relaxes at once: <= kPassRows / T rows of ONE local step of a tile (T lanes per row).
// Everything is laid out per SWEEP DIRECTION, in the order that sweep walks it (the kernel has no notion of direction):
//   * the entries of a pass form two dense slabs, one word per thread and slot, consecutive threads consecutive words
//     (width = rows * T rounded up to a warp):
//       FAR  slots (<= kPassFar):  value + COLUMN INDEX — x is loaded from global memory a few passes ahead of its use: old values
//            of later-ordered neighbours, new values of other tiles (guarded by the pass requirements), new values of this tile
//            more than kPassNear passes old.  Padding: value 0, column = a row of the pass.
//       NEAR slots (<= kPassNearSlots): value + BYTE OFFSET of the x value in the shared-memory window — relaxed by this tile at
//            most kPassNear passes earlier; read after the hand-off of the previous pass.  Padding: value 0, offset of a zero.
//     The entries of a row are dealt to its T lanes round-robin inside each class, in the reference's entry order
//     (src/smoother.jl:81-86); the diagonal is not stored (the engine keeps a per-row array).
//   * the slab of a pass (far slots, then near slots) is one contiguous piece of the value / index arrays: the unit the
//     producer thread moves into shared memory with bulk copies;
//   * per pass the list of {tile, passes of it that must be complete} it needs from OTHER tiles (only increases are recorded).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "block_plan.h"

namespace b200amg {

constexpr int kPassGroup = 256;          // threads of one compute group: 7 working warps + 1 signalling warp
constexpr int kPassRows = 224;           // threads of a group that relax rows
constexpr int kPassFar = 6;              // far slots per lane and pass (registers)
constexpr int kPassNearSlots = 4;        // near slots per lane and pass (registers)
constexpr int kPassNear = 4;             // a value this many passes old (or younger) is read from the window
constexpr int kPassWindow = 2048;        // x values of the tile kept in shared memory (power of two)
static_assert(kPassWindow >= (kPassNear + 2) * kPassRows, "window too small");
// layout of gs_pass_kernel's dynamic shared memory (pass_gs.cuh): byte ring of whole passes, then the x window, then a zero
constexpr int kPgRingBytes = 192 * 1024;      // slabs + right-hand side + diagonal of the passes in flight
constexpr int kPgWinOff = kPgRingBytes;       // byte offsets inside the dynamic shared memory
constexpr int kPgZeroOff = kPgWinOff + kPassWindow * (int)sizeof(double);

struct PassDir {              // one sweep direction, everything in SWEEP order
  std::vector<BI2> tile;      // per tile: {first pass, end pass}                                            (global indices)
  std::vector<BI4> pass;      // {first row, rows | far slots << 16 | near slots << 24, offset of its slab in val / idx, 0}
  std::vector<BI2> preq;      // per pass: {first requirement, requirements}
  std::vector<BI2> req;       // {tile, passes of it that must be complete}
  std::vector<double> val;    // slabs: per pass far slots, then near slots
  std::vector<int> idx;       // column index (far) / byte offset into the window (near)
  int64_t nentries = 0;
};

struct PassPlan {
  bool ok = false;
  std::string why;
  int lanes = 1;
  int64_t npasses = 0, nnz_stored = 0;
  PassDir dir[2];             // [0] forward, [1] backward
};

// wp: the walked matrix in the plan's NEW numbering (permute_sym(w, P.perm)), sorted columns.  zero_off / win_off: byte offsets
// of the 16 zero bytes / of the window inside the kernel's dynamic shared memory (pass_gs.cuh).
static inline PassPlan build_pass_plan(const BlockPlan& P, const HostCsr& wp, int win_off, int zero_off) {
  PassPlan Q;
  if (!P.ok) { Q.why = "no blocked plan: " + P.why; return Q; }
  const int64_t n = P.n;
  std::vector<int> tile_of((size_t)n, 0);
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < P.ntiles; ++t)
    for (int r = P.tile[(size_t)t].z; r < P.tile[(size_t)t].w; ++r) tile_of[(size_t)r] = t;
  // the plan's steps (rows of a tile with the same local step, contiguous in the new numbering; a step the plan split at a
  // stage boundary stays split — harmless)
  std::vector<int> step_end((size_t)n, 0), step_begin((size_t)n, 0);
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < P.ntiles; ++t) {
    const BI4 TT = P.tile[(size_t)t];
    for (int g = TT.x; g < TT.y; ++g) {
      const BI4 m = P.stage_meta[(size_t)g];
      const BI4 ax = P.stage_aux[(size_t)g];
      int lo = m.x;
      for (int s = 0; s < ax.y; ++s) {
        const int hi = P.steps[(size_t)(ax.x + s)];
        for (int r = lo; r < hi; ++r) { step_begin[(size_t)r] = lo; step_end[(size_t)r] = hi; }
        lo = hi;
      }
    }
  }
  // ---- lanes: the smallest T for which every row's far / near entries fit the slots, in both directions.  The near / far
  // split depends on the passes, which depend on T: try T = 1, 2, 4, ...
  for (int T = 1; T <= 32; T *= 2) {
    const int RP = kPassRows / T;
    bool fits = true;
    for (int d = 0; d < 2 && fits; ++d) {
      PassDir& D = Q.dir[d];
      D = PassDir();
      const bool bwd = d == 1;
      std::vector<int> pass_of_row((size_t)n, 0);   // LOCAL pass index (sweep order) inside the row's tile
      D.tile.assign((size_t)P.ntiles, BI2{0, 0});
      for (int t = 0; t < P.ntiles; ++t) {
        const BI4 TT = P.tile[(size_t)t];
        const int pass_begin = (int)D.pass.size();
        if (!bwd) {
          int r = TT.z;
          while (r < TT.w) {
            const int e = step_end[(size_t)r];
            for (int r0 = r; r0 < e; r0 += RP) {
              const int r1 = std::min(e, r0 + RP);
              const int pl = (int)D.pass.size() - pass_begin;
              D.pass.push_back(BI4{r0, r1 - r0, 0, 0});
              for (int q = r0; q < r1; ++q) pass_of_row[(size_t)q] = pl;
            }
            r = e;
          }
        } else {
          int r = TT.w;
          while (r > TT.z) {
            const int b0 = step_begin[(size_t)(r - 1)];
            for (int r1 = r; r1 > b0; r1 -= RP) {
              const int r0 = std::max(b0, r1 - RP);
              const int pl = (int)D.pass.size() - pass_begin;
              D.pass.push_back(BI4{r0, r1 - r0, 0, 0});
              for (int q = r0; q < r1; ++q) pass_of_row[(size_t)q] = pl;
            }
            r = b0;
          }
        }
        D.tile[(size_t)t] = BI2{pass_begin, (int)D.pass.size()};
      }
      const int64_t np = (int64_t)D.pass.size();
      // near = same tile, earlier in the sweep, at most kPassNear passes earlier
      auto is_earlier = [&](int r, int c) { return bwd ? c > r : c < r; };
      auto is_near = [&](int r, int c) {
        return tile_of[(size_t)c] == tile_of[(size_t)r] && is_earlier(r, c) && pass_of_row[(size_t)r] - pass_of_row[(size_t)c] <= kPassNear;
      };
      std::vector<int> nfar((size_t)np, 0), nnear((size_t)np, 0);
      bool ok_slots = true;
#pragma omp parallel for schedule(dynamic, 64) reduction(&& : ok_slots)
      for (int64_t p = 0; p < np; ++p) {
        const BI4 pr = D.pass[(size_t)p];
        int mf = 0, mn = 0;
        for (int r = pr.x; r < pr.x + pr.y; ++r) {
          int f = 0, nr = 0;
          for (int k = wp.ptr[r]; k < wp.ptr[r + 1]; ++k) {
            const int c = wp.idx[k];
            if (c == r) continue;
            if (is_near(r, c)) ++nr; else ++f;
          }
          mf = std::max(mf, (f + T - 1) / T);
          mn = std::max(mn, (nr + T - 1) / T);
        }
        nfar[(size_t)p] = mf;
        nnear[(size_t)p] = mn;
        if (mf > kPassFar || mn > kPassNearSlots) ok_slots = false;
      }
      if (!ok_slots) { fits = false; break; }
      // slab offsets
      int64_t ell = 0;
      for (int64_t p = 0; p < np; ++p) {
        BI4& pr = D.pass[(size_t)p];
        const int rows = pr.y;
        const int width = (rows * T + 31) & ~31;
        pr.y = rows | (nfar[(size_t)p] << 16) | (nnear[(size_t)p] << 24);
        pr.z = (int)ell;
        pr.w = 0;
        ell += (int64_t)(nfar[(size_t)p] + nnear[(size_t)p]) * width;
        if (ell > (int64_t)0x7ffffff0) { Q.why = "slab arrays exceed 2^31 entries"; return Q; }
      }
      D.nentries = ell;
      D.val.assign((size_t)ell + 64, 0.0);
      D.idx.assign((size_t)ell + 64, 0);
      D.preq.assign((size_t)np, BI2{0, 0});
      std::vector<std::vector<BI2>> treq((size_t)P.ntiles);        // per tile: requirements of its passes, concatenated
      std::vector<std::vector<int>> treq_cnt((size_t)P.ntiles);    // per tile: requirements per pass
      int64_t stored = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : stored)
      for (int t = 0; t < P.ntiles; ++t) {
        const BI2 QT = D.tile[(size_t)t];
        std::vector<BI2> have;   // running maximum per predecessor tile
        treq_cnt[(size_t)t].assign((size_t)(QT.y - QT.x), 0);
        for (int p = QT.x; p < QT.y; ++p) {
          const BI4 pr = D.pass[(size_t)p];
          const int rows = pr.y & 0xffff, jf = (pr.y >> 16) & 0xff, jn = (pr.y >> 24) & 0xff;
          const int width = (rows * T + 31) & ~31;
          const int64_t base = pr.z;
          const int64_t nbase = base + (int64_t)jf * width;
          // padding first: far slots point at a row of the pass (value 0), near slots at the zero
          for (int j = 0; j < jf; ++j)
            for (int q = 0; q < width; ++q) D.idx[(size_t)(base + (int64_t)j * width + q)] = pr.x + std::min(q / T, rows - 1);
          for (int j = 0; j < jn; ++j)
            for (int q = 0; q < width; ++q) D.idx[(size_t)(nbase + (int64_t)j * width + q)] = zero_off;
          const size_t req_from = treq[(size_t)t].size();
          for (int q = 0; q < rows; ++q) {
            const int r = pr.x + q;
            int ef = 0, en = 0;
            for (int k = wp.ptr[r]; k < wp.ptr[r + 1]; ++k) {
              const int c = wp.idx[k];
              if (c == r) continue;
              ++stored;
              if (is_near(r, c)) {
                const int64_t at = nbase + (int64_t)(en / T) * width + (int64_t)q * T + en % T;
                D.val[(size_t)at] = wp.val[k];
                D.idx[(size_t)at] = win_off + 8 * (c & (kPassWindow - 1));
                ++en;
                continue;
              }
              const int64_t at = base + (int64_t)(ef / T) * width + (int64_t)q * T + ef % T;
              D.val[(size_t)at] = wp.val[k];
              D.idx[(size_t)at] = c;
              ++ef;
              const int pt = tile_of[(size_t)c];
              if (pt == t || !is_earlier(r, c)) continue;
              // a new value of another tile: that tile's pass must be complete
              const int cnt = pass_of_row[(size_t)c] + 1;
              BI2* hq = nullptr;
              for (BI2& h2 : have)
                if (h2.x == pt) hq = &h2;
              if (hq && cnt <= hq->y) continue;
              if (hq) hq->y = cnt; else have.push_back(BI2{pt, cnt});
              bool merged = false;   // one record per tile inside a pass
              for (size_t z = req_from; z < treq[(size_t)t].size(); ++z)
                if (treq[(size_t)t][z].x == pt) { treq[(size_t)t][z].y = cnt; merged = true; }
              if (!merged) treq[(size_t)t].push_back(BI2{pt, cnt});
            }
          }
          treq_cnt[(size_t)t][(size_t)(p - QT.x)] = (int)(treq[(size_t)t].size() - req_from);
        }
      }
      for (int t = 0; t < P.ntiles; ++t) {
        const BI2 QT = D.tile[(size_t)t];
        size_t o = 0;
        for (int p = QT.x; p < QT.y; ++p) {
          const int cnt = treq_cnt[(size_t)t][(size_t)(p - QT.x)];
          D.preq[(size_t)p] = BI2{(int)D.req.size(), cnt};
          for (int z = 0; z < cnt; ++z) D.req.push_back(treq[(size_t)t][o++]);
        }
      }
      D.req.push_back(BI2{0, 0});
      Q.nnz_stored = stored;
      Q.npasses = np;
    }
    if (fits) {
      Q.lanes = T;
      Q.ok = true;
      return Q;
    }
  }
  Q.why = "a row has more far / near entries than 32 lanes hold";
  Q.dir[0] = PassDir();
  Q.dir[1] = PassDir();
  return Q;
}

// Host emulation of gs_pass_kernel for one sweep: tiles in ticket order, passes in sweep order, x values taken from exactly
// where the kernel takes them.  The window remembers which row each slot holds; FAR reads are checked against what the kernel's
// look-ahead can legally see (a value of this tile younger than kPassNear + 1 passes must be NEAR; a new value of another tile
// must be covered by a requirement).  Returns an empty string, or what went wrong.  x, b, diag: NEW numbering.
static inline std::string emulate_pass_sweep(const BlockPlan& P, const PassPlan& Q, const HostCsr& wp, std::vector<double>& x,
                                             const std::vector<double>& b, const std::vector<double>& diag, double omega, bool sor,
                                             bool backward, int win_off, int zero_off) {
  char buf[256];
  const int T = Q.lanes;
  const PassDir& D = Q.dir[backward ? 1 : 0];
  std::vector<double> win((size_t)kPassWindow, 0.0);
  std::vector<int> win_row((size_t)kPassWindow, -1);
  std::vector<int> done((size_t)P.ntiles, 0);          // passes complete per tile
  std::vector<int> tile_of((size_t)P.n, -1), pass_of_row((size_t)P.n, -1);
  std::vector<char> seen((size_t)P.n, 0);
  for (int t = 0; t < P.ntiles; ++t) {
    const BI2 QT = D.tile[(size_t)t];
    for (int p = QT.x; p < QT.y; ++p)
      for (int r = D.pass[(size_t)p].x; r < D.pass[(size_t)p].x + (D.pass[(size_t)p].y & 0xffff); ++r) {
        if (r < P.tile[(size_t)t].z || r >= P.tile[(size_t)t].w) return "a pass holds a row of another tile";
        if (pass_of_row[(size_t)r] >= 0) return "a row belongs to two passes";
        tile_of[(size_t)r] = t;
        pass_of_row[(size_t)r] = p - QT.x;
      }
  }
  for (int64_t r = 0; r < P.n; ++r)
    if (pass_of_row[(size_t)r] < 0) return "a row belongs to no pass";
  const std::vector<int>& order = backward ? P.order_bwd : P.order_fwd;
  for (int tk = 0; tk < P.ntiles; ++tk) {
    const int t = order[(size_t)tk];
    const BI2 QT = D.tile[(size_t)t];
    std::fill(win_row.begin(), win_row.end(), -1);
    std::vector<BI2> have;
    for (int p = QT.x; p < QT.y; ++p) {
      const BI4 pr = D.pass[(size_t)p];
      const int pl = p - QT.x;
      const int rows = pr.y & 0xffff, jf = (pr.y >> 16) & 0xff, jn = (pr.y >> 24) & 0xff;
      if (rows < 1 || rows * T > kPassRows || jf > kPassFar || jn > kPassNearSlots) return "pass shape out of range";
      const int width = (rows * T + 31) & ~31;
      if ((int64_t)pr.z + (int64_t)(jf + jn) * width > D.nentries || (pr.z & 31)) return "pass slab out of range";
      for (int z = 0; z < D.preq[(size_t)p].y; ++z) {
        const BI2 rq = D.req[(size_t)(D.preq[(size_t)p].x + z)];
        if (rq.x < 0 || rq.x >= P.ntiles || rq.x == t) return "bad requirement";
        if (done[(size_t)rq.x] < rq.y) {
          snprintf(buf, sizeof buf, "tile %d pass %d needs %d passes of tile %d, which has %d when the tile is reached in ticket order", t, pl,
                   rq.y, rq.x, done[(size_t)rq.x]);
          return buf;
        }
        bool f = false;
        for (BI2& hq : have)
          if (hq.x == rq.x) { hq.y = std::max(hq.y, rq.y); f = true; }
        if (!f) have.push_back(rq);
      }
      const int64_t base = pr.z, nbase = base + (int64_t)jf * width;
      std::vector<double> newx((size_t)rows);
      int entries_seen = 0;
      for (int q = 0; q < rows; ++q) {
        const int r = pr.x + q;
        double lane_sum[32];
        for (int l = 0; l < T; ++l) {
          volatile double far = 0.0, near = 0.0;
          for (int j = 0; j < jf; ++j) {
            const int64_t at = base + (int64_t)j * width + (int64_t)q * T + l;
            const int c = D.idx[(size_t)at];
            const double v = D.val[(size_t)at];
            if (c < 0 || c >= P.n) return "far column out of range";
            if (v == 0.0 && c == r) continue;   // padding
            ++entries_seen;
            if (tile_of[(size_t)c] == t) {
              const int dp = pl - pass_of_row[(size_t)c];
              if (dp > 0 && dp <= kPassNear) {
                snprintf(buf, sizeof buf, "row %d reads row %d (%d passes old) from global memory", r, c, dp);
                return buf;
              }
              if (dp == 0) return "a row depends on a row of its own pass";
            } else if (backward ? c > r : c < r) {
              int got = 0;
              for (const BI2& hq : have)
                if (hq.x == tile_of[(size_t)c]) got = hq.y;
              if (got < pass_of_row[(size_t)c] + 1) {
                snprintf(buf, sizeof buf, "row %d (tile %d) reads row %d of tile %d without a requirement that covers it", r, t, c, tile_of[(size_t)c]);
                return buf;
              }
            }
            volatile double prod = v * x[(size_t)c];
            far = far + prod;
          }
          for (int j = 0; j < jn; ++j) {
            const int64_t at = nbase + (int64_t)j * width + (int64_t)q * T + l;
            const int off = D.idx[(size_t)at];
            const double v = D.val[(size_t)at];
            if (off == zero_off) {
              if (v != 0.0) return "padding with a non-zero value";
              continue;
            }
            ++entries_seen;
            const int slot = (off - win_off) / 8;
            if (off < win_off || (off - win_off) % 8 || slot >= kPassWindow) return "near offset outside the window";
            const int cr = win_row[(size_t)slot];
            if (cr < 0) return "window slot never written";
            const int dp = pl - pass_of_row[(size_t)cr];
            if (dp < 1 || dp > kPassNear) {
              snprintf(buf, sizeof buf, "row %d reads window slot %d holding row %d (%d passes old)", r, slot, cr, dp);
              return buf;
            }
            volatile double prod = v * win[(size_t)slot];
            near = near + prod;
          }
          lane_sum[l] = far + near;
        }
        volatile double rsum = 0.0;
        for (int l = 0; l < T; ++l) rsum = rsum + lane_sum[l];
        const double d = diag[(size_t)r];
        double xn;
        if (d != 0.0) {
          volatile double res = b[(size_t)r] - rsum;
          if (sor) {
            volatile double t1 = (1.0 - omega) * x[(size_t)r], t2 = omega / d;
            volatile double t3 = t2 * res;
            xn = t1 + t3;
          } else {
            xn = res / d;
          }
        } else {
          xn = x[(size_t)r];
        }
        newx[(size_t)q] = xn;
        if (seen[(size_t)r]) return "a row is relaxed twice";
        seen[(size_t)r] = 1;
      }
      int want = 0;
      for (int r = pr.x; r < pr.x + rows; ++r)
        for (int k = wp.ptr[r]; k < wp.ptr[r + 1]; ++k) want += wp.idx[k] != r;
      if (want != entries_seen) {
        snprintf(buf, sizeof buf, "pass %d of tile %d stores %d entries, its rows have %d", pl, t, entries_seen, want);
        return buf;
      }
      for (int q = 0; q < rows; ++q) {
        const int r = pr.x + q;
        x[(size_t)r] = newx[(size_t)q];
        win[(size_t)(r & (kPassWindow - 1))] = newx[(size_t)q];
        win_row[(size_t)(r & (kPassWindow - 1))] = r;
      }
      done[(size_t)t] = pl + 1;
    }
  }
  return "";
}

// Timing model of gs_pass_kernel's schedule (no arithmetic): `ncta` CTAs claim tiles in ticket order; a pass takes t_pass; pass i
// of a tile cannot start before the requirements of pass i + look are met (its group issues that pass's far loads first); a pass
// of another tile becomes visible `lam` after it ends; claiming a tile costs t_tile.  Returns the sweep time in the unit of the
// arguments; *busy = sum of pass times / (ncta * sweep time).
static inline double simulate_pass_sweep(const BlockPlan& P, const PassPlan& Q, bool backward, int ncta, double t_pass, double lam, double t_tile,
                                         int look, double* busy, double* wait_first) {
  const PassDir& D = Q.dir[backward ? 1 : 0];
  const std::vector<int>& order = backward ? P.order_bwd : P.order_fwd;
  std::vector<std::vector<double>> pub((size_t)P.ntiles);   // per tile: time its k-th pass (sweep order) is visible elsewhere
  std::vector<double> free_at((size_t)ncta, 0.0);
  double end = 0.0, work = 0.0, wfirst = 0.0;
  for (int tk = 0; tk < P.ntiles; ++tk) {
    const int t = order[(size_t)tk];
    const BI2 QT = D.tile[(size_t)t];
    const int np = QT.y - QT.x;
    int c = 0;
    for (int q = 1; q < ncta; ++q)
      if (free_at[(size_t)q] < free_at[(size_t)c]) c = q;
    const double t0 = free_at[(size_t)c] + t_tile;
    std::vector<double> gate((size_t)np, 0.0);
    double run = 0.0;
    for (int i = 0; i < np; ++i) {
      const BI2 pq = D.preq[(size_t)(QT.x + i)];
      for (int z = 0; z < pq.y; ++z) {
        const BI2 rq = D.req[(size_t)(pq.x + z)];
        run = std::max(run, pub[(size_t)rq.x][(size_t)(rq.y - 1)]);
      }
      gate[(size_t)i] = run;
    }
    pub[(size_t)t].assign((size_t)np, 0.0);
    double now = t0;
    for (int i = 0; i < np; ++i) {
      const double start = std::max(now, gate[(size_t)std::min(np - 1, i + look)]);
      if (i == 0) wfirst += start - t0;
      now = start + t_pass;
      work += t_pass;
      pub[(size_t)t][(size_t)i] = now + lam;
    }
    free_at[(size_t)c] = now;
    end = std::max(end, now);
  }
  if (busy) *busy = end > 0 ? work / (end * ncta) : 0.0;
  if (wait_first) *wait_first = wfirst / std::max(1, P.ntiles);
  return end;
}

}  // namespace b200amg
