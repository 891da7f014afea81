// Host-only entry points of libb200amg.so (no CUDA device needed): the halo / send lists one rank of a row partition derives
// (b200amg_partition_plan, _plan_child — what the CPU gloo tests exercise) and the plans of the blocked Gauss-Seidel sweeps
// with their host emulation (b200amg_block_plan_check).
#include "engine_base.h"
#include "staging.h"
#include "partition.h"
#include "block_params.h"

using namespace b200amg;

extern "C" {

// Host-only: the plan one rank of a `world`-way partition would use (no device needed; what the CPU
// world_size-2 tests exercise).  Array capacities: row_split/coarse_split/recv_off/send_off world+1,
// cx_lo/cx_hi world, halo_cols/send_idx `cap` entries.
static void partition_plan_impl(const b200amg_csc_t* A, const b200amg_csc_t* P, const b200amg_csc_t* R, const b200amg_csc_t* parentP,
                                const int64_t* parent_row_split, const int64_t* parent_coarse_split, int32_t rank, int32_t world,
                                int64_t* row_split, int64_t* coarse_split, int32_t* halo_cols, int64_t* nhalo, int32_t* recv_off,
                                int32_t* send_idx, int64_t* nsend, int32_t* send_off, int64_t* cx_lo, int64_t* cx_hi, int64_t cap) {
  REQUIRE(A && P && R && row_split && coarse_split && halo_cols && nhalo && recv_off && send_idx && nsend && send_off && cx_lo && cx_hi,
          B200AMG_ERR_BAD_ARG, "null argument");
  REQUIRE(world >= 1 && rank >= 0 && rank < world, B200AMG_ERR_BAD_ARG, "bad rank / world");
  HostCsr hAt = stage_csc_as_rows_of_transpose(A);
  HostCsr hA = transpose(hAt);
  const bool sym = bit_equal(hA, hAt);
  HostCsr hP = stage_operator_by_rows(P), hR = stage_operator_by_rows(R);
  PartPlan pl;
  if (parentP) {
    REQUIRE(parent_row_split && parent_coarse_split, B200AMG_ERR_BAD_ARG, "a child plan needs the parent's row and coarse splits");
    HostCsr hPP = stage_operator_by_rows(parentP);
    REQUIRE(hPP.ncols == hA.nrows, B200AMG_ERR_DIM_MISMATCH, "the parent's P has %lld columns, this level %lld rows", (long long)hPP.ncols,
            (long long)hA.nrows);
    const std::vector<int64_t> given(parent_coarse_split, parent_coarse_split + world + 1), prs(parent_row_split, parent_row_split + world + 1);
    pl = make_part_plan(rank, world, hA, sym ? nullptr : &hAt, hR, hP, &given, &hPP, &prs);
  } else {
    pl = make_part_plan(rank, world, hA, sym ? nullptr : &hAt, hR, hP);
  }
  REQUIRE((int64_t)pl.halo_cols.size() <= cap && (int64_t)pl.send_idx.size() <= cap, B200AMG_ERR_BAD_ARG, "capacity too small");
  std::copy(pl.row_split.begin(), pl.row_split.end(), row_split);
  std::copy(pl.coarse_split.begin(), pl.coarse_split.end(), coarse_split);
  std::copy(pl.halo_cols.begin(), pl.halo_cols.end(), halo_cols);
  std::copy(pl.recv_off.begin(), pl.recv_off.end(), recv_off);
  std::copy(pl.send_idx.begin(), pl.send_idx.end(), send_idx);
  std::copy(pl.send_off.begin(), pl.send_off.end(), send_off);
  std::copy(pl.cx_lo_all.begin(), pl.cx_lo_all.end(), cx_lo);
  std::copy(pl.cx_hi_all.begin(), pl.cx_hi_all.end(), cx_hi);
  *nhalo = pl.nhalo;
  *nsend = (int64_t)pl.send_idx.size();
}

int32_t b200amg_partition_plan(const b200amg_csc_t* A, const b200amg_csc_t* P, const b200amg_csc_t* R, int32_t rank, int32_t world,
                               int64_t* row_split, int64_t* coarse_split, int32_t* halo_cols, int64_t* nhalo, int32_t* recv_off,
                               int32_t* send_idx, int64_t* nsend, int32_t* send_off, int64_t* cx_lo, int64_t* cx_hi, int64_t cap) {
  API_BEGIN
  partition_plan_impl(A, P, R, nullptr, nullptr, nullptr, rank, world, row_split, coarse_split, halo_cols, nhalo, recv_off, send_idx, nsend,
                      send_off, cx_lo, cx_hi, cap);
  API_END
}

int32_t b200amg_partition_plan_child(const b200amg_csc_t* A, const b200amg_csc_t* P, const b200amg_csc_t* R, const b200amg_csc_t* parent_P,
                                     const int64_t* parent_row_split, const int64_t* parent_coarse_split, int32_t rank, int32_t world,
                                     int64_t* row_split, int64_t* coarse_split, int32_t* halo_cols, int64_t* nhalo, int32_t* recv_off,
                                     int32_t* send_idx, int64_t* nsend, int32_t* send_off, int64_t* cx_lo, int64_t* cx_hi, int64_t cap) {
  API_BEGIN
  REQUIRE(parent_P, B200AMG_ERR_BAD_ARG, "null parent P");
  partition_plan_impl(A, P, R, parent_P, parent_row_split, parent_coarse_split, rank, world, row_split, coarse_split, halo_cols, nhalo,
                      recv_off, send_idx, nsend, send_off, cx_lo, cx_hi, cap);
  API_END
}

// Host-only (no device needed): build the blocked Gauss-Seidel plan of a matrix (block_plan.h), check every invariant
// the kernel relies on and, when x / b are given, run the host emulation of the kernel's sweep (same stage / step / window
// / far-gather rules) so the CPU tests can compare it with the sequential sweep.
int32_t b200amg_block_plan_check(const b200amg_csc_t* A, const int64_t* params, int64_t* stats, int32_t* new_of_old, const double* x,
                                 const double* b, double* x_out, double omega, int32_t sor, int32_t sweep, char* msg, int64_t msg_cap) {
  API_BEGIN
  REQUIRE(A && stats, B200AMG_ERR_BAD_ARG, "null argument");
  if (msg && msg_cap > 0) msg[0] = 0;
  HostCsr w = stage_csc_as_rows_of_transpose(A);   // the rows the "fast" smoothers walk (smoother.jl:81-86)
  REQUIRE(w.nrows == w.ncols, B200AMG_ERR_DIM_MISMATCH, "matrix must be square");
  REQUIRE(symmetry_kind(w) >= 1, B200AMG_ERR_UNSUPPORTED, "the blocked sweep needs a structurally symmetric pattern");
  BlockPlanParams prm = block_params_from_env();
  if (params) {
    if (params[0] > 0) prm.force_tile_rows = (int)params[0];
    if (params[1] > 0) prm.force_a = (int)params[1];
    if (params[2] > 0) prm.force_b = (int)params[2];
    if (params[3] > 0) prm.stage_nnz = (int)params[3];
    if (params[4] > 0) prm.stage_rows = (int)params[4];
    if (params[5] > 0) prm.window = (int)params[5];
    if (params[6] > 0) prm.depth = (int)params[6];
    prm.verbose = (int)params[7];
  }
  BlockPlan P = build_block_plan(w, prm);
  for (int q = 0; q < 16; ++q) stats[q] = 0;
  stats[0] = P.ok;
  if (!P.ok) {
    if (msg && msg_cap > 0) snprintf(msg, (size_t)msg_cap, "%s", P.why.c_str());
    return B200AMG_OK;
  }
  stats[1] = P.ntiles; stats[2] = P.nstages; stats[3] = P.nsteps; stats[4] = P.lanes; stats[5] = P.global_wavefronts;
  stats[6] = (int64_t)(1000.0 * P.mean_step_rows); stats[7] = (int64_t)P.theta; stats[8] = P.block_a; stats[9] = P.block_b;
  stats[10] = P.max_tile_rows; stats[11] = P.max_tile_steps; stats[12] = (int64_t)P.req_fwd.size(); stats[13] = (int64_t)P.req_bwd.size();
  stats[14] = P.k_extent; stats[15] = P.j_extent;
  HostCsr wp = permute_sym(w, P.perm);
  const std::string err = validate_block_plan(P, wp);
  if (!err.empty()) {
    stats[0] = -1;
    if (msg && msg_cap > 0) snprintf(msg, (size_t)msg_cap, "%s", err.c_str());
  }
  if (new_of_old) std::copy(P.perm.new_of_old.begin(), P.perm.new_of_old.end(), new_of_old);
  if (x && b && x_out) {
    const int64_t n = w.nrows;
    std::vector<double> xp((size_t)n), bp((size_t)n);
    for (int64_t q = 0; q < n; ++q) { xp[q] = x[P.perm.old_of_new[q]]; bp[q] = b[P.perm.old_of_new[q]]; }
    if (params && params[8] == 1) {   // the pass sweep's layout and addressing rules (pass_plan.h) on the same plan
      PassPlan Q = build_pass_plan(P, wp, kPgWinOff, kPgZeroOff);
      if (!Q.ok) {
        stats[0] = -2;
        if (msg && msg_cap > 0) snprintf(msg, (size_t)msg_cap, "pass plan: %s", Q.why.c_str());
        return B200AMG_OK;
      }
      std::vector<double> dg((size_t)n, 0.0);
      for (int64_t q = 0; q < n; ++q)
        for (int k = wp.ptr[q]; k < wp.ptr[q + 1]; ++k)
          if (wp.idx[k] == q) dg[(size_t)q] = wp.val[k];
      std::string e2;
      if (sweep == 1 || sweep == 3) e2 = emulate_pass_sweep(P, Q, wp, xp, bp, dg, omega, sor != 0, false, kPgWinOff, kPgZeroOff);
      if (e2.empty() && (sweep == 2 || sweep == 3)) e2 = emulate_pass_sweep(P, Q, wp, xp, bp, dg, omega, sor != 0, true, kPgWinOff, kPgZeroOff);
      if (!e2.empty()) {
        stats[0] = -3;
        if (msg && msg_cap > 0) snprintf(msg, (size_t)msg_cap, "pass emulation: %s", e2.c_str());
      }
      stats[4] = Q.lanes; stats[2] = 0; stats[3] = Q.npasses;
      stats[12] = (int64_t)Q.dir[0].req.size(); stats[13] = (int64_t)Q.dir[1].req.size();
      if (prm.verbose) {   // timing model of the schedule: where a sweep's time would go
        const double tp = env_int("B200AMG_MODEL_TPASS_NS", 250) * 1e-3;
        const int ncs[3] = {148, 148, 100000};
        const double lams[3] = {1.5, 0.0, 1.5};
        for (int q = 0; q < 3; ++q) {
          double busy = 0, wf = 0;
          const double us = simulate_pass_sweep(P, Q, false, ncs[q], tp, lams[q], 3.0, 2, &busy, &wf);
          fprintf(stderr, "[b200amg] pass model: n=%lld lanes=%d tiles=%d passes=%lld wavefronts=%d | CTAs %d lam %.1f t_pass %.2f us -> forward sweep %.1f us "
                  "(%.2f us per wavefront), CTAs busy %.0f %%, mean wait before a tile's first pass %.1f us\n", (long long)P.n, Q.lanes, P.ntiles,
                  (long long)Q.npasses, P.global_wavefronts, ncs[q], lams[q], tp, us, us / std::max(1, P.global_wavefronts), 100.0 * busy, wf);
        }
      }
    } else {
      if (sweep == 1 || sweep == 3) emulate_block_sweep(P, wp, xp, bp, omega, sor != 0, false);
      if (sweep == 2 || sweep == 3) emulate_block_sweep(P, wp, xp, bp, omega, sor != 0, true);
    }
    for (int64_t q = 0; q < n; ++q) x_out[P.perm.old_of_new[q]] = xp[q];
  }
  API_END
}


}  // extern "C"
