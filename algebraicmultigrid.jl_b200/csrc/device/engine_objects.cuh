// Device-resident objects of a hierarchy: compressed operators and their tile plans, the smoother matrices with their sweep
// schedules / plans (upload-time work), levels, partitioned levels and the handle.  Part of engine.cu (one translation unit).
#pragma once

struct DevCsr {
  int64_t nrows = 0, ncols = 0, nnz = 0;
  int* ptr = nullptr;
  int* idx = nullptr;
  double* val = nullptr;
  float* val32 = nullptr;   // the same values in binary32 when EVERY one of them is exactly representable (else nullptr): the
                            // stream kernels then read 8 instead of 12 bytes per entry and compute the same fp64 products
  int lanes = 8;  // lanes per row of the vector kernels
  // tile plan of the TMA stream kernels (stream.cuh); ntiles == 0: not streamable (a row > kTileNnz)
  int4* meta = nullptr;
  int ntiles = 0;
  // row-partitioned levels: the same tiles sorted into INTERIOR ones (no column in the halo part of the vector) and BOUNDARY
  // ones; the interior kernel runs while the halo exchange is in flight (meta_split = [interior..., boundary...])
  int4* meta_split = nullptr;
  int ntiles_int = 0, ntiles_bnd = 0;
  int stream_lanes = 1;
  int stream_burst = 8;
  int64_t halo_begin = -1;
  bool owner = false;
  void upload(const HostCsr& h, int64_t halo_start = -1) {
    nrows = h.nrows; ncols = h.ncols; nnz = h.nnz();
    halo_begin = halo_start;
    ptr = dev_upload(h.ptr, 8);
    idx = dev_upload(h.idx, 8);
    val = dev_upload(h.val, 8);
    owner = true;
    if (env_int("B200AMG_FP32_STORAGE", 0) && nnz > 0) {   // lossless narrow storage (opt-in: see H::fp32_storage)
      bool exact = true;
      const int64_t nz = nnz;
#pragma omp parallel for schedule(static) reduction(&& : exact)
      for (int64_t k = 0; k < nz; ++k) exact = exact && ((double)(float)h.val[(size_t)k] == h.val[(size_t)k]);
      if (exact) {
        std::vector<float> v32((size_t)nz);
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < nz; ++k) v32[(size_t)k] = (float)h.val[(size_t)k];
        val32 = dev_upload(v32, 16);
      }
    }
    const double mean = nrows ? (double)nnz / (double)nrows : 0.0;
    lanes = 2;
    while (lanes < 32 && lanes < mean) lanes *= 2;
    plan_tiles(h, mean);
  }
  void plan_tiles(const HostCsr& h, double mean) {
    ntiles = 0;
    if (nrows == 0 || env_int("B200AMG_NO_STREAM", 0)) return;
    // lanes per row.  Measured (tools/tune_kernels.py, 256^3 RS hierarchy): stencil rows (<= 8 entries) are
    // fastest with one thread per row and a single gather burst; for 19-110 entries per row FEWER lanes with the
    // unrolled loop beat more lanes with bursts (4070 vs 3830 GB/s at 19 entries per row).
    stream_lanes = 1;
    while (stream_lanes < 32 && mean > 12.0 * stream_lanes) stream_lanes *= 2;
    stream_lanes = env_int("B200AMG_STREAM_LANES", stream_lanes);
    stream_burst = ((stream_lanes == 2 || stream_lanes == 4) && env_int("B200AMG_STREAM_BURST16", 1)) ? 16 : 8;
    const int G = kStreamThreads / stream_lanes;
    int passes = (int)(kTileNnz / std::max(1.0, G * std::max(mean, 1.0)));
    passes = std::min(std::max(passes, 1), 2);   // the kernel prefetches the epilogue operands of two passes
    const int rows_per_tile = std::min(G * passes, kTileRowsMax);
    std::vector<int4> m;
    m.reserve((size_t)(nnz / kTileNnz + nrows / rows_per_tile + 2));
    int64_t r = 0;
    while (r < nrows) {
      int64_t e = r;
      const int k0 = h.ptr[r];
      while (e < nrows && e - r < rows_per_tile && h.ptr[e + 1] - k0 <= kTileNnz) ++e;
      if (e == r) return;  // a single row exceeds the tile: leave ntiles = 0 (vector kernels take over)
      m.push_back(make_int4((int)r, (int)e, k0, h.ptr[e]));
      r = e;
    }
    meta = dev_upload(m);
    ntiles = (int)m.size();
    if (halo_begin >= 0) {
      std::vector<int4> mi, mb;
      for (const int4& t : m) {
        bool bnd = false;
        for (int k = t.z; k < t.w && !bnd; ++k) bnd = h.idx[k] >= halo_begin;
        (bnd ? mb : mi).push_back(t);
      }
      ntiles_int = (int)mi.size();
      ntiles_bnd = (int)mb.size();
      mi.insert(mi.end(), mb.begin(), mb.end());
      meta_split = dev_upload(mi);
    }
  }
  void alias(const DevCsr& o) { *this = o; owner = false; }
  void release() {
    if (owner) { cudaFree(ptr); cudaFree(idx); cudaFree(val); cudaFree(val32); cudaFree(meta); cudaFree(meta_split); }
    ptr = idx = nullptr; val = nullptr; val32 = nullptr; meta = nullptr; meta_split = nullptr; ntiles = ntiles_int = ntiles_bnd = 0; owner = false;
  }
};

struct SweepItem {
  int lv_begin, lv_end;  // wavefront range (in sweep order)
  bool single_cta;
};
// One sweep direction over a level that has been renumbered into wavefront order: wavefront w of the
// forward sweep is the contiguous row range [fwd_lvlptr[w], fwd_lvlptr[w+1]); the backward sweep takes the
// same ranges last to first.
struct DevSchedule {
  int nlev = 0;
  int64_t n = 0;
  int backward = 0;
  int* rows = nullptr;     // rows in sweep order (identity or reversed blocks): per-wavefront fallback kernels only
  int* lvlptr = nullptr;
  std::vector<int> h_lvlptr;
  std::vector<SweepItem> items;
  bool built = false;
  // dataflow sweep (stream.cuh: gs_dataflow_kernel)
  int df_lanes = 1, df_threads = 128, ntasks = 0;
  int4* tasks = nullptr;
  unsigned* counters = nullptr;   // [0] ticket, [(1 + w) * kGsCounterStride] finished tasks of wavefront w
  void upload(const std::vector<int>& fwd_lvlptr, bool backward_, double mean_row, int lanes) {
    backward = backward_ ? 1 : 0;
    nlev = (int)fwd_lvlptr.size() - 1;
    n = nlev > 0 ? fwd_lvlptr[nlev] : 0;
    // wavefronts in sweep order, as (begin, end) row ranges
    std::vector<std::pair<int, int>> wave(nlev);
    for (int w = 0; w < nlev; ++w) {
      const int src = backward ? nlev - 1 - w : w;
      wave[w] = {fwd_lvlptr[src], fwd_lvlptr[src + 1]};
    }
    std::vector<int> h_rows((size_t)n);
    h_lvlptr.assign(nlev + 1, 0);
    {
      size_t o = 0;
      for (int w = 0; w < nlev; ++w) {
        for (int r = wave[w].first; r < wave[w].second; ++r) h_rows[o++] = r;
        h_lvlptr[w + 1] = (int)o;
      }
    }
    rows = dev_upload(h_rows);
    lvlptr = dev_upload(h_lvlptr);
    // group runs of narrow wavefronts into single-CTA items (fallback mode)
    const int narrow = 4 * (kCtaThreads / lanes);  // <= 4 passes of one CTA
    int l = 0;
    while (l < nlev) {
      const int cnt = h_lvlptr[l + 1] - h_lvlptr[l];
      if (cnt <= narrow) {
        int e = l + 1;
        while (e < nlev && h_lvlptr[e + 1] - h_lvlptr[e] <= narrow) ++e;
        items.push_back({l, e, true});
        l = e;
      } else {
        items.push_back({l, l + 1, false});
        ++l;
      }
    }
    // ---- dataflow tasks ----
    df_lanes = 1;
    while (df_lanes < 32 && kGsPrefetch * df_lanes < (mean_row <= kGsPrefetch ? mean_row : 1.25 * mean_row)) df_lanes *= 2;
    df_lanes = env_int("B200AMG_GS_LANES", df_lanes);
    df_threads = env_int("B200AMG_GS_THREADS", 128) == 256 ? 256 : 128;
    const int R = df_threads / df_lanes;
    std::vector<int4> tk;
    int prev = 0;
    for (int w = 0; w < nlev; ++w) {
      int cnt = 0;
      for (int p = wave[w].first; p < wave[w].second; p += R, ++cnt)
        tk.push_back(make_int4(p, std::min(R, wave[w].second - p), w, prev));
      prev = cnt;
    }
    ntasks = (int)tk.size();
    tasks = dev_upload(tk);
    counters = dev_alloc<unsigned>((int64_t)(nlev + 2) * kGsCounterStride);
    built = true;
  }
  void release() {
    cudaFree(rows); cudaFree(lvlptr); cudaFree(tasks); cudaFree(counters);
    rows = lvlptr = nullptr; tasks = nullptr; counters = nullptr;
    built = false;
  }
};

struct SmootherCfg {
  int kind = 0, sweep = 3, iter = 1;
  double omega = 1.0;
};
static SmootherCfg to_cfg(const b200amg_smoother_t* s) {
  SmootherCfg c;
  if (!s) { c.kind = 0; return c; }
  REQUIRE(s->kind >= 0 && s->kind <= 3, B200AMG_ERR_BAD_ARG, "unknown smoother kind %d", s->kind);
  c.kind = s->kind; c.sweep = s->sweep; c.iter = s->iter; c.omega = s->omega;
  if (c.kind == B200AMG_SMOOTHER_GS || c.kind == B200AMG_SMOOTHER_SOR)
    REQUIRE(c.sweep >= 1 && c.sweep <= 3, B200AMG_ERR_BAD_ARG, "unknown sweep %d", c.sweep);
  REQUIRE(c.iter >= 0, B200AMG_ERR_BAD_ARG, "negative iteration count");
  return c;
}

// Device copy of the blocked-sweep plan (block_plan.h / block_gs.cuh)
struct DevBlockPlan {
  bool ok = false;
  int ntiles = 0, nstages = 0, lanes = 1, wavefronts = 0;
  int4 *tile = nullptr, *stage_meta = nullptr, *stage_aux = nullptr;
  int2 *stage_auxb = nullptr, *req_fwd = nullptr, *req_bwd = nullptr;
  int* steps = nullptr;
  int *order_fwd = nullptr, *order_bwd = nullptr;
  int *code_fwd = nullptr, *code_bwd = nullptr, *dpos = nullptr;   // per-entry codes of the walked matrix (build_block_codes)
  unsigned* ctl = nullptr;   // [0] ticket, [kBgCtlProgress + t] published stages of tile t
  size_t ctl_words = 0;
  void upload(const BlockPlan& P) {
    static_assert(sizeof(BI4) == sizeof(int4) && sizeof(BI2) == sizeof(int2), "plan records are uploaded as int4 / int2");
    ntiles = P.ntiles; nstages = P.nstages; lanes = P.lanes; wavefronts = P.global_wavefronts;
    auto up4 = [](const std::vector<BI4>& v) {
      int4* p = dev_alloc<int4>((int64_t)v.size() + 2);
      if (!v.empty()) CUDA_OK(cudaMemcpy(p, v.data(), sizeof(int4) * v.size(), cudaMemcpyHostToDevice));
      return p;
    };
    auto up2 = [](const std::vector<BI2>& v) {
      int2* p = dev_alloc<int2>((int64_t)v.size() + 2);
      if (!v.empty()) CUDA_OK(cudaMemcpy(p, v.data(), sizeof(int2) * v.size(), cudaMemcpyHostToDevice));
      return p;
    };
    tile = up4(P.tile); stage_meta = up4(P.stage_meta); stage_aux = up4(P.stage_aux);
    stage_auxb = up2(P.stage_auxb); req_fwd = up2(P.req_fwd); req_bwd = up2(P.req_bwd);
    steps = dev_upload(P.steps, 8);
    order_fwd = dev_upload(P.order_fwd, 8);
    order_bwd = dev_upload(P.order_bwd, 8);
    ctl_words = (size_t)kBgCtlProgress + (size_t)ntiles + 8;
    ctl = dev_alloc<unsigned>((int64_t)ctl_words);
    CUDA_OK(cudaMemset(ctl, 0, sizeof(unsigned) * ctl_words));
    ok = true;
  }
  void release() {
    cudaFree(tile); cudaFree(stage_meta); cudaFree(stage_aux); cudaFree(stage_auxb); cudaFree(req_fwd); cudaFree(req_bwd);
    cudaFree(steps); cudaFree(ctl); cudaFree(order_fwd); cudaFree(order_bwd); cudaFree(code_fwd); cudaFree(code_bwd); cudaFree(dpos);
    order_fwd = order_bwd = code_fwd = code_bwd = dpos = nullptr;
    tile = stage_meta = stage_aux = nullptr; stage_auxb = req_fwd = req_bwd = nullptr; steps = nullptr; ctl = nullptr;
    ok = false;
  }
};
// layout of the pass sweep (pass_plan.h / pass_gs.cuh) on the device: slabs of values and per-direction codes, pass / chunk /
// tile records; tiles, stages, requirements, ticket order and progress counters are the blocked plan's (DevBlockPlan)
struct DevPassPlan {
  bool ok = false;
  int lanes = 1;
  int64_t npasses = 0;
  struct Dir {
    int4* pass = nullptr;
    int2 *tile = nullptr, *preq = nullptr, *req = nullptr;
    double* val = nullptr;
    int* idx = nullptr;
  } dir[2];
  void upload(const PassPlan& Q) {
    lanes = Q.lanes; npasses = Q.npasses;
    auto up4 = [](const std::vector<BI4>& v) {
      int4* p = dev_alloc<int4>((int64_t)v.size() + 2);
      if (!v.empty()) CUDA_OK(cudaMemcpy(p, v.data(), sizeof(int4) * v.size(), cudaMemcpyHostToDevice));
      return p;
    };
    auto up2 = [](const std::vector<BI2>& v) {
      int2* p = dev_alloc<int2>((int64_t)v.size() + 34);
      if (!v.empty()) CUDA_OK(cudaMemcpy(p, v.data(), sizeof(int2) * v.size(), cudaMemcpyHostToDevice));
      return p;
    };
    for (int d = 0; d < 2; ++d) {
      const PassDir& D = Q.dir[d];
      dir[d].tile = up2(D.tile); dir[d].pass = up4(D.pass);
      dir[d].preq = up2(D.preq); dir[d].req = up2(D.req);
      dir[d].val = dev_upload(D.val, 8);
      dir[d].idx = dev_upload(D.idx, 8);
    }
    ok = true;
  }
  void release() {
    for (int d = 0; d < 2; ++d) {
      cudaFree(dir[d].tile); cudaFree(dir[d].pass); cudaFree(dir[d].preq); cudaFree(dir[d].req);
      cudaFree(dir[d].val); cudaFree(dir[d].idx);
      dir[d] = Dir();
    }
    ok = false;
  }
};
// A matrix prepared for relaxation: the rows the smoother walks + wavefront schedules + diagonal.
// When a Gauss-Seidel / SOR sweep is requested the level is renumbered into wavefront order (perm).
struct SmootherMatrix {
  DevCsr A;      // true A by rows
  DevCsr At;     // rows of A' (== the reference's CSC columns); aliases A when A is bit-symmetric
  bool symmetric_bits = false;
  int symmetry = B200AMG_SYMMETRY_HERMITIAN;
  DevSchedule fwd, bwd;
  double* diag = nullptr;   // diagonal of the walked matrix (same for A and A')
  int64_t n = 0;
  HostPerm perm;            // identity unless a sweep smoother renumbered the level
  int *d_new_of_old = nullptr, *d_old_of_new = nullptr;
  // mailbox sweep (stream.cuh: gs_mail_kernel): only for structurally symmetric patterns
  // wavefront-aligned tile plan of the walked matrix (stream.cuh: gs_tile_kernel)
  int4* gs_meta = nullptr;
  int* gs_tile_wave = nullptr;
  int gs_ntiles = 0, gs_lanes = 1;
  mutable int gs_tile_ctas = 0;   // persistent CTAs of gs_tile_kernel chosen by tune_tile_ctas (0: all that fit)
  int* d_fwd_lvlptr = nullptr;   // forward wavefront boundaries (single-CTA sweep)
  int nlev = 0;
  bool pattern_symmetric = false;
  // one-cluster sweep with x in distributed shared memory (dsm_gs.cuh): wavefront-aligned tiles of <= 256/T rows
  int4* dsm_meta = nullptr;
  int2* dsm_aux = nullptr;
  int *dsm_code = nullptr, *dsm_rowof = nullptr, *dsm_own_off = nullptr, *dsm_wave_tiles = nullptr;
  int dsm_ntiles = 0, dsm_lanes = 0, dsm_threads = 256, dsm_log_nc = 0, dsm_slots_max = 0;
  int* dsm_status = nullptr;
  uint4* mail = nullptr;
  unsigned* mail_ctl = nullptr;
  DevBlockPlan block;       // blocked sweep (block_gs.cuh): the default for structurally symmetric patterns
  DevPassPlan pass;         // pass sweep (pass_gs.cuh) on the same plan
  const DevCsr& walked() const { return symmetry == B200AMG_SYMMETRY_HERMITIAN ? At : A; }

  // hAt_in: rows of A' (the staged CSC)
  void build(const HostCsr& hAt_in, int symmetry_, bool need_fwd, bool need_bwd, bool need_true_A) {
    symmetry = symmetry_;
    n = hAt_in.nrows;
    int sym_kind;
    { UploadTimer t("symmetry check"); sym_kind = symmetry_kind(hAt_in); }
    symmetric_bits = sym_kind == 2;
    pattern_symmetric = sym_kind >= 1;
    HostCsr hA_own;                     // the true A by rows: only materialised when it differs from A'
    if (!symmetric_bits) hA_own = transpose(hAt_in);
    const HostCsr& hA_in = symmetric_bits ? hAt_in : hA_own;
    std::vector<int> lvlptr;
    HostCsr hAt_p, hA_p;
    const HostCsr* hAt = &hAt_in;
    const HostCsr* hA = &hA_in;
    bool blocked = false;
    // Which exact-order sweep.  Measured on B200 (256^3 RS hierarchy, SGS ms, blocked vs wavefront kernels; profiles/
    // r02_gs_block_vs_wavefront_256.log): stencil-like rows (7 entries, one lane per row) 3.54 vs 3.78, tiny levels (<= ~1000
    // rows) 0.19 / 0.088 vs 0.22 / 0.093; on the irregular coarse levels in between (19-124 entries per row) the blocked
    // sweep's per-stage pipeline latency loses (10.3 / 8.7 / 7.2 / 9.0 / 2.6 vs 7.8 / 6.3 / 5.0 / 6.2 / 2.2).
    // B200AMG_GS_BLOCK: 0 never, 1 (default) by that rule, 2 always.
    // B200AMG_GS_MULTICOLOR=1 (NOT parity: the sweep relaxes colour after colour instead of in index order): wavefront kernels
    // on a greedy colouring, see greedy_colours().
    const bool multicolor = env_int("B200AMG_GS_MULTICOLOR", 0) != 0;
    const int block_mode = multicolor ? 0 : env_int("B200AMG_GS_BLOCK", 1);
    const double mean_row = n ? (double)hAt_in.nnz() / (double)n : 0.0;
    // (tiny levels with longer rows: the two-group one-CTA sweep gs_dsm2_kernel is ahead of the blocked sweep — 800 / 181 / 51
    // rows: 0.34 / 0.12 / 0.058 ms against 0.47 / 0.15 / 0.074 — so they only go to the blocked sweep when it is switched off)
    const bool tiny_blocked = n <= 1024 && env_int("B200AMG_GS_DSM2", 1) == 0;
    const bool block_wanted = block_mode >= 2 || (block_mode == 1 && (mean_row <= 8.0 || tiny_blocked));
    if ((need_fwd || need_bwd) && n > 0 && pattern_symmetric && block_wanted) {
      // blocked sweep: tiles of rows relaxed by one CTA each, rows renumbered (tile, local step, old index)
      const HostCsr& w0 = symmetry == B200AMG_SYMMETRY_HERMITIAN ? hAt_in : hA_in;
      BlockPlan plan;
      { UploadTimer t("block plan"); plan = build_block_plan(w0, block_params_from_env()); }
      if (plan.ok) {
        UploadTimer t_perm("renumbering + permute");
        perm = std::move(plan.perm);
        hAt_p = permute_sym(hAt_in, perm);
        hAt = &hAt_p;
        if (!symmetric_bits) { hA_p = permute_sym(hA_in, perm); hA = &hA_p; } else hA = &hAt_p;
        d_new_of_old = dev_upload(perm.new_of_old);
        d_old_of_new = dev_upload(perm.old_of_new);
        block.upload(plan);
        if (env_int("B200AMG_GS_PASS", 0) >= 1) {   // the pass sweep (pass_gs.cuh) on this plan
          UploadTimer t_pass("pass slabs");
          PassPlan Q = build_pass_plan(plan, symmetry == B200AMG_SYMMETRY_HERMITIAN ? *hAt : *hA, kPgWinOff, kPgZeroOff);
          if (Q.ok) pass.upload(Q);
          if (env_int("B200AMG_BLOCK_VERBOSE", 0))
            fprintf(stderr, "[b200amg] pass plan: %s lanes=%d passes=%lld slab entries=%lld (%.2f x nnz) requirements=%lld\n",
                    Q.ok ? "ok" : Q.why.c_str(), Q.lanes, (long long)Q.npasses, (long long)Q.dir[0].nentries,
                    (double)Q.dir[0].nentries / (double)std::max<int64_t>(1, plan.nnz), (long long)Q.dir[0].req.size());
        }
        if (!pass.ok) {
          UploadTimer t_codes("block entry codes");
          hvec<int> cf, cb, dp;
          build_block_codes(plan, symmetry == B200AMG_SYMMETRY_HERMITIAN ? *hAt : *hA, cf, cb, dp);
          block.code_fwd = dev_upload(cf, 8);
          block.code_bwd = dev_upload(cb, 8);
          block.dpos = dev_upload(dp, 8);
        }
        nlev = plan.global_wavefronts;
        blocked = true;
        if (env_int("B200AMG_BLOCK_VERBOSE", 0))
          fprintf(stderr, "[b200amg] block plan: n=%lld nnz=%lld wavefronts=%d lanes=%d tiles=%d stages=%d steps=%d rows/step %.1f (target %.1f) theta=%.0f a=%d b=%d max tile rows %lld steps %d\n",
                  (long long)n, (long long)plan.nnz, plan.global_wavefronts, plan.lanes, plan.ntiles, plan.nstages, plan.nsteps,
                  plan.mean_step_rows, plan.target_step_rows, plan.theta, plan.block_a, plan.block_b, (long long)plan.max_tile_rows,
                  plan.max_tile_steps);
      } else if (env_int("B200AMG_BLOCK_VERBOSE", 0)) {
        fprintf(stderr, "[b200amg] block plan rejected (%s): wavefront sweeps\n", plan.why.c_str());
      }
    }
    if ((need_fwd || need_bwd) && n > 0 && !blocked) {
      const HostCsr& w0 = symmetry == B200AMG_SYMMETRY_HERMITIAN ? hAt_in : hA_in;
      const HostCsr& wt0 = symmetric_bits ? w0 : (symmetry == B200AMG_SYMMETRY_HERMITIAN ? hA_in : hAt_in);
      int nlev = 0;
      std::vector<int> level;
      { UploadTimer t("wavefront levels"); level = multicolor ? greedy_colours(w0, wt0, &nlev) : wavefront_levels(w0, wt0, &nlev); }
      UploadTimer t_perm("renumbering + permute");
      lvlptr.assign(nlev + 1, 0);
      for (int64_t i = 0; i < n; ++i) lvlptr[level[i] + 1]++;
      for (int l = 0; l < nlev; ++l) lvlptr[l + 1] += lvlptr[l];
      perm.old_of_new.resize(n);
      perm.new_of_old.resize(n);
      std::vector<int> next(lvlptr.begin(), lvlptr.end() - 1);
      for (int64_t i = 0; i < n; ++i) {   // ascending old index inside a wavefront
        const int q = next[level[i]]++;
        perm.old_of_new[q] = (int)i;
        perm.new_of_old[i] = q;
      }
      hAt_p = permute_sym(hAt_in, perm);
      hAt = &hAt_p;
      if (!symmetric_bits) { hA_p = permute_sym(hA_in, perm); hA = &hA_p; } else hA = &hAt_p;
      d_new_of_old = dev_upload(perm.new_of_old);
      d_old_of_new = dev_upload(perm.old_of_new);
    }
    { UploadTimer t("operator to device + tiles"); At.upload(*hAt); }
    if (symmetric_bits) A.alias(At);
    else if (need_true_A || symmetry == B200AMG_SYMMETRY_NONE) A.upload(*hA);
    const HostCsr& w = symmetry == B200AMG_SYMMETRY_HERMITIAN ? *hAt : *hA;
    std::vector<double> d(n, 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
      for (int k = w.ptr[i]; k < w.ptr[i + 1]; ++k)
        if (w.idx[k] == i) d[i] = w.val[k];
    diag = dev_upload(d, 8);
    if (symmetry == B200AMG_SYMMETRY_NONE && (need_fwd || need_bwd)) {
      // DiagonalIndices(A): SingularException on a missing / zero diagonal  (smoother.jl:233-248)
      int64_t bad = -1;   // the reference reports the first (lowest) column without a usable diagonal
      for (int64_t i = 0; i < n; ++i)
        if (d[i] == 0.0) {
          const int64_t old = perm.identity() ? i : perm.old_of_new[i];
          if (bad < 0 || old < bad) bad = old;
        }
      REQUIRE(bad < 0, B200AMG_ERR_SINGULAR, "SingularException(%lld)", (long long)(bad + 1));
    }
    const double mean = n ? (double)w.nnz() / (double)n : 0.0;
    if (blocked) {   // the wavefront schedules of the other sweep kernels do not exist in this numbering
      fwd.backward = 0; bwd.backward = 1;
      fwd.nlev = bwd.nlev = nlev;
      fwd.n = bwd.n = n;
      return;
    }
    { UploadTimer t("sweep schedules"); if (need_fwd) fwd.upload(lvlptr, false, mean, walked().lanes);
    if (need_bwd) bwd.upload(lvlptr, true, mean, walked().lanes); }
    UploadTimer t_plans("sweep tile plans (dsm / tile / mailboxes)");
    if ((need_fwd || need_bwd) && n > 0) {
      d_fwd_lvlptr = dev_upload(lvlptr, 8);
      nlev = (int)lvlptr.size() - 1;
    }
    if ((need_fwd || need_bwd) && n > 0 && n <= (int64_t)16 * 28000 && mean >= 6.0 && lvlptr.size() >= 2) {
      // plan of the distributed-shared-memory sweep (dsm_gs.cuh): tiles never cross a wavefront, <= 256/T rows,
      // <= kDsmTileNnz entries; tile t belongs to CTA t % NC, which also keeps the x of the tile's rows
      int T = 4;
      while (T < 32 && kDsmBurst * T < 1.6 * mean) T *= 2;   // one gather burst covers all but the longest rows
      T = std::min(32, std::max(4, env_int("B200AMG_DSM_LANES", T)));
      const int threads = env_int("B200AMG_DSM_THREADS", kDsmThreads) == 512 ? 512 : 256;
      const int G = threads / T;
      std::vector<int4> tm;
      std::vector<int2> ta;
      bool ok = true;
      const int nl = (int)lvlptr.size() - 1;
      std::vector<int> wave_tiles((size_t)nl, 0);
      for (int wv = 0; wv < nl && ok; ++wv) {
        int r = lvlptr[wv];
        while (r < lvlptr[wv + 1]) {
          int e2 = r;
          const int k0 = w.ptr[r];
          while (e2 < lvlptr[wv + 1] && e2 - r < G && w.ptr[e2 + 1] - k0 <= kDsmTileNnz) ++e2;
          if (e2 == r) { ok = false; break; }   // a row longer than a tile
          tm.push_back(make_int4(r, e2, k0, w.ptr[e2]));
          ta.push_back(make_int2(wv, 0));
          ++wave_tiles[wv];
          r = e2;
        }
      }
      if (ok) {
        auto slots_max_for = [&](int lnc) {
          std::vector<int64_t> cnt((size_t)1 << lnc, 0);
          for (size_t t = 0; t < tm.size(); ++t) cnt[t & ((1u << lnc) - 1)] += tm[t].y - tm[t].x;
          return *std::max_element(cnt.begin(), cnt.end());
        };
        int lnc = 0;
        while (lnc <= 4 && dsm_smem_bytes(slots_max_for(lnc), nl) > (size_t)kDsmMaxDynSmem) ++lnc;
        const int lnc_fit = lnc;
        const int lnc_max = std::min(4, std::max(0, env_int("B200AMG_GS_DSM_MAX_LOG_NC", 4)));
        const double wave_rows = (double)n / (double)nl;
        while (lnc < lnc_max && wave_rows > (double)G * (double)(1 << lnc)) ++lnc;   // one pass of all CTAs covers a mean wavefront
        const int forced = env_int("B200AMG_GS_DSM_LOG_NC", -1);
        if (forced >= 0) lnc = std::min(4, std::max(lnc_fit, forced));
        if (lnc <= 4) {
          const int NC = 1 << lnc;
          std::vector<int> running(NC, 0), code_of_row((size_t)n, 0);
          std::vector<std::vector<int>> rows_of(NC);
          for (size_t t = 0; t < tm.size(); ++t) {
            const int owner = (int)(t & (size_t)(NC - 1));
            ta[t].y = running[owner];
            for (int r = tm[t].x; r < tm[t].y; ++r) {
              code_of_row[r] = (running[owner] << lnc) | owner;
              rows_of[owner].push_back(r);
              ++running[owner];
            }
          }
          std::vector<int> own_off(NC + 2, 0), rowof;
          rowof.reserve((size_t)n);
          for (int c = 0; c < NC; ++c) {
            own_off[c + 1] = own_off[c] + running[c];
            rowof.insert(rowof.end(), rows_of[c].begin(), rows_of[c].end());
          }
          own_off[NC + 1] = *std::max_element(running.begin(), running.end());
          std::vector<int> code(w.idx.size());
          for (size_t k = 0; k < w.idx.size(); ++k) code[k] = code_of_row[w.idx[k]];
          dsm_meta = dev_upload(tm);
          dsm_aux = dev_upload(ta);
          dsm_code = dev_upload(code, 8);
          dsm_rowof = dev_upload(rowof, 8);
          dsm_own_off = dev_upload(own_off);
          dsm_wave_tiles = dev_upload(wave_tiles);
          dsm_ntiles = (int)tm.size();
          dsm_lanes = T;
          dsm_threads = threads;
          dsm_log_nc = lnc;
          dsm_slots_max = own_off[NC + 1];
          dsm_status = dev_alloc<int>(4);
          CUDA_OK(cudaMemset(dsm_status, 0, 4 * sizeof(int)));
          if (env_int("B200AMG_GS_DSM_VERBOSE", 0))
            fprintf(stderr, "[b200amg] dsm plan: n=%lld nnz=%lld wavefronts=%d lanes=%d threads=%d ctas=%d tiles=%d slots/cta=%d smem=%zu\n",
                    (long long)n, (long long)w.nnz(), nl, T, threads, NC, dsm_ntiles, dsm_slots_max, dsm_smem_bytes(dsm_slots_max, nl));
        }
      }
    }
    if ((need_fwd || need_bwd) && pattern_symmetric && n > 0) {
      gs_lanes = 1;
      while (gs_lanes < 32 && kGsPrefetch * gs_lanes < (mean <= kGsPrefetch ? mean : 1.25 * mean)) gs_lanes *= 2;
      gs_lanes = env_int("B200AMG_GS_LANES", gs_lanes);
      const int G = kGsTileThreads / gs_lanes;
      const int rows_per_tile = G * std::min(std::max(env_int("B200AMG_GS_TILE_PASSES", 1), 1), 4);
      std::vector<int4> tm;
      std::vector<int> tw;
      bool ok = true;
      for (int wv = 0; wv + 1 < (int)lvlptr.size() && ok; ++wv) {
        int r = lvlptr[wv];
        while (r < lvlptr[wv + 1]) {
          int e2 = r;
          const int k0 = w.ptr[r];
          while (e2 < lvlptr[wv + 1] && e2 - r < rows_per_tile && w.ptr[e2 + 1] - k0 <= kTileNnz) ++e2;
          if (e2 == r) { ok = false; break; }   // a row longer than a tile: the other sweeps take over
          tm.push_back(make_int4(r, e2, k0, w.ptr[e2]));
          tw.push_back(wv);
          r = e2;
        }
      }
      if (ok) {
        gs_meta = dev_upload(tm);
        gs_tile_wave = dev_upload(tw);
        gs_ntiles = (int)tm.size();
      }
      mail = dev_alloc<uint4>(n + 8);
      CUDA_OK(cudaMemset(mail, 0, sizeof(uint4) * (size_t)(n + 8)));
      const int64_t words = (int64_t)(lvlptr.size() + 4) * kGsCounterStride;
      mail_ctl = dev_alloc<unsigned>(words);
      CUDA_OK(cudaMemset(mail_ctl, 0, sizeof(unsigned) * (size_t)words));
    }
  }
  void release() {
    A.release(); At.release(); fwd.release(); bwd.release(); block.release(); pass.release();
    cudaFree(diag); cudaFree(d_new_of_old); cudaFree(d_old_of_new); cudaFree(mail); cudaFree(mail_ctl); cudaFree(d_fwd_lvlptr); cudaFree(gs_meta); cudaFree(gs_tile_wave);
    cudaFree(dsm_meta); cudaFree(dsm_aux); cudaFree(dsm_status); cudaFree(dsm_code); cudaFree(dsm_rowof); cudaFree(dsm_own_off); cudaFree(dsm_wave_tiles);
    dsm_meta = nullptr; dsm_aux = nullptr; dsm_status = nullptr; dsm_code = dsm_rowof = dsm_own_off = dsm_wave_tiles = nullptr; dsm_ntiles = 0;
    d_fwd_lvlptr = nullptr; gs_meta = nullptr; gs_tile_wave = nullptr; gs_ntiles = 0;
    diag = nullptr; d_new_of_old = d_old_of_new = nullptr; mail = nullptr; mail_ctl = nullptr;
  }
};

static bool cfg_needs_fwd(const SmootherCfg& c) {
  return (c.kind == B200AMG_SMOOTHER_GS || c.kind == B200AMG_SMOOTHER_SOR) && (c.sweep == 1 || c.sweep == 3);
}
static bool cfg_needs_bwd(const SmootherCfg& c) {
  return (c.kind == B200AMG_SMOOTHER_GS || c.kind == B200AMG_SMOOTHER_SOR) && (c.sweep == 2 || c.sweep == 3);
}

struct Level {
  int64_t n = 0, nc = 0;
  int64_t nnz_a = 0, nnz_p = 0;   // kept for level_info (a partitioned / remote level has no full device copy)
  bool remote = false;            // this rank holds no device data for the level (rank != 0 of a partition)
  SmootherMatrix M;
  DevCsr P, R;
  // P and R wait on the host until the NEXT level's numbering is known (add_level / set_coarse)
  HostCsr pendP, pendR;
  bool pending = false;
  SmootherCfg pre, post;
  double *res = nullptr, *coarse_x = nullptr, *coarse_b = nullptr, *temp = nullptr;
};

// The fine level of a row-partitioned hierarchy as one rank sees it (partition.h has the plan).
struct Part {
  PartPlan plan;
  int64_t n = 0, nc = 0;
  DevCsr A, At, R, P;            // local blocks; At aliases A when A is bit-symmetric
  int symmetry = B200AMG_SYMMETRY_HERMITIAN;
  SmootherCfg pre, post;
  double* diag = nullptr;        // diagonal of the owned rows
  int* send_idx = nullptr;
  double* sendbuf = nullptr;
  double *x = nullptr, *b = nullptr, *res = nullptr, *temp = nullptr;   // [owned | halo]
  double *cb = nullptr, *cx = nullptr;   // my coarse_b rows / my coarse_x window (alias the full vectors on rank 0)
  bool own_cb = false, own_cx = false;
  double* xfull = nullptr;       // staging for the final all-gather when the caller's x is host memory
  // The level below may be partitioned too (B200AMG_OPT_PART_LEVELS): then cb / cx ARE the child's b / x
  // ([owned | halo]) and P's columns are the child's local ids.  P is therefore uploaded only when the next
  // add_level / set_coarse call tells what the level below looks like.
  int level = 0;
  Part* child = nullptr;
  HostCsr pendP;
  bool pendingP = false;
  const DevCsr& walked() const { return symmetry == B200AMG_SYMMETRY_HERMITIAN ? At : A; }
  void release() {
    A.release(); At.release(); R.release(); P.release();
    cudaFree(diag); cudaFree(send_idx); cudaFree(sendbuf); cudaFree(x); cudaFree(b); cudaFree(res); cudaFree(temp);
    if (own_cb) cudaFree(cb);
    if (own_cx) cudaFree(cx);
    cudaFree(xfull);
  }
};

struct b200amg_hierarchy {
  int device = 0;
  // row partition of the fine level (world == 1: none)
  int rank = 0, world = 1;
  ncclUniqueId nccl_id;
  ncclComm_t comm = nullptr;
  std::vector<std::unique_ptr<Part>> parts;   // partitioned levels 0 .. parts.size()-1
  Part* part = nullptr;                       // parts[0]: what solve / cycle / precond load and store
  int part_levels = 1;                        // how many of the finest levels are partitioned (world > 1)
  cudaStream_t stream = nullptr;
  int num_sms = kNumSM;                       // queried at create
  std::vector<std::unique_ptr<Level>> levels;
  // coarsest
  bool have_coarse = false;
  int64_t nfinal = 0;
  DevCsr finalA;
  double* coarse_inv = nullptr;
  double* res_final = nullptr;
  // coarse solver as a host callable (b200amg_set_coarse_callback): pinned staging vectors, the callable, its last status
  b200amg_coarse_fn coarse_fn = nullptr;
  void* coarse_user = nullptr;
  double *coarse_hb = nullptr, *coarse_hx = nullptr;
  volatile int32_t coarse_fn_status = 0;
  int64_t coarse_fn_calls = 0;
  // level-0 work vectors
  int64_t n0 = 0;
  double *x0 = nullptr, *b0 = nullptr;
  // reductions
  double* partial = nullptr;
  double* scalars = nullptr;  // device scalars: [0] norm, [1] rho, [2] rho_prev, [3] uq, [4] scratch
  double* h_scalars = nullptr;  // pinned
  // PCG
  double *pcg_u = nullptr, *pcg_q = nullptr, *pcg_x = nullptr;
  // graphs
  cudaGraphExec_t cycle_graph[3] = {nullptr, nullptr, nullptr};
  int64_t cycle_graph_launches[3] = {0, 0, 0};
  cudaGraphExec_t resnorm_graph = nullptr;
  bool use_graphs = true;
  bool part_graphs = true;   // partitioned handles: rank 0 replays the levels below the fine one as a graph
  bool part_overlap = true;   // halo exchange on a second stream, overlapped with the interior rows of the kernel that needs it
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  bool part_whole_graph = true;   // partitioned handles: the whole cycle (kernels + NCCL groups) is one captured graph per rank
  cudaGraphExec_t part_cycle_graph[3] = {nullptr, nullptr, nullptr};
  int64_t part_cycle_launches[3] = {0, 0, 0}, part_cycle_collectives[3] = {0, 0, 0};
  int stream_chunk = 4;   // consecutive tiles per CTA run of the stream kernels (0: contiguous split)
  int64_t gs_cta_rows = 12288;   // levels up to this many rows are swept by ONE CTA (bar.sync per wavefront, x in smem)
  int gs_cluster = 0;                 // one-cluster sweep (x in distributed shared memory) for mid-size levels: measured
                                      // 2.4-3.9 us per wavefront vs 2.2-2.5 for the counter sweep, so off by default
  int64_t gs_cluster_rows = 380000;
  int gs_cluster_log_nc = 3, gs_cluster_threads = 256;
  int gs_dsm = 1;                     // 1: one-cluster sweep with x in distributed shared memory + per-wavefront mbarriers
                                      // (dsm_gs.cuh) on narrow-wavefront levels that fit gs_dsm_max_log_nc CTAs; 2: required
  int gs_dsm_max_log_nc = 4;          // measured (256^3 RS hierarchy, SGS ms), gs_dsm_kernel: 1 CTA 2.56 -> 1.83 (5 195 rows), 4 CTAs
                                      // 7.42 -> 5.63 (38 260 rows), 16 CTAs 4.69 -> 5.16 (228 538 rows); gs_dsm2_kernel: 1.29 / 4.71 /
                                      // 3.75 -> up to 16 CTAs with the two-group kernel, up to 4 without it
  int gs_dsm2 = 1;                    // 1: gs_dsm2_kernel (two consumer groups alternate the tiles: preparation off the hand-off path)
  int gs_dsm_fence = 0;               // bit 0 / 1: cluster-scope fence on the producer / consumer side of the hand-off
  int gs_counter_mail = 1;            // counter sweep publishes mailboxes instead of fencing (symmetric patterns)
  int gs_tile_any_lanes = 1;          // 1: use the TMA-fed mailbox sweep for multi-lane rows too
  int64_t gs_mail_min_width = 1024;   // mean rows per wavefront from which the mailbox sweep is used
  int gs_poll_masked = -1;            // TMA-fed mailbox sweep: 1 poll only the mailboxes a row still waits for (measured: -2.7 %);
                                      // 2 additionally spin on one outstanding mailbox between rounds; -1 (default): 2 on rows
                                      // of >= 8 lanes, else 1 (measured, see launch_gs_tile_T)
  int gs_gate_dist = 2;               // a tile of wavefront w stays off the mailboxes until wavefront w - gs_gate_dist has begun to finish
  int gs_tile_cta_limit = 0;          // experiment knob (B200AMG_GS_TILE_CTAS): cap on the persistent CTAs of gs_tile_kernel
  int gs_poll_sleep = 0, gs_gate_sleep = 100;   // ns between failed mailbox polls / throttle polls
  int opaque_zero = 0;    // a zero the compiler cannot see (scheduling fence in gs_dataflow_kernel)
  int gs_acquire = 0;     // consumer-side acquire of the dataflow sweep: 0 none (see stream.cuh), 1 ld.acquire, 2 fence
  unsigned long long* gs_debug = nullptr;   // 8 timestamps per task of the last dataflow sweep (diagnostics)
  int* gs_fault = nullptr;                  // set by a sweep kernel whose watchdog fired (checked after every stream sync of an entry point)
  int gs_mode = 2;        // 2: per-row mailbox sweep (symmetric patterns; else 1), 1: wavefront-counter dataflow sweep,
                          // 0: one launch per wavefront (fallback / A-B)
  bool finalized = false;
  bool capturing = false;
  int64_t launches = 0;       // kernels launched (graph replays add their node counts)
  // stream kernels read the binary32 copy of an operator's values where one exists (DevCsr::val32).  OFF by default: measured on
  // B200 (256^3 fine level, profiles/r02_fp32_storage_ab.md) the residual takes 0.392 ms with 4-byte values against 0.347 ms
  // with 8-byte values — the kernel is co-limited by instruction issue, and seven F2F.F64.F32 conversions per row (quarter
  // rate) plus shorter bulk copies cost more than the 25 % fewer bytes save.  B200AMG_FP32_STORAGE=1 / B200AMG_OPT_FP32_STORAGE.
  bool fp32_storage = env_int("B200AMG_FP32_STORAGE", 0) != 0;
  int64_t collectives = 0;    // NCCL groups / collectives enqueued (partitioned handles)
  // halo exchange over peer memory (peer_halo.cuh): on when every rank could map its neighbours' vectors
  struct PeerCtx {
    bool on = false;
    unsigned long long* sync = nullptr;   // my flag / ack / counter words (exported)
    unsigned* tickets = nullptr;          // one per (level, channel)
    PeerTables* d_tab = nullptr;          // [levels * channels]
    std::vector<void*> opened;            // peer mappings to close
  } peer;
  Part* peer_pending = nullptr;           // the exchange whose halo has not been acknowledged yet
  int peer_pending_ch = 0;
  int64_t peer_exchanges = 0;
  int64_t part_cycle_peer[3] = {0, 0, 0};
  int64_t capture_count = 0;  // kernels recorded into the graph being captured
  // staging for renumbered vectors crossing the ABI
  double* io_tmp = nullptr;
  int64_t io_cap = 0;
  // L2 flush buffer for time_kernel
  void* flush = nullptr;
  size_t flush_bytes = 0;
  // profiling
  bool profiling = false;
  std::vector<double>* prof_ms = nullptr;
  // per-iteration timing of the fine-level convergence residual (B200AMG_OPT_TIME_RESIDUAL)
  bool time_residual = false;
  std::vector<cudaEvent_t> res_events;   // 2 per iteration
  int res_events_used = 0;
};
typedef b200amg_hierarchy H;

static void cycle_body_part(H* h, int cycle);
static inline void count_launch(H* h) {
  if (h->capturing) h->capture_count++; else h->launches++;
}
static inline unsigned grid_for(int64_t work_items) {
  return (unsigned)std::max<int64_t>(1, (work_items + kThreads - 1) / kThreads);
}

