// Kernel launchers: stream kernels, the Gauss-Seidel sweep kernels and their per-level selection, smoothers, reductions.
// Part of engine.cu (one translation unit).
#pragma once

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
// ---- TMA stream kernels ---------------------------------------------------------------------
template <int T, int MODE>
static void stream_set_attr() {
  CUDA_OK(cudaFuncSetAttribute(csr_stream_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemBytes));
  CUDA_OK(cudaFuncSetAttribute(csr_stream_kernel<T, MODE, kStreamBurst, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemBytes));
}
template <int MODE>
static void stream_set_attr_all() {
  CUDA_OK(cudaFuncSetAttribute(csr_stream_kernel<2, MODE, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemBytes));
  CUDA_OK(cudaFuncSetAttribute(csr_stream_kernel<4, MODE, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemBytes));
  CUDA_OK(cudaFuncSetAttribute(csr_stream_kernel<2, MODE, 16, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemBytes));
  CUDA_OK(cudaFuncSetAttribute(csr_stream_kernel<4, MODE, 16, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemBytes));
  stream_set_attr<1, MODE>(); stream_set_attr<2, MODE>(); stream_set_attr<4, MODE>();
  stream_set_attr<8, MODE>(); stream_set_attr<16, MODE>(); stream_set_attr<32, MODE>();
}
static void stream_kernels_init() {   // once per device context: opt in to 86 KB of dynamic shared memory
  stream_set_attr_all<0>(); stream_set_attr_all<1>(); stream_set_attr_all<2>(); stream_set_attr_all<3>(); stream_set_attr_all<4>();
}
// part: 0 every tile, 1 the interior tiles, 2 the boundary tiles (row-partitioned levels, DevCsr::meta_split)
template <int MODE, typename VT>
static void launch_stream_vt(H* h, const DevCsr& A, const VT* val, int ctas, int chunk, const double* x, const double* b, double* y,
                             double omega, const double* diagvals) {
#define B200AMG_STREAM_CASE(TT)                                                                                                        \
  case TT:                                                                                                                             \
    csr_stream_kernel<TT, MODE, kStreamBurst, VT><<<ctas, kStreamThreads, kStreamSmemBytes, h->stream>>>(A.ntiles, chunk, A.meta, A.ptr, A.idx, \
                                                                                                        val, x, b, y, omega, diagvals); \
    break;
  if (A.stream_burst == 16) {   // 13-64 entries per row: two / four lanes, one burst of 16 gathers each
    if (A.stream_lanes == 2)
      csr_stream_kernel<2, MODE, 16, VT><<<ctas, kStreamThreads, kStreamSmemBytes, h->stream>>>(A.ntiles, chunk, A.meta, A.ptr, A.idx, val, x, b, y,
                                                                                               omega, diagvals);
    else
      csr_stream_kernel<4, MODE, 16, VT><<<ctas, kStreamThreads, kStreamSmemBytes, h->stream>>>(A.ntiles, chunk, A.meta, A.ptr, A.idx, val, x, b, y,
                                                                                               omega, diagvals);
    count_launch(h);
    return;
  }
  switch (A.stream_lanes) {
    B200AMG_STREAM_CASE(1) B200AMG_STREAM_CASE(2) B200AMG_STREAM_CASE(4) B200AMG_STREAM_CASE(8) B200AMG_STREAM_CASE(16)
    default:
      csr_stream_kernel<32, MODE, kStreamBurst, VT><<<ctas, kStreamThreads, kStreamSmemBytes, h->stream>>>(A.ntiles, chunk, A.meta, A.ptr, A.idx, val,
                                                                                                          x, b, y, omega, diagvals);
  }
#undef B200AMG_STREAM_CASE
  count_launch(h);
}
template <int MODE>
static void launch_stream(H* h, const DevCsr& A0, const double* x, const double* b, double* y, double omega,
                          const double* diagvals, int part = 0) {
  DevCsr A = A0;   // (a shallow view: tile list and count swapped for the requested part)
  A.owner = false;
  if (part == 1) { A.meta = A0.meta_split; A.ntiles = A0.ntiles_int; }
  else if (part == 2) { A.meta = A0.meta_split + A0.ntiles_int; A.ntiles = A0.ntiles_bnd; }
  if (A.ntiles == 0) return;
  const int ctas = std::min(A.ntiles, h->num_sms * 2);
  const int chunk = h->stream_chunk > 0 ? h->stream_chunk : (A.ntiles + ctas - 1) / ctas;
  if (A.val32 && h->fp32_storage) launch_stream_vt<MODE, float>(h, A, A.val32, ctas, chunk, x, b, y, omega, diagvals);
  else launch_stream_vt<MODE, double>(h, A, A.val, ctas, chunk, x, b, y, omega, diagvals);
}

template <int MODE>
static void launch_csr(H* h, const DevCsr& A, const double* x, const double* b, double* y, int part = 0) {
  if (A.nrows == 0) return;
  if (A.ntiles > 0) { launch_stream<MODE>(h, A, x, b, y, 0.0, nullptr, part); return; }
  if (part == 1) return;   // not streamable: everything runs as the "boundary" part, after the exchange
  const unsigned g = grid_for(A.nrows * A.lanes);
  switch (A.lanes) {
    case 2: csr_vec_kernel<2, MODE><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, x, b, y); break;
    case 4: csr_vec_kernel<4, MODE><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, x, b, y); break;
    case 8: csr_vec_kernel<8, MODE><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, x, b, y); break;
    case 16: csr_vec_kernel<16, MODE><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, x, b, y); break;
    default: csr_vec_kernel<32, MODE><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, x, b, y); break;
  }
  count_launch(h);
}
static void spmv(H* h, const DevCsr& A, const double* x, double* y, int part = 0) { launch_csr<0>(h, A, x, nullptr, y, part); }
static void residual(H* h, const DevCsr& A, const double* x, const double* b, double* r, int part = 0) { launch_csr<1>(h, A, x, b, r, part); }
static void spmv_add(H* h, const DevCsr& A, const double* x, double* y, int part = 0) { launch_csr<2>(h, A, x, nullptr, y, part); }

static void launch_jacobi_fast(H* h, const DevCsr& A, const double* xin, const double* b, double* xout, double w, int part = 0) {
  if (A.ntiles > 0) { launch_stream<3>(h, A, xin, b, xout, w, nullptr, part); return; }
  if (part == 1) return;
  const unsigned g = grid_for(A.nrows * A.lanes);
  switch (A.lanes) {
    case 2: jacobi_fast_kernel<2><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, xin, b, xout, w); break;
    case 4: jacobi_fast_kernel<4><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, xin, b, xout, w); break;
    case 8: jacobi_fast_kernel<8><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, xin, b, xout, w); break;
    case 16: jacobi_fast_kernel<16><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, xin, b, xout, w); break;
    default: jacobi_fast_kernel<32><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, xin, b, xout, w); break;
  }
  count_launch(h);
}
static void launch_jacobi_general(H* h, const DevCsr& A, const double* diag, const double* xin, const double* b,
                                  double* xout, double w, int part = 0) {
  if (A.ntiles > 0) { launch_stream<4>(h, A, xin, b, xout, w, diag, part); return; }
  if (part == 1) return;
  const unsigned g = grid_for(A.nrows * A.lanes);
  switch (A.lanes) {
    case 2: jacobi_general_kernel<2><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, diag, xin, b, xout, w); break;
    case 4: jacobi_general_kernel<4><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, diag, xin, b, xout, w); break;
    case 8: jacobi_general_kernel<8><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, diag, xin, b, xout, w); break;
    case 16: jacobi_general_kernel<16><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, diag, xin, b, xout, w); break;
    default: jacobi_general_kernel<32><<<g, kThreads, 0, h->stream>>>(A.nrows, A.ptr, A.idx, A.val, diag, xin, b, xout, w); break;
  }
  count_launch(h);
}

template <int T>
static void launch_sweep_T(H* h, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w, int sor) {
  for (const SweepItem& it : sc.items) {
    if (it.single_cta) {
      gs_cta_levels_kernel<T><<<1, kCtaThreads, 0, h->stream>>>(sc.rows, sc.lvlptr, it.lv_begin, it.lv_end, A.ptr, A.idx,
                                                               A.val, x, b, w, sor);
    } else {
      const int s = sc.h_lvlptr[it.lv_begin], cnt = sc.h_lvlptr[it.lv_begin + 1] - s;
      gs_wavefront_kernel<T><<<grid_for((int64_t)cnt * T), kThreads, 0, h->stream>>>(sc.rows + s, cnt, A.ptr, A.idx, A.val, x,
                                                                                  b, w, sor);
    }
    count_launch(h);
  }
}
template <int T, int BS, bool MAIL>
static int gs_dataflow_ctas() {   // co-resident CTAs of the persistent dataflow sweep
  static int cached = 0;
  if (!cached) {
    int per_sm = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gs_dataflow_kernel<T, BS, MAIL>, BS, 0));
    cached = std::max(1, per_sm);   // per SM: the call sites multiply by the SM count of the handle's device
  }
  return cached;
}
template <int T, int BS>
static void launch_dataflow_T(H* h, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w, int sor,
                              uint4* mail, unsigned* mail_ctl) {
  CUDA_OK(cudaMemsetAsync(sc.counters, 0, sizeof(unsigned) * (size_t)(sc.nlev + 2) * kGsCounterStride, h->stream));
  if (mail && h->gs_counter_mail && !h->gs_debug) {
    const int ctas = std::min(sc.ntasks, gs_dataflow_ctas<T, BS, true>() * h->num_sms);
    gs_mail_prepare_kernel<<<1, 32, 0, h->stream>>>(mail_ctl);   // new epoch for the mailbox flags
    count_launch(h);
    gs_dataflow_kernel<T, BS, true><<<ctas, BS, 0, h->stream>>>(sc.ntasks, sc.tasks, sc.counters, A.ptr, A.idx, A.val, x, b, w, sor,
                                                               sc.backward, h->gs_acquire, h->opaque_zero, nullptr, mail, mail_ctl);
  } else {
    const int ctas = std::min(sc.ntasks, gs_dataflow_ctas<T, BS, false>() * h->num_sms);
    gs_dataflow_kernel<T, BS, false><<<ctas, BS, 0, h->stream>>>(sc.ntasks, sc.tasks, sc.counters, A.ptr, A.idx, A.val, x, b, w, sor,
                                                                sc.backward, h->gs_acquire, h->opaque_zero, h->gs_debug, nullptr, nullptr);
  }
  count_launch(h);
}
static void launch_dataflow(H* h, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w, int sor,
                            uint4* mail = nullptr, unsigned* mail_ctl = nullptr) {
  if (sc.ntasks == 0) return;
#define B200AMG_DF_CASE(TT)                                                    \
  case TT:                                                                     \
    if (sc.df_threads == 128) launch_dataflow_T<TT, 128>(h, A, sc, x, b, w, sor, mail, mail_ctl); \
    else launch_dataflow_T<TT, 256>(h, A, sc, x, b, w, sor, mail, mail_ctl);                      \
    break;
  switch (sc.df_lanes) {
    B200AMG_DF_CASE(1) B200AMG_DF_CASE(2) B200AMG_DF_CASE(4) B200AMG_DF_CASE(8) B200AMG_DF_CASE(16)
    default:
      if (sc.df_threads == 128) launch_dataflow_T<32, 128>(h, A, sc, x, b, w, sor, mail, mail_ctl);
      else launch_dataflow_T<32, 256>(h, A, sc, x, b, w, sor, mail, mail_ctl);
  }
#undef B200AMG_DF_CASE
}
template <int T, int BS>
static int gs_mail_ctas() {
  static int cached = 0;
  if (!cached) {
    int per_sm = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gs_mail_kernel<T, BS>, BS, 0));
    cached = std::max(1, per_sm);   // per SM: the call sites multiply by the SM count of the handle's device
  }
  return cached;
}
template <int T, int BS>
static void launch_mail_T(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w,
                          int sor) {
  const int ctas = std::min(sc.ntasks, gs_mail_ctas<T, BS>() * h->num_sms);
  gs_mail_prepare_kernel<<<1, 32, 0, h->stream>>>(M.mail_ctl);
  count_launch(h);
  gs_mail_kernel<T, BS><<<ctas, BS, 0, h->stream>>>(sc.ntasks, sc.tasks, M.mail_ctl, A.ptr, A.idx, A.val, x, b, M.mail, w, sor,
                                                   sc.backward, h->opaque_zero, h->gs_poll_sleep, h->gs_gate_sleep);
  count_launch(h);
}
static void launch_mail(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w,
                        int sor) {
  if (sc.ntasks == 0) return;
#define B200AMG_ML_CASE(TT)                                                       \
  case TT:                                                                        \
    if (sc.df_threads == 128) launch_mail_T<TT, 128>(h, M, A, sc, x, b, w, sor);  \
    else launch_mail_T<TT, 256>(h, M, A, sc, x, b, w, sor);                       \
    break;
  switch (sc.df_lanes) {
    B200AMG_ML_CASE(1) B200AMG_ML_CASE(2) B200AMG_ML_CASE(4) B200AMG_ML_CASE(8) B200AMG_ML_CASE(16)
    default:
      if (sc.df_threads == 128) launch_mail_T<32, 128>(h, M, A, sc, x, b, w, sor);
      else launch_mail_T<32, 256>(h, M, A, sc, x, b, w, sor);
  }
#undef B200AMG_ML_CASE
}
template <int T>
static int gs_tile_ctas() {
  static int cached = 0;
  if (!cached) {
    CUDA_OK(cudaFuncSetAttribute(gs_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kStages * sizeof(GsCtaStage))));
    int per_sm = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gs_tile_kernel<T>, kGsTileThreads, kStages * sizeof(GsCtaStage)));
    cached = std::max(1, per_sm);   // per SM: the call sites multiply by the SM count of the handle's device
  }
  return cached;
}
template <int T>
static void launch_gs_tile_T(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w,
                             int sor) {
  int ctas = std::min(M.gs_ntiles, gs_tile_ctas<T>() * h->num_sms);
  if (h->gs_tile_cta_limit > 0) ctas = std::min(ctas, h->gs_tile_cta_limit);   // experiment knob: fewer tiles in flight
  else if (M.gs_tile_ctas > 0) ctas = std::min(ctas, M.gs_tile_ctas);          // measured at finalize (tune_tile_ctas)
  // poll mode -1 (default): the focused spin pays on rows of >= 8 lanes (256^3 level 2: 5.82 -> 5.55 ms) and costs on
  // 4-lane rows (level 1: 7.27 -> 7.90), profiles/r02_tile_knobs_256.log
  const int poll_masked = h->gs_poll_masked >= 0 ? h->gs_poll_masked : (T >= 8 ? 2 : 1);
  gs_mail_prepare_kernel<<<1, 32, 0, h->stream>>>(M.mail_ctl);
  count_launch(h);
  gs_tile_kernel<T><<<ctas, kGsTileThreads, kStages * sizeof(GsCtaStage), h->stream>>>(
      M.gs_ntiles, M.gs_meta, M.gs_tile_wave, M.nlev, M.mail_ctl, A.ptr, A.idx, A.val, x, b, M.mail, w, sor, sc.backward, h->opaque_zero,
      h->gs_poll_sleep, h->gs_gate_sleep, poll_masked, std::max(1, h->gs_gate_dist), h->gs_debug);
  count_launch(h);
}
static void launch_gs_tile(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w,
                           int sor) {
  switch (M.gs_lanes) {
    case 1: launch_gs_tile_T<1>(h, M, A, sc, x, b, w, sor); break;
    case 2: launch_gs_tile_T<2>(h, M, A, sc, x, b, w, sor); break;
    case 4: launch_gs_tile_T<4>(h, M, A, sc, x, b, w, sor); break;
    case 8: launch_gs_tile_T<8>(h, M, A, sc, x, b, w, sor); break;
    case 16: launch_gs_tile_T<16>(h, M, A, sc, x, b, w, sor); break;
    default: launch_gs_tile_T<32>(h, M, A, sc, x, b, w, sor); break;
  }
}
// ---- one-cluster sweep for mid-size levels (cluster_gs.cuh) ----
template <int LOG_NC, int BS>
static bool launch_gs_cluster_T(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b,
                                double w, int sor) {
  constexpr int NC = 1 << LOG_NC;
  const size_t smem = (size_t)((M.n + NC - 1) / NC) * sizeof(double) + (size_t)(M.nlev + 1) * sizeof(int) + 16;
  static int state = 0;   // 0 unknown, 1 usable, -1 not schedulable on this device
  if (state < 0 || smem > 200 * 1024) return false;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(NC, 1, 1);
  cfg.blockDim = dim3(BS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (state == 0) {
    int nclusters = 0;
    if (cudaFuncSetAttribute(gs_cluster_kernel<LOG_NC, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        (NC > 8 && cudaFuncSetAttribute(gs_cluster_kernel<LOG_NC, BS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ||
        cudaOccupancyMaxActiveClusters(&nclusters, gs_cluster_kernel<LOG_NC, BS>, &cfg) != cudaSuccess || nclusters < 1) {
      cudaGetLastError();
      state = -1;
      return false;
    }
    state = 1;
  }
  CUDA_OK(cudaLaunchKernelEx(&cfg, gs_cluster_kernel<LOG_NC, BS>, (int)M.n, M.nlev, (const int*)M.d_fwd_lvlptr, (const int*)A.ptr,
                             (const int*)A.idx, (const double*)A.val, x, b, w, sor, sc.backward));
  count_launch(h);
  return true;
}
static bool launch_gs_cluster(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w,
                              int sor) {
  const int nc = h->gs_cluster_log_nc, bs = h->gs_cluster_threads;
#define B200AMG_CL(LN, BSZ) if (nc == LN && bs == BSZ && launch_gs_cluster_T<LN, BSZ>(h, M, A, sc, x, b, w, sor)) return true;
  B200AMG_CL(1, 1024) B200AMG_CL(2, 1024) B200AMG_CL(3, 1024) B200AMG_CL(4, 1024)
  B200AMG_CL(1, 256) B200AMG_CL(2, 256) B200AMG_CL(3, 256) B200AMG_CL(4, 256)
#undef B200AMG_CL
  return launch_gs_cluster_T<4, 1024>(h, M, A, sc, x, b, w, sor);
}
// attributes + schedulability of one instantiation, probed once (at b200amg_create: never inside a stream capture)
template <int LOG_NC, int T, int BS>
static int dsm_state() {
  static int state = 0;   // 1 usable, -1 not schedulable on this device
  if (state != 0) return state;
  constexpr int NC = 1 << LOG_NC;
  cudaLaunchConfig_t probe = {};
  probe.gridDim = dim3(NC, 1, 1);
  probe.blockDim = dim3(BS + 32, 1, 1);
  probe.dynamicSmemBytes = kDsmMaxDynSmem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  probe.attrs = attr;
  probe.numAttrs = NC > 1 ? 1 : 0;
  int nclusters = 1;
  if (cudaFuncSetAttribute(gs_dsm_kernel<LOG_NC, T, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDsmMaxDynSmem) != cudaSuccess ||
      (NC > 8 && cudaFuncSetAttribute(gs_dsm_kernel<LOG_NC, T, BS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ||
      (NC > 1 && (cudaOccupancyMaxActiveClusters(&nclusters, gs_dsm_kernel<LOG_NC, T, BS>, &probe) != cudaSuccess || nclusters < 1))) {
    cudaGetLastError();
    state = -1;
  } else {
    state = 1;
  }
  return state;
}
// the two-group variant (gs_dsm2_kernel): 2 x 256 consumer threads + the producer warp
template <int LOG_NC, int T>
static int dsm2_state() {
  static int state = 0;
  if (state != 0) return state;
  constexpr int NC = 1 << LOG_NC;
  cudaLaunchConfig_t probe = {};
  probe.gridDim = dim3(NC, 1, 1);
  probe.blockDim = dim3(2 * 256 + 32, 1, 1);
  probe.dynamicSmemBytes = kDsmMaxDynSmem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  probe.attrs = attr;
  probe.numAttrs = NC > 1 ? 1 : 0;
  int nclusters = 1;
  if (cudaFuncSetAttribute(gs_dsm2_kernel<LOG_NC, T, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDsmMaxDynSmem) != cudaSuccess ||
      (NC > 8 && cudaFuncSetAttribute(gs_dsm2_kernel<LOG_NC, T, 256>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ||
      (NC > 1 && (cudaOccupancyMaxActiveClusters(&nclusters, gs_dsm2_kernel<LOG_NC, T, 256>, &probe) != cudaSuccess || nclusters < 1))) {
    cudaGetLastError();
    state = -1;
  } else {
    state = 1;
  }
  return state;
}
template <int LOG_NC>
static void dsm_init_nc() {
  dsm2_state<LOG_NC, 4>(); dsm2_state<LOG_NC, 8>(); dsm2_state<LOG_NC, 16>(); dsm2_state<LOG_NC, 32>();
  dsm_state<LOG_NC, 4, 256>(); dsm_state<LOG_NC, 8, 256>(); dsm_state<LOG_NC, 16, 256>(); dsm_state<LOG_NC, 32, 256>();
  dsm_state<LOG_NC, 4, 512>(); dsm_state<LOG_NC, 8, 512>(); dsm_state<LOG_NC, 16, 512>(); dsm_state<LOG_NC, 32, 512>();
}
static void dsm_kernels_init() { dsm_init_nc<0>(); dsm_init_nc<1>(); dsm_init_nc<2>(); dsm_init_nc<3>(); dsm_init_nc<4>(); }
// ---- one-cluster sweep, x in distributed shared memory, dataflow hand-off through shared memory (dsm_gs.cuh) ----
template <int LOG_NC, int T, int BS>
static bool launch_gs_dsm_T(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w,
                            int sor) {
  constexpr int NC = 1 << LOG_NC;
  const size_t smem = dsm_smem_bytes(M.dsm_slots_max, M.nlev);
  if (dsm_state<LOG_NC, T, BS>() < 0 || smem > (size_t)kDsmMaxDynSmem) return false;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(NC, 1, 1);
  cfg.blockDim = dim3(BS + 32, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = NC > 1 ? 1 : 0;
  if (BS == 256 && h->gs_dsm2 && dsm2_state<LOG_NC, T>() > 0) {   // two consumer groups alternate the tiles (same plan)
    cfg.blockDim = dim3(2 * 256 + 32, 1, 1);
    CUDA_OK(cudaLaunchKernelEx(&cfg, gs_dsm2_kernel<LOG_NC, T, 256>, (int)M.n, M.dsm_ntiles, M.nlev, (const int4*)M.dsm_meta,
                               (const int2*)M.dsm_aux, (const int*)A.ptr, (const int*)M.dsm_code, (const double*)A.val,
                               (const int*)M.dsm_rowof, (const int*)M.dsm_own_off, (const int*)M.dsm_wave_tiles, x, b, w, sor,
                               sc.backward, h->opaque_zero, h->gs_dsm_fence, M.dsm_status, h->gs_debug));
    count_launch(h);
    return true;
  }
  CUDA_OK(cudaLaunchKernelEx(&cfg, gs_dsm_kernel<LOG_NC, T, BS>, (int)M.n, M.dsm_ntiles, M.nlev, (const int4*)M.dsm_meta,
                             (const int2*)M.dsm_aux, (const int*)A.ptr, (const int*)M.dsm_code, (const double*)A.val,
                             (const int*)M.dsm_rowof, (const int*)M.dsm_own_off, (const int*)M.dsm_wave_tiles, x, b, w, sor,
                             sc.backward, h->opaque_zero,
                             h->gs_dsm_fence, M.dsm_status, h->gs_debug));
  count_launch(h);
  return true;
}
template <int LOG_NC>
static bool launch_gs_dsm_NC(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b,
                             double w, int sor) {
  if (M.dsm_threads == 512) {
    switch (M.dsm_lanes) {
      case 4: return launch_gs_dsm_T<LOG_NC, 4, 512>(h, M, A, sc, x, b, w, sor);
      case 8: return launch_gs_dsm_T<LOG_NC, 8, 512>(h, M, A, sc, x, b, w, sor);
      case 16: return launch_gs_dsm_T<LOG_NC, 16, 512>(h, M, A, sc, x, b, w, sor);
      case 32: return launch_gs_dsm_T<LOG_NC, 32, 512>(h, M, A, sc, x, b, w, sor);
      default: return false;
    }
  }
  switch (M.dsm_lanes) {
    case 4: return launch_gs_dsm_T<LOG_NC, 4, 256>(h, M, A, sc, x, b, w, sor);
    case 8: return launch_gs_dsm_T<LOG_NC, 8, 256>(h, M, A, sc, x, b, w, sor);
    case 16: return launch_gs_dsm_T<LOG_NC, 16, 256>(h, M, A, sc, x, b, w, sor);
    case 32: return launch_gs_dsm_T<LOG_NC, 32, 256>(h, M, A, sc, x, b, w, sor);
    default: return false;
  }
}
static bool launch_gs_dsm(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w,
                          int sor) {
  if (M.dsm_ntiles <= 0 || M.dsm_lanes < 4 || M.nlev <= 0 || !M.dsm_code) return false;
  switch (M.dsm_log_nc) {
    case 0: return launch_gs_dsm_NC<0>(h, M, A, sc, x, b, w, sor);
    case 1: return launch_gs_dsm_NC<1>(h, M, A, sc, x, b, w, sor);
    case 2: return launch_gs_dsm_NC<2>(h, M, A, sc, x, b, w, sor);
    case 3: return launch_gs_dsm_NC<3>(h, M, A, sc, x, b, w, sor);
    case 4: return launch_gs_dsm_NC<4>(h, M, A, sc, x, b, w, sor);
    default: return false;
  }
}
// ---- blocked sweep (block_gs.cuh) ----
template <int T>
static void gs_block_set_attr() {
  CUDA_OK(cudaFuncSetAttribute(gs_block_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBgSmemBytes));
}
static void gs_block_kernels_init() {
  gs_block_set_attr<1>(); gs_block_set_attr<2>(); gs_block_set_attr<4>(); gs_block_set_attr<8>(); gs_block_set_attr<16>(); gs_block_set_attr<32>();
}
static void launch_gs_block(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w,
                            int sor) {
  const DevBlockPlan& B = M.block;
  CUDA_OK(cudaMemsetAsync(B.ctl, 0, sizeof(unsigned) * B.ctl_words, h->stream));
  const int ctas = std::min(B.ntiles, h->num_sms);
  const int2* req = sc.backward ? B.req_bwd : B.req_fwd;
  const int* order = sc.backward ? B.order_bwd : B.order_fwd;
  const int* code = sc.backward ? B.code_bwd : B.code_fwd;
#define B200AMG_BG_CASE(TT)                                                                                                       \
  case TT:                                                                                                                        \
    gs_block_kernel<TT><<<ctas, kBgThreads, kBgSmemBytes, h->stream>>>(B.ntiles, B.tile, B.stage_meta, B.stage_aux, B.stage_auxb, \
                                                                      B.steps, req, order, B.ctl, A.ptr, code, B.dpos, A.val, x, b, w, \
                                                                      sor,                                                    \
                                                                      sc.backward, h->gs_fault, h->gs_debug);                    \
    break;
  switch (B.lanes) {
    B200AMG_BG_CASE(1) B200AMG_BG_CASE(2) B200AMG_BG_CASE(4) B200AMG_BG_CASE(8) B200AMG_BG_CASE(16)
    default:
      gs_block_kernel<32><<<ctas, kBgThreads, kBgSmemBytes, h->stream>>>(B.ntiles, B.tile, B.stage_meta, B.stage_aux, B.stage_auxb, B.steps,
                                                                        req, order, B.ctl, A.ptr, code, B.dpos, A.val, x, b, w, sor, sc.backward,
                                                                        h->gs_fault, h->gs_debug);
  }
#undef B200AMG_BG_CASE
  count_launch(h);
}
// ---- pass sweep (pass_gs.cuh) ----
template <int T>
static void gs_pass_set_attr() {
  CUDA_OK(cudaFuncSetAttribute(gs_pass_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPgSmemBytes));
}
static void gs_pass_kernels_init() {
  gs_pass_set_attr<1>(); gs_pass_set_attr<2>(); gs_pass_set_attr<4>(); gs_pass_set_attr<8>(); gs_pass_set_attr<16>(); gs_pass_set_attr<32>();
}
static void launch_gs_pass(H* h, const SmootherMatrix& M, const DevSchedule& sc, double* x, const double* b, double w, int sor) {
  const DevBlockPlan& B = M.block;
  const DevPassPlan& Q = M.pass;
  const DevPassPlan::Dir& D = Q.dir[sc.backward ? 1 : 0];
  CUDA_OK(cudaMemsetAsync(B.ctl, 0, sizeof(unsigned) * B.ctl_words, h->stream));
  const int ctas = std::min(B.ntiles, h->num_sms);
  const int* order = sc.backward ? B.order_bwd : B.order_fwd;
#define B200AMG_PG_CASE(TT)                                                                                                                 \
  case TT:                                                                                                                                  \
    gs_pass_kernel<TT><<<ctas, kPgThreads, kPgSmemBytes, h->stream>>>(B.ntiles, D.tile, D.pass, D.preq, D.req, order, B.ctl, D.val, D.idx,          \
                                                                     M.diag, x, b, w, sor, h->gs_fault, h->gs_debug);                       \
    break;
  switch (Q.lanes) {
    B200AMG_PG_CASE(1) B200AMG_PG_CASE(2) B200AMG_PG_CASE(4) B200AMG_PG_CASE(8) B200AMG_PG_CASE(16) B200AMG_PG_CASE(32)
    default: REQUIRE(false, B200AMG_ERR_STATE, "pass sweep: unsupported lane count %d", Q.lanes);
  }
#undef B200AMG_PG_CASE
  count_launch(h);
}
constexpr int64_t kGsCtaXsRows = 12288;   // x of the level fits next to the tile ring in shared memory
template <int T, bool XS>
static void gs_cta_set_attr() {
  CUDA_OK(cudaFuncSetAttribute(gs_cta_kernel<T, XS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(kStages * sizeof(GsCtaStage) + (XS ? kGsCtaXsRows * sizeof(double) : 0))));
}
static void gs_cta_kernels_init() {
  gs_cta_set_attr<1, false>(); gs_cta_set_attr<2, false>(); gs_cta_set_attr<4, false>(); gs_cta_set_attr<8, false>();
  gs_cta_set_attr<16, false>(); gs_cta_set_attr<32, false>();
  gs_cta_set_attr<1, true>(); gs_cta_set_attr<2, true>(); gs_cta_set_attr<4, true>(); gs_cta_set_attr<8, true>();
  gs_cta_set_attr<16, true>(); gs_cta_set_attr<32, true>();
}
template <int T>
static void launch_gs_cta_T(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w,
                            int sor) {
  const bool xs = M.n <= kGsCtaXsRows;
  const size_t smem = kStages * sizeof(GsCtaStage) + (xs ? (size_t)M.n * sizeof(double) : 0);
  if (xs)
    gs_cta_kernel<T, true><<<1, kGsCtaThreads, smem, h->stream>>>((int)M.n, A.ntiles, A.meta, A.ptr, A.idx, A.val, M.d_fwd_lvlptr, M.nlev,
                                                                 x, b, w, sor, sc.backward, h->opaque_zero);
  else
    gs_cta_kernel<T, false><<<1, kGsCtaThreads, smem, h->stream>>>((int)M.n, A.ntiles, A.meta, A.ptr, A.idx, A.val, M.d_fwd_lvlptr,
                                                                  M.nlev, x, b, w, sor, sc.backward, h->opaque_zero);
  count_launch(h);
}
static void launch_gs_cta(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w,
                          int sor) {
  const double mean = M.n ? (double)A.nnz / (double)M.n : 0.0;
  int T = 1;
  while (T < 32 && T < mean) T *= 2;
  switch (T) {
    case 1: launch_gs_cta_T<1>(h, M, A, sc, x, b, w, sor); break;
    case 2: launch_gs_cta_T<2>(h, M, A, sc, x, b, w, sor); break;
    case 4: launch_gs_cta_T<4>(h, M, A, sc, x, b, w, sor); break;
    case 8: launch_gs_cta_T<8>(h, M, A, sc, x, b, w, sor); break;
    case 16: launch_gs_cta_T<16>(h, M, A, sc, x, b, w, sor); break;
    default: launch_gs_cta_T<32>(h, M, A, sc, x, b, w, sor); break;
  }
}
static void launch_sweep(H* h, const SmootherMatrix& M, const DevSchedule& sc, double* x, const double* b, double w, int sor) {
  const DevCsr& A = M.walked();
  if (M.pass.ok) { launch_gs_pass(h, M, sc, x, b, w, sor); return; }
  if (M.block.ok) { launch_gs_block(h, M, A, sc, x, b, w, sor); return; }
  if (h->gs_mode >= 1 && h->gs_dsm && M.d_fwd_lvlptr && M.dsm_ntiles > 0 && M.dsm_log_nc <= h->gs_dsm_max_log_nc &&
      !(sc.nlev > 0 && M.n / sc.nlev >= h->gs_mail_min_width)) {
    if (launch_gs_dsm(h, M, A, sc, x, b, w, sor)) return;
    REQUIRE(h->gs_dsm < 2 || M.dsm_ntiles <= 0, B200AMG_ERR_CUDA, "the distributed-shared-memory sweep could not be launched (n = %lld, %d CTAs)",
            (long long)M.n, 1 << M.dsm_log_nc);
  }
  // Which sweep: measured on B200 (tools/tune_kernels.py, profiles/): one CTA wins while x fits in shared
  // memory (~1 us per wavefront); the per-row mailbox sweep wins on wide wavefronts (>= ~1000 rows); the
  // wavefront-counter sweep in between.
  if (h->gs_mode >= 1 && M.n <= h->gs_cta_rows && A.ntiles > 0 && M.d_fwd_lvlptr) { launch_gs_cta(h, M, A, sc, x, b, w, sor); return; }
  const bool wide = sc.nlev > 0 && M.n / sc.nlev >= h->gs_mail_min_width;
  // mid-size level with long rows and narrow wavefronts: one cluster, x in distributed shared memory
  if (h->gs_mode >= 1 && h->gs_cluster && !wide && M.d_fwd_lvlptr && M.n <= h->gs_cluster_rows && A.nrows > 0 &&
      (double)A.nnz / (double)A.nrows >= 16.0 && launch_gs_cluster(h, M, A, sc, x, b, w, sor))
    return;
  // measured, 256^3 RS hierarchy (us per wavefront): TMA-fed mailbox sweep 2.3 at one thread per row (stencil rows)
  // but 6-7 with several lanes per row, where the ticket mailbox sweep does 3.0-4.7 and the counter sweep 4.6-6.0
  if (h->gs_mode == 2 && M.mail && M.gs_ntiles > 0 && wide && (M.gs_lanes == 1 || h->gs_tile_any_lanes)) { launch_gs_tile(h, M, A, sc, x, b, w, sor); return; }
  if (h->gs_mode >= 2 && M.mail && wide) { launch_mail(h, M, A, sc, x, b, w, sor); return; }
  if (h->gs_mode >= 1) { launch_dataflow(h, A, sc, x, b, w, sor, M.mail, M.mail_ctl); return; }
  switch (A.lanes) {
    case 2: launch_sweep_T<2>(h, A, sc, x, b, w, sor); break;
    case 4: launch_sweep_T<4>(h, A, sc, x, b, w, sor); break;
    case 8: launch_sweep_T<8>(h, A, sc, x, b, w, sor); break;
    case 16: launch_sweep_T<16>(h, A, sc, x, b, w, sor); break;
    default: launch_sweep_T<32>(h, A, sc, x, b, w, sor); break;
  }
}

// smooth!(x, s, b) for one configured smoother on a prepared matrix.  temp: n scratch doubles.
// x_is_zero: the caller guarantees x == 0 on entry (enables the exact zero-guess Jacobi shortcut).
static void smooth(H* h, const SmootherMatrix& M, const SmootherCfg& c, double* x, const double* b, double* temp,
                   bool x_is_zero) {
  if (c.kind == B200AMG_SMOOTHER_NONE || M.n == 0) return;
  const DevCsr& A = M.walked();
  if (c.kind == B200AMG_SMOOTHER_JACOBI) {
    const bool general = M.symmetry == B200AMG_SYMMETRY_NONE;
    double* cur = x;
    double* other = temp;
    for (int it = 0; it < c.iter; ++it) {
      if (it == 0 && x_is_zero) {
        // elementwise: safe in place, no buffer swap
        jacobi_zero_guess_kernel<<<grid_for(M.n), kThreads, 0, h->stream>>>(M.n, M.diag, b, cur, c.omega, general ? 1 : 0);
        count_launch(h);
        continue;
      } else if (general) {
        launch_jacobi_general(h, A, M.diag, cur, b, other, c.omega);
      } else {
        launch_jacobi_fast(h, A, cur, b, other, c.omega);
      }
      std::swap(cur, other);
    }
    if (cur != x) CUDA_OK(cudaMemcpyAsync(x, cur, sizeof(double) * M.n, cudaMemcpyDeviceToDevice, h->stream));
    return;
  }
  const int sor = c.kind == B200AMG_SMOOTHER_SOR;
  for (int it = 0; it < c.iter; ++it) {
    if (c.sweep == 1 || c.sweep == 3) launch_sweep(h, M, M.fwd, x, b, c.omega, sor);
    if (c.sweep == 2 || c.sweep == 3) launch_sweep(h, M, M.bwd, x, b, c.omega, sor);
  }
}

static void norm2_async(H* h, int64_t n, const double* v, double* out_dev) {
  dot_partial_kernel<<<kRedBlocks, kThreads, 0, h->stream>>>(n, v, v, h->partial);
  count_launch(h);
  reduce_final_kernel<<<1, kThreads, 0, h->stream>>>(kRedBlocks, h->partial, out_dev, 1);
  count_launch(h);
}
static void dot_async(H* h, int64_t n, const double* a, const double* b, double* out_dev) {
  dot_partial_kernel<<<kRedBlocks, kThreads, 0, h->stream>>>(n, a, b, h->partial);
  count_launch(h);
  reduce_final_kernel<<<1, kThreads, 0, h->stream>>>(kRedBlocks, h->partial, out_dev, 0);
  count_launch(h);
}
static double read_scalar(H* h, const double* dev) {
  CUDA_OK(cudaMemcpyAsync(h->h_scalars, dev, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return h->h_scalars[0];
}

// the host-callable coarse solver, run by the CUDA runtime between the two copies of coarse_solve (stream order; also
// inside a captured cycle graph, as a host node).  No CUDA calls in here.
static void CUDART_CB coarse_host_trampoline(void* p) {
  H* h = static_cast<H*>(p);
  const int32_t rc = h->coarse_fn(h->coarse_user, h->nfinal, 1, h->coarse_hx, h->coarse_hb);
  ++h->coarse_fn_calls;
  if (rc != 0 && h->coarse_fn_status == 0) h->coarse_fn_status = rc;
}

static void coarse_solve(H* h, double* x, const double* b) {
  if (h->nfinal == 0) return;
  if (h->coarse_fn) {   // cs(x, b) on the host: src/multilevel.jl:180,228 with a callable from src/coarse_solver.jl:24-58
    const size_t bytes = sizeof(double) * (size_t)h->nfinal;
    CUDA_OK(cudaMemcpyAsync(h->coarse_hb, b, bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaLaunchHostFunc(h->stream, coarse_host_trampoline, h));
    CUDA_OK(cudaMemcpyAsync(x, h->coarse_hx, bytes, cudaMemcpyHostToDevice, h->stream));
    return;
  }
  dense_gemv_kernel<<<grid_for(h->nfinal), kThreads, 0, h->stream>>>((int)h->nfinal, h->coarse_inv, b, x);
  count_launch(h);
}

