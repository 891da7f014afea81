// The cycle (__solve!), its captured graphs and the phase timers.  Part of engine.cu (one translation unit).
#pragma once

// ------------------------------------------------------------------------------------------
// the cycle: __solve!(x, ml, cycle, b, lvl)  — src/multilevel.jl:214-239, recursion :200-212
// ------------------------------------------------------------------------------------------
// The six sections the reference times with @timeit_debug (src/multilevel.jl:216-236), under the same names: an NVTX range
// around the launches of every phase (nsys / ncu --nvtx line the device work up with the reference's timer labels; when the
// cycle is replayed as a CUDA graph the ranges mark its capture), and CUDA-event timers for b200amg_profile_cycle.
static const char* const kPhaseNames[6] = {"Presmoother", "Residual eval", "Restriction", "Coarse solve", "Prolongation", "Postsmoother"};
struct PhaseTimer {
  H* h;
  int slot;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  PhaseTimer(H* h_, int lvl, int phase) : h(h_), slot(lvl * 6 + phase) {
    char name[48];
    snprintf(name, sizeof name, "%s L%d", kPhaseNames[phase], lvl);
    nvtxRangePushA(name);
    if (h->profiling) {
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0, h->stream);
    }
  }
  ~PhaseTimer() {
    if (h->profiling) {
      cudaEventRecord(e1, h->stream);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (h->prof_ms && slot < (int)h->prof_ms->size()) (*h->prof_ms)[slot] += ms;
      cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    nvtxRangePop();
  }
};

static void solve_level(H* h, double* x, int cycle, const double* b, int lvl, bool x_is_zero) {
  Level& L = *h->levels[lvl];
  { PhaseTimer t(h, lvl, 0); smooth(h, L.M, L.pre, x, b, L.temp, x_is_zero); }                    // :216
  { PhaseTimer t(h, lvl, 1); residual(h, L.M.A, x, b, L.res); }                                  // :219-220
  { PhaseTimer t(h, lvl, 2); spmv(h, L.R, L.res, L.coarse_b); }                                  // :223
  CUDA_OK(cudaMemsetAsync(L.coarse_x, 0, sizeof(double) * (size_t)std::max<int64_t>(L.nc, 1), h->stream));  // :226
  if (lvl == (int)h->levels.size() - 1) {
    PhaseTimer t(h, lvl, 3);
    coarse_solve(h, L.coarse_x, L.coarse_b);                                                     // :228
  } else if (cycle == B200AMG_CYCLE_V) {
    solve_level(h, L.coarse_x, B200AMG_CYCLE_V, L.coarse_b, lvl + 1, true);                       // :200-202
  } else if (cycle == B200AMG_CYCLE_W) {
    solve_level(h, L.coarse_x, B200AMG_CYCLE_W, L.coarse_b, lvl + 1, true);                       // :204-207
    solve_level(h, L.coarse_x, B200AMG_CYCLE_W, L.coarse_b, lvl + 1, false);
  } else {
    solve_level(h, L.coarse_x, B200AMG_CYCLE_F, L.coarse_b, lvl + 1, true);                       // :209-212
    solve_level(h, L.coarse_x, B200AMG_CYCLE_V, L.coarse_b, lvl + 1, false);
  }
  { PhaseTimer t(h, lvl, 4); spmv_add(h, L.P, L.coarse_x, x); }                                  // :233-234
  { PhaseTimer t(h, lvl, 5); smooth(h, L.M, L.post, x, b, L.temp, false); }                      // :236
}

// one "iteration body" of _solve! on the internal level-0 vectors (multilevel.jl:179-183)
static void cycle_body(H* h, int cycle, bool x_is_zero) {
  if (h->levels.empty()) coarse_solve(h, h->x0, h->b0);
  else solve_level(h, h->x0, cycle, h->b0, 0, x_is_zero);
}

static int64_t estimate_launches(H* h, int cycle, int lvl) {
  if (h->levels.empty()) return 1;
  const Level& L = *h->levels[lvl];
  auto sm = [&](const SmootherCfg& c) -> int64_t {
    if (c.kind == 0) return 0;
    if (c.kind == B200AMG_SMOOTHER_JACOBI) return c.iter + 1;
    int64_t per = 0;
    if (c.sweep == 1 || c.sweep == 3) per += h->gs_mode >= 1 ? 2 : (int64_t)L.M.fwd.items.size();
    if (c.sweep == 2 || c.sweep == 3) per += h->gs_mode >= 1 ? 2 : (int64_t)L.M.bwd.items.size();
    return per * c.iter;
  };
  int64_t n = sm(L.pre) + sm(L.post) + 4;
  if (lvl == (int)h->levels.size() - 1) return n + 1;
  if (cycle == B200AMG_CYCLE_V) return n + estimate_launches(h, cycle, lvl + 1);
  if (cycle == B200AMG_CYCLE_W) return n + 2 * estimate_launches(h, cycle, lvl + 1);
  return n + estimate_launches(h, B200AMG_CYCLE_F, lvl + 1) + estimate_launches(h, B200AMG_CYCLE_V, lvl + 1);
}

static const int64_t kMaxGraphNodes = 150000;

static void ensure_cycle_graph(H* h, int cycle) {
  if (!h->use_graphs || h->cycle_graph[cycle] || h->cycle_graph_launches[cycle] < 0) return;
  if (estimate_launches(h, cycle, 0) > kMaxGraphNodes) { h->cycle_graph_launches[cycle] = -1; return; }
  cudaGraph_t g = nullptr;
  h->capturing = true;
  h->capture_count = 0;
  CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  try {
    cycle_body(h, cycle, false);
  } catch (...) {
    cudaStreamEndCapture(h->stream, &g);
    if (g) cudaGraphDestroy(g);
    h->capturing = false;
    throw;
  }
  CUDA_OK(cudaStreamEndCapture(h->stream, &g));
  h->capturing = false;
  CUDA_OK(cudaGraphInstantiate(&h->cycle_graph[cycle], g, 0));
  CUDA_OK(cudaGraphDestroy(g));
  h->cycle_graph_launches[cycle] = h->capture_count;
}

static void run_cycle(H* h, int cycle) {
  if (h->part) { cycle_body_part(h, cycle); return; }
  ensure_cycle_graph(h, cycle);
  if (h->use_graphs && h->cycle_graph[cycle]) {
    CUDA_OK(cudaGraphLaunch(h->cycle_graph[cycle], h->stream));
    h->launches += h->cycle_graph_launches[cycle];
  } else {
    cycle_body(h, cycle, false);
  }
}

// res = b0 - A x0 ; scalars[0] = ||res||      (multilevel.jl:188-190)
static void residual_norm(H* h) {
  const DevCsr& A = h->levels.empty() ? h->finalA : h->levels[0]->M.A;
  double* res = h->levels.empty() ? h->res_final : h->levels[0]->res;
  const bool timed = h->time_residual && h->res_events_used + 2 <= (int)h->res_events.size();
  if (timed) CUDA_OK(cudaEventRecord(h->res_events[h->res_events_used], h->stream));
  residual(h, A, h->x0, h->b0, res);
  if (timed) {
    CUDA_OK(cudaEventRecord(h->res_events[h->res_events_used + 1], h->stream));
    h->res_events_used += 2;
  }
  norm2_async(h, h->n0, res, h->scalars);
}


