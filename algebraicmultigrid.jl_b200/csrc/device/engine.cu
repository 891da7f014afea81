// libb200amg.so — the C-ABI device engine declared in include/b200amg.h.
//
// Owns device copies of every level of an AMG hierarchy and runs the whole solve phase of
// AlgebraicMultigrid.jl (src/multilevel.jl:152-239, src/smoother.jl, src/preconditioner.jl:12-24)
// on one B200: the cycle is a static sequence of kernels captured once per cycle type into a
// CUDA graph and replayed per iteration.  There is no CPU fallback: without a device every
// compute entry point fails with B200AMG_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>   // header-only: ranges are no-ops unless a profiler is attached
#include <nccl.h>   // types and prototypes only: the library is dlopen'ed when a partition is requested

#include "b200amg.h"
#include "kernels.cuh"
#include "stream.cuh"
#include "host_csr.h"
#include "partition.h"
#include "cluster_gs.cuh"
#include "dsm_gs.cuh"
#include "block_plan.h"
#include "block_gs.cuh"
#include "pass_plan.h"
#include "pass_gs.cuh"
#include "peer_halo.cuh"
#include "spgemm.cuh"

using namespace b200amg;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int32_t fail(int32_t code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
struct AmgError {
  int32_t code;
  std::string msg;
};
#define CUDA_OK(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      char _b[512];                                                                                \
      snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      throw AmgError{_e == cudaErrorMemoryAllocation ? B200AMG_ERR_OOM : B200AMG_ERR_CUDA, _b};    \
    }                                                                                              \
  } while (0)
#define REQUIRE(cond, code, ...)                                     \
  do {                                                               \
    if (!(cond)) {                                                   \
      char _b[512];                                                  \
      snprintf(_b, sizeof _b, __VA_ARGS__);                          \
      throw AmgError{code, _b};                                      \
    }                                                                \
  } while (0)
#define API_BEGIN try {
#define API_END                                         \
  }                                                     \
  catch (const AmgError& e) {                           \
    return fail(e.code, "%s", e.msg.c_str());           \
  }                                                     \
  catch (const std::bad_alloc&) {                       \
    return fail(B200AMG_ERR_OOM, "host out of memory"); \
  }                                                     \
  catch (const std::exception& e) {                     \
    return fail(B200AMG_ERR_BAD_ARG, "%s", e.what());   \
  }                                                     \
  return B200AMG_OK;

// ------------------------------------------------------------------------------------------
// NCCL, resolved at run time (a single-GPU process never needs libnccl).  RTLD_NOLOAD first: a host
// that already initialised torch.distributed has its NCCL loaded under the same soname.
// ------------------------------------------------------------------------------------------
struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  bool ok = false;
};
static NcclApi& nccl_api() {
  static NcclApi api;
  if (api.ok) return api;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) throw AmgError{B200AMG_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror()};
#define B200AMG_NCCL_SYM(field, name)                                                         \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(lib, name));                          \
  if (!api.field) throw AmgError{B200AMG_ERR_NCCL, std::string("libnccl lacks ") + name};
  B200AMG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  B200AMG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  B200AMG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  B200AMG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
  B200AMG_NCCL_SYM(GroupStart, "ncclGroupStart")
  B200AMG_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  B200AMG_NCCL_SYM(Send, "ncclSend")
  B200AMG_NCCL_SYM(Recv, "ncclRecv")
  B200AMG_NCCL_SYM(AllReduce, "ncclAllReduce")
  B200AMG_NCCL_SYM(Broadcast, "ncclBroadcast")
  B200AMG_NCCL_SYM(AllGather, "ncclAllGather")
#undef B200AMG_NCCL_SYM
  api.ok = true;
  return api;
}
#define NCCL_OK(expr)                                                                                      \
  do {                                                                                                     \
    ncclResult_t _r = (expr);                                                                              \
    if (_r != ncclSuccess) {                                                                               \
      char _b[512];                                                                                        \
      snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, nccl_api().GetErrorString(_r), __FILE__, __LINE__); \
      throw AmgError{B200AMG_ERR_NCCL, _b};                                                                \
    }                                                                                                      \
  } while (0)


// The CSC arrays of an m x n matrix ARE the CSR arrays of its n x m transpose.
static HostCsr stage_csc_as_rows_of_transpose(const b200amg_csc_t* M) {
  REQUIRE(M && M->colptr && (M->index_bits == 32 || M->index_bits == 64) && (M->index_base == 0 || M->index_base == 1),
          B200AMG_ERR_BAD_ARG, "bad matrix descriptor (index_bits must be 32/64, index_base 0/1)");
  REQUIRE(M->m >= 0 && M->n >= 0 && M->m < INT32_MAX && M->n < INT32_MAX, B200AMG_ERR_UNSUPPORTED,
          "matrix dimension does not fit the int32 device index width");
  HostCsr out;
  out.nrows = M->n;
  out.ncols = M->m;
  out.ptr.resize(M->n + 1);
  const int base = M->index_base;
  int64_t nnz;
  if (M->index_bits == 64) {
    const int64_t* cp = (const int64_t*)M->colptr;
    nnz = cp[M->n] - base;
    REQUIRE(nnz >= 0 && nnz < INT32_MAX, B200AMG_ERR_UNSUPPORTED, "nnz does not fit the int32 device index width");
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j <= M->n; ++j) out.ptr[j] = (int)(cp[j] - base);
  } else {
    const int32_t* cp = (const int32_t*)M->colptr;
    nnz = cp[M->n] - base;
    REQUIRE(nnz >= 0, B200AMG_ERR_BAD_ARG, "negative nnz");
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j <= M->n; ++j) out.ptr[j] = cp[j] - base;
  }
  REQUIRE(nnz == 0 || (M->rowval && M->nzval), B200AMG_ERR_BAD_ARG, "null rowval/nzval");
  out.idx.resize(nnz);
  out.val.resize(nnz);
  const double* nz = M->nzval;
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nnz; ++k) out.val[k] = nz[k];
  if (M->index_bits == 64) {
    const int64_t* rv = (const int64_t*)M->rowval;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nnz; ++k) out.idx[k] = (int)(rv[k] - base);
  } else {
    const int32_t* rv = (const int32_t*)M->rowval;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nnz; ++k) out.idx[k] = rv[k] - base;
  }
  // validation (exceptions must not leave an OpenMP region: collect the first kind of violation, report after)
  int bad = 0;
  for (int64_t j = 0; j < M->n && !bad; ++j)
    if (out.ptr[j] > out.ptr[j + 1] || out.ptr[j] < 0 || out.ptr[j + 1] > nnz) bad = 1;
  REQUIRE(!bad && (M->n == 0 || out.ptr[0] == 0), B200AMG_ERR_BAD_ARG, "colptr not monotone");
  const int64_t mrows = M->m;
#pragma omp parallel for schedule(static) reduction(max : bad)
  for (int64_t j = 0; j < M->n; ++j)
    for (int k = out.ptr[j]; k < out.ptr[j + 1]; ++k) {
      if (out.idx[k] < 0 || out.idx[k] >= mrows) bad = std::max(bad, 2);
      else if (k != out.ptr[j] && out.idx[k - 1] >= out.idx[k]) bad = std::max(bad, 1);
    }
  REQUIRE(bad != 2, B200AMG_ERR_BAD_ARG, "row index out of range");
  REQUIRE(bad != 1, B200AMG_ERR_BAD_ARG, "row indices must be sorted and unique inside each column");
  return out;
}


// operator given as (stored CSC, adjoint flag) -> the operator compressed by ITS rows
static HostCsr stage_operator_by_rows(const b200amg_csc_t* M) {
  HostCsr t = stage_csc_as_rows_of_transpose(M);  // rows of stored'
  if (M->adjoint) return t;                       // operator == stored'
  return transpose(t);                            // operator == stored
}


// ------------------------------------------------------------------------------------------
// device objects
// ------------------------------------------------------------------------------------------
template <typename T>
static T* dev_alloc(int64_t count) {
  T* p = nullptr;
  CUDA_OK(cudaMalloc(&p, sizeof(T) * (size_t)std::max<int64_t>(count, 1)));
  return p;
}
template <typename T, typename Al>
static T* dev_upload(const std::vector<T, Al>& v, int64_t pad = 0) {
  T* p = dev_alloc<T>((int64_t)v.size() + pad);
  if (!v.empty()) CUDA_OK(cudaMemcpy(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return p;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
// B200AMG_VERBOSE_UPLOAD=1: wall-clock of the host-side stages of add_level on stderr
struct UploadTimer {
  const char* what;
  double t0;
  bool on;
  static double now() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
  }
  explicit UploadTimer(const char* w) : what(w), t0(now()), on(env_int("B200AMG_VERBOSE_UPLOAD", 0) != 0) {}
  ~UploadTimer() {
    if (on) fprintf(stderr, "[b200amg] upload %-28s %8.3f s\n", what, now() - t0);
  }
};

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
static BlockPlanParams block_params_from_env() {
  BlockPlanParams prm;
  prm.stage_nnz = kBgStageNnz; prm.stage_rows = kBgStageRows; prm.window = kBgWindow; prm.depth = kBgDepth;
  prm.step_us = 1e-3 * env_int("B200AMG_BLOCK_STEP_NS", 220);
  prm.cta_gbs = env_int("B200AMG_BLOCK_CTA_GBS", 55);
  prm.cap_step_to_stage = env_int("B200AMG_BLOCK_XCAP", 1);
  prm.force_tile_rows = env_int("B200AMG_BLOCK_TILE_ROWS", 0);
  prm.force_a = env_int("B200AMG_BLOCK_A", 0);
  prm.force_b = env_int("B200AMG_BLOCK_B", 0);
  prm.max_lanes = 32;
  prm.verbose = env_int("B200AMG_BLOCK_VERBOSE", 0);
  return prm;
}

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
    cached = std::max(1, per_sm) * kNumSM;
  }
  return cached;
}
template <int T, int BS>
static void launch_dataflow_T(H* h, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w, int sor,
                              uint4* mail, unsigned* mail_ctl) {
  CUDA_OK(cudaMemsetAsync(sc.counters, 0, sizeof(unsigned) * (size_t)(sc.nlev + 2) * kGsCounterStride, h->stream));
  if (mail && h->gs_counter_mail && !h->gs_debug) {
    const int ctas = std::min(sc.ntasks, gs_dataflow_ctas<T, BS, true>());
    gs_mail_prepare_kernel<<<1, 32, 0, h->stream>>>(mail_ctl);   // new epoch for the mailbox flags
    count_launch(h);
    gs_dataflow_kernel<T, BS, true><<<ctas, BS, 0, h->stream>>>(sc.ntasks, sc.tasks, sc.counters, A.ptr, A.idx, A.val, x, b, w, sor,
                                                               sc.backward, h->gs_acquire, h->opaque_zero, nullptr, mail, mail_ctl);
  } else {
    const int ctas = std::min(sc.ntasks, gs_dataflow_ctas<T, BS, false>());
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
    cached = std::max(1, per_sm) * kNumSM;
  }
  return cached;
}
template <int T, int BS>
static void launch_mail_T(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w,
                          int sor) {
  const int ctas = std::min(sc.ntasks, gs_mail_ctas<T, BS>());
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
    cached = std::max(1, per_sm) * kNumSM;
  }
  return cached;
}
template <int T>
static void launch_gs_tile_T(H* h, const SmootherMatrix& M, const DevCsr& A, const DevSchedule& sc, double* x, const double* b, double w,
                             int sor) {
  int ctas = std::min(M.gs_ntiles, gs_tile_ctas<T>());
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


// ------------------------------------------------------------------------------------------
// Row-partitioned fine level (config C4): halo exchange over NCCL, coarse levels on rank 0
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) halo_pack_kernel(int n, const int* __restrict__ idx, const double* __restrict__ v,
                                                             double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v[idx[i]];
}

// v is laid out [owned | halo]: gather what the neighbours need, exchange, receive straight into the halo
static int peer_channel(const Part& P, const double* v) { return v == P.x ? 0 : v == P.temp ? 1 : v == P.res ? 2 : -1; }
// the kernel(s) that read the halo of the last exchange have been enqueued on the compute stream: tell the senders
static void halo_consumed(H* h) {
  if (!h->peer_pending) return;
  const int slot = h->peer_pending->level * kPeerChannels + h->peer_pending_ch;
  halo_ack_kernel<<<1, 32, 0, h->stream>>>(h->peer.d_tab + slot, h->peer.sync + (size_t)slot * kPeerWords);
  count_launch(h);
  h->peer_pending = nullptr;
}
static void halo_exchange(H* h, Part& P, double* v) {
  const PartPlan& pl = P.plan;
  if (h->peer.on) {
    const int ch = peer_channel(P, v);
    REQUIRE(ch >= 0, B200AMG_ERR_STATE, "peer halo exchange of a vector that was not exported");
    REQUIRE(!h->peer_pending, B200AMG_ERR_STATE, "internal: a halo exchange was started before the previous one was acknowledged");
    const int slot = P.level * kPeerChannels + ch;
    unsigned long long* sync = h->peer.sync + (size_t)slot * kPeerWords;
    const int nsend = pl.send_off[pl.world];
    const int blocks = std::max(1, std::min(64, (nsend + 255) / 256));
    halo_push_kernel<<<blocks, 256, 0, h->stream>>>(h->peer.d_tab + slot, P.send_idx, v, sync, h->peer.tickets + slot, h->gs_fault);
    count_launch(h);
    halo_wait_kernel<<<1, 32, 0, h->stream>>>(h->peer.d_tab + slot, sync, h->gs_fault);
    count_launch(h);
    h->peer_pending = &P;
    h->peer_pending_ch = ch;
    h->peer_exchanges++;
    return;
  }
  NcclApi& nc = nccl_api();
  const int nsend = pl.send_off[pl.world];
  if (nsend > 0) {
    halo_pack_kernel<<<grid_for(nsend), kThreads, 0, h->stream>>>(nsend, P.send_idx, v, P.sendbuf);
    count_launch(h);
  }
  NCCL_OK(nc.GroupStart());
  for (int q = 0; q < pl.world; ++q) {
    if (q == pl.rank) continue;
    const int ns = pl.send_off[q + 1] - pl.send_off[q], nr = pl.recv_off[q + 1] - pl.recv_off[q];
    if (ns > 0) NCCL_OK(nc.Send(P.sendbuf + pl.send_off[q], (size_t)ns, ncclDouble, q, h->comm, h->stream));
    if (nr > 0) NCCL_OK(nc.Recv(v + pl.nloc + pl.recv_off[q], (size_t)nr, ncclDouble, q, h->comm, h->stream));
  }
  NCCL_OK(nc.GroupEnd());
  h->collectives++;
}

// The same exchange, split in two so that work which does not touch the halo can run in between: _begin forks onto the
// communication stream (after everything enqueued so far on the compute stream: producers of v's owned part, earlier readers
// of its halo part), _end joins it back.  Both are captured into the whole-cycle graph as a fork / join.
static void halo_exchange_begin(H* h, Part& P, double* v) {
  if (!h->part_overlap) { halo_exchange(h, P, v); return; }
  cudaStream_t compute = h->stream;
  CUDA_OK(cudaEventRecord(h->ev_ready, compute));
  CUDA_OK(cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0));
  h->stream = h->comm_stream;
  try {
    halo_exchange(h, P, v);
  } catch (...) {
    h->stream = compute;
    throw;
  }
  h->stream = compute;
  CUDA_OK(cudaEventRecord(h->ev_done, h->comm_stream));
}
static void halo_exchange_end(H* h) {
  if (!h->part_overlap) return;
  CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_done, 0));
}

static void smooth_part(H* h, Part& P, const SmootherCfg& c) {
  if (c.kind == B200AMG_SMOOTHER_NONE || P.plan.nloc == 0) {
    if (c.kind != B200AMG_SMOOTHER_NONE)
      for (int it = 0; it < c.iter; ++it) { halo_exchange(h, P, P.x); halo_consumed(h); }   // keep the exchanges matched
    return;
  }
  double* cur = P.x;
  double* other = P.temp;
  for (int it = 0; it < c.iter; ++it) {
    halo_exchange_begin(h, P, cur);
    for (int part = 1; part <= 2; ++part) {   // interior rows while the halo is in flight, boundary rows after it has landed
      if (part == 2) halo_exchange_end(h);
      const int sel = h->part_overlap ? part : (part == 2 ? 0 : -1);
      if (sel < 0) continue;
      if (P.symmetry == B200AMG_SYMMETRY_NONE) launch_jacobi_general(h, P.A, P.diag, cur, P.b, other, c.omega, sel);
      else launch_jacobi_fast(h, P.walked(), cur, P.b, other, c.omega, sel);
    }
    halo_consumed(h);
    std::swap(cur, other);
  }
  if (cur != P.x) CUDA_OK(cudaMemcpyAsync(P.x, cur, sizeof(double) * (size_t)P.plan.nloc, cudaMemcpyDeviceToDevice, h->stream));
}

static void solve_level(H* h, double* x, int cycle, const double* b, int lvl, bool x_is_zero);
static void coarse_solve(H* h, double* x, const double* b);

// __solve!(x, ml, cycle, b, lvl) for a level split by rows (multilevel.jl:214-239)
static void cycle_part_level(H* h, int lvl, int cycle) {
  Part& P = *h->parts[lvl];
  const PartPlan& pl = P.plan;
  NcclApi& nc = nccl_api();
  Level& L0 = *h->levels[lvl];
  smooth_part(h, P, P.pre);                                                      // :216
  auto split = [&](auto&& launch) {   // interior part, join the exchange, boundary part (or everything after a blocking exchange)
    if (h->part_overlap) { launch(1); halo_exchange_end(h); launch(2); }
    else launch(0);
    halo_consumed(h);
  };
  halo_exchange_begin(h, P, P.x);
  split([&](int part) { residual(h, P.A, P.x, P.b, P.res, part); });               // :219-220
  halo_exchange_begin(h, P, P.res);
  split([&](int part) { spmv(h, P.R, P.res, P.cb, part); });                       // :223 (my coarse rows)
  if (P.child) {
    // the level below is partitioned too: the restriction wrote straight into its b (the rows I own there)
    Part& C = *P.child;
    CUDA_OK(cudaMemsetAsync(C.x, 0, sizeof(double) * (size_t)std::max<int64_t>(C.plan.nloc, 1), h->stream));   // :226
    if (cycle == B200AMG_CYCLE_V) {
      cycle_part_level(h, lvl + 1, B200AMG_CYCLE_V);
    } else if (cycle == B200AMG_CYCLE_W) {
      cycle_part_level(h, lvl + 1, B200AMG_CYCLE_W);
      cycle_part_level(h, lvl + 1, B200AMG_CYCLE_W);
    } else {
      cycle_part_level(h, lvl + 1, B200AMG_CYCLE_F);
      cycle_part_level(h, lvl + 1, B200AMG_CYCLE_V);
    }
    halo_exchange_begin(h, C, C.x);                                              // my rows of P reach into the neighbours' coarse entries
    split([&](int part) { spmv_add(h, P.P, C.x, P.x, part); });                    // :233-234
    smooth_part(h, P, P.post);                                                   // :236
    return;
  }
  NCCL_OK(nc.GroupStart());                                                      // coarse_b -> rank 0
  if (pl.rank == 0) {
    for (int q = 1; q < pl.world; ++q) {
      const int64_t cnt = pl.coarse_split[q + 1] - pl.coarse_split[q];
      if (cnt > 0) NCCL_OK(nc.Recv(L0.coarse_b + pl.coarse_split[q], (size_t)cnt, ncclDouble, q, h->comm, h->stream));
    }
  } else if (pl.ncloc > 0) {
    NCCL_OK(nc.Send(P.cb, (size_t)pl.ncloc, ncclDouble, 0, h->comm, h->stream));
  }
  NCCL_OK(nc.GroupEnd());
  h->collectives++;
  if (pl.rank == 0) {
    // everything below the partitioned level is a static kernel sequence on this rank: one graph per cycle type
    auto coarse_part = [&]() {
      CUDA_OK(cudaMemsetAsync(L0.coarse_x, 0, sizeof(double) * (size_t)std::max<int64_t>(L0.nc, 1), h->stream));   // :226
      if ((int)h->levels.size() == lvl + 1) {
        coarse_solve(h, L0.coarse_x, L0.coarse_b);                                  // :228
      } else if (cycle == B200AMG_CYCLE_V) {
        solve_level(h, L0.coarse_x, B200AMG_CYCLE_V, L0.coarse_b, lvl + 1, true);
      } else if (cycle == B200AMG_CYCLE_W) {
        solve_level(h, L0.coarse_x, B200AMG_CYCLE_W, L0.coarse_b, lvl + 1, true);
        solve_level(h, L0.coarse_x, B200AMG_CYCLE_W, L0.coarse_b, lvl + 1, false);
      } else {
        solve_level(h, L0.coarse_x, B200AMG_CYCLE_F, L0.coarse_b, lvl + 1, true);
        solve_level(h, L0.coarse_x, B200AMG_CYCLE_V, L0.coarse_b, lvl + 1, false);
      }
    };
    if (h->part_graphs && !h->part_whole_graph && !h->capturing && !h->cycle_graph[cycle] && h->cycle_graph_launches[cycle] >= 0 && !h->profiling) {
      cudaGraph_t gr = nullptr;
      h->capturing = true;
      h->capture_count = 0;
      CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      try {
        coarse_part();
      } catch (...) {
        cudaStreamEndCapture(h->stream, &gr);
        if (gr) cudaGraphDestroy(gr);
        h->capturing = false;
        throw;
      }
      CUDA_OK(cudaStreamEndCapture(h->stream, &gr));
      h->capturing = false;
      CUDA_OK(cudaGraphInstantiate(&h->cycle_graph[cycle], gr, 0));
      CUDA_OK(cudaGraphDestroy(gr));
      h->cycle_graph_launches[cycle] = h->capture_count;
    }
    if (h->part_graphs && !h->part_whole_graph && !h->capturing && h->cycle_graph[cycle]) {
      CUDA_OK(cudaGraphLaunch(h->cycle_graph[cycle], h->stream));
      h->launches += h->cycle_graph_launches[cycle];
    } else {
      coarse_part();
    }
  }
  NCCL_OK(nc.GroupStart());                                                      // coarse_x windows <- rank 0
  if (pl.rank == 0) {
    for (int q = 1; q < pl.world; ++q) {
      const int64_t cnt = pl.cx_hi_all[q] - pl.cx_lo_all[q];
      if (cnt > 0) NCCL_OK(nc.Send(L0.coarse_x + pl.cx_lo_all[q], (size_t)cnt, ncclDouble, q, h->comm, h->stream));
    }
  } else if (pl.cx_hi > pl.cx_lo) {
    NCCL_OK(nc.Recv(P.cx, (size_t)(pl.cx_hi - pl.cx_lo), ncclDouble, 0, h->comm, h->stream));
  }
  NCCL_OK(nc.GroupEnd());
  h->collectives++;
  spmv_add(h, P.P, P.cx, P.x);                                                   // :233-234
  smooth_part(h, P, P.post);                                                     // :236
}
// The whole partitioned cycle — kernels, memsets and the NCCL point-to-point groups of every level — is a static sequence on
// every rank, so it is captured ONCE per cycle type into a CUDA graph and replayed (NCCL >= 2.9 records its kernels into a
// capturing stream): ~60 launches + ~18 communication groups per V-cycle become one graph launch per rank.
// B200AMG_PART_WHOLE_GRAPH=0 (or USE_GRAPHS=0 after finalize) goes back to eager launches.
static void cycle_body_part(H* h, int cycle) {
  if (!h->part_whole_graph || h->profiling) { cycle_part_level(h, 0, cycle); return; }
  if (!h->part_cycle_graph[cycle]) {
    cudaGraph_t g = nullptr;
    const int64_t coll0 = h->collectives, peer0 = h->peer_exchanges;
    h->capturing = true;
    h->capture_count = 0;
    CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    try {
      cycle_part_level(h, 0, cycle);
    } catch (...) {
      cudaStreamEndCapture(h->stream, &g);
      if (g) cudaGraphDestroy(g);
      h->capturing = false;
      throw;
    }
    CUDA_OK(cudaStreamEndCapture(h->stream, &g));
    h->capturing = false;
    CUDA_OK(cudaGraphInstantiate(&h->part_cycle_graph[cycle], g, 0));
    CUDA_OK(cudaGraphDestroy(g));
    h->part_cycle_launches[cycle] = h->capture_count;
    h->part_cycle_collectives[cycle] = h->collectives - coll0;
    h->collectives = coll0;
    h->part_cycle_peer[cycle] = h->peer_exchanges - peer0;
    h->peer_exchanges = peer0;
  }
  CUDA_OK(cudaGraphLaunch(h->part_cycle_graph[cycle], h->stream));
  h->launches += h->part_cycle_launches[cycle];
  h->collectives += h->part_cycle_collectives[cycle];
  h->peer_exchanges += h->part_cycle_peer[cycle];
}

// sum over ranks of a device scalar, in place; every rank gets the same bits
static void allreduce_scalar(H* h, double* dev) {
  NCCL_OK(nccl_api().AllReduce(dev, dev, 1, ncclDouble, ncclSum, h->comm, h->stream));
  h->collectives++;
}
// scalars[slot] = sum over all ranks of v.v over the owned entries (NOT square-rooted)
static void sumsq_part(H* h, const double* v, double* out_dev) {
  dot_partial_kernel<<<kRedBlocks, kThreads, 0, h->stream>>>(h->part->plan.nloc, v, v, h->partial);
  count_launch(h);
  reduce_final_kernel<<<1, kThreads, 0, h->stream>>>(kRedBlocks, h->partial, out_dev, 0);
  count_launch(h);
  allreduce_scalar(h, out_dev);
}
static void residual_norm_part(H* h) {   // scalars[0] = ||b - A x||^2 over all ranks
  Part& P = *h->part;
  halo_exchange(h, P, P.x);   // (blocking here: the kernel below is the one bench.py times on its own)
  const bool timed = h->time_residual && h->res_events_used + 2 <= (int)h->res_events.size();
  if (timed) CUDA_OK(cudaEventRecord(h->res_events[h->res_events_used], h->stream));
  residual(h, P.A, P.x, P.b, P.res);
  if (timed) {
    CUDA_OK(cudaEventRecord(h->res_events[h->res_events_used + 1], h->stream));
    h->res_events_used += 2;
  }
  halo_consumed(h);
  sumsq_part(h, P.res, h->scalars);
}
// owned slice in, assembled vector out
static void part_load(H* h, double* dst, const double* src_full, int memkind) {
  const PartPlan& pl = h->part->plan;
  if (pl.nloc == 0) return;
  CUDA_OK(cudaMemcpyAsync(dst, src_full + pl.row_split[pl.rank], sizeof(double) * (size_t)pl.nloc,
                          memkind == B200AMG_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, h->stream));
}
static void part_store(H* h, double* dst_full, const double* src_local, int memkind) {
  Part& P = *h->part;
  const PartPlan& pl = P.plan;
  NcclApi& nc = nccl_api();
  double* full = dst_full;
  if (memkind == B200AMG_MEM_HOST) {
    if (!P.xfull) P.xfull = dev_alloc<double>(P.n);
    full = P.xfull;
  }
  NCCL_OK(nc.GroupStart());
  for (int q = 0; q < pl.world; ++q) {
    const int64_t cnt = pl.row_split[q + 1] - pl.row_split[q];
    if (cnt > 0) NCCL_OK(nc.Broadcast(q == pl.rank ? src_local : full + pl.row_split[q], full + pl.row_split[q], (size_t)cnt, ncclDouble, q, h->comm, h->stream));
  }
  NCCL_OK(nc.GroupEnd());
  h->collectives++;
  if (memkind == B200AMG_MEM_HOST) CUDA_OK(cudaMemcpyAsync(dst_full, full, sizeof(double) * (size_t)P.n, cudaMemcpyDeviceToHost, h->stream));
}

// ------------------------------------------------------------------------------------------
// API helpers
// ------------------------------------------------------------------------------------------
static void set_device(H* h) { CUDA_OK(cudaSetDevice(h->device)); }
// stream-synchronise and report a sweep kernel whose watchdog fired (a hand-off that never came: the result is not valid)
static void check_coarse_callback(H* h) {
  if (h->coarse_fn_status != 0) {
    const int32_t rc = h->coarse_fn_status;
    h->coarse_fn_status = 0;
    char msg[128];
    std::snprintf(msg, sizeof msg, "the coarse-solver callback returned %d: result discarded", (int)rc);
    throw AmgError{B200AMG_ERR_CALLBACK, msg};
  }
}
static void sync_and_check(H* h) {
  CUDA_OK(cudaStreamSynchronize(h->stream));
  check_coarse_callback(h);
  if (!h->gs_fault) return;
  int f = 0;
  CUDA_OK(cudaMemcpy(&f, h->gs_fault, sizeof(int), cudaMemcpyDeviceToHost));
  if (f) {
    CUDA_OK(cudaMemset(h->gs_fault, 0, sizeof(int)));
    throw AmgError{B200AMG_ERR_CUDA, "a Gauss-Seidel sweep kernel timed out waiting for another tile (watchdog): result discarded"};
  }
}
static void check_ready(H* h) {
  REQUIRE(h, B200AMG_ERR_BAD_ARG, "null handle");
  REQUIRE(h->finalized, B200AMG_ERR_STATE, "hierarchy not finalized (call b200amg_finalize first)");
  set_device(h);
}
// ---- vectors cross the ABI in the caller's (reference) numbering; renumbered levels permute on the way ----
__global__ void __launch_bounds__(kThreads) gather_kernel(int64_t n, const int* __restrict__ idx, const double* __restrict__ src,
                                                          double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
static double* io_scratch(H* h, int64_t n) {
  if (h->io_cap < n) {
    cudaFree(h->io_tmp);
    h->io_tmp = nullptr;
    h->io_cap = 0;
    h->io_tmp = dev_alloc<double>(n + 8);
    h->io_cap = n;
  }
  return h->io_tmp;
}
// dst (device, level numbering) <- src (caller, natural numbering); M == nullptr or identity: plain copy
static void vec_in(H* h, const SmootherMatrix* M, double* dst, const double* src, int64_t n, int memkind) {
  if (n == 0) return;
  const cudaMemcpyKind kind = memkind == B200AMG_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  if (!M || M->perm.identity()) {
    CUDA_OK(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, kind, h->stream));
    return;
  }
  const double* dsrc = src;
  if (memkind == B200AMG_MEM_HOST) {
    double* tmp = io_scratch(h, n);
    CUDA_OK(cudaMemcpyAsync(tmp, src, sizeof(double) * (size_t)n, kind, h->stream));
    dsrc = tmp;
  }
  gather_kernel<<<grid_for(n), kThreads, 0, h->stream>>>(n, M->d_old_of_new, dsrc, dst);   // dst[p] = src[old_of_new[p]]
  count_launch(h);
}
// dst (caller, natural numbering) <- src (device, level numbering)
static void vec_out(H* h, const SmootherMatrix* M, double* dst, const double* src, int64_t n, int memkind) {
  if (n == 0) return;
  const cudaMemcpyKind kind = memkind == B200AMG_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (!M || M->perm.identity()) {
    CUDA_OK(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, kind, h->stream));
    return;
  }
  if (memkind == B200AMG_MEM_HOST) {
    double* tmp = io_scratch(h, n);
    gather_kernel<<<grid_for(n), kThreads, 0, h->stream>>>(n, M->d_new_of_old, src, tmp);   // tmp[i] = src[new_of_old[i]]
    count_launch(h);
    CUDA_OK(cudaMemcpyAsync(dst, tmp, sizeof(double) * (size_t)n, kind, h->stream));
  } else {
    gather_kernel<<<grid_for(n), kThreads, 0, h->stream>>>(n, M->d_new_of_old, src, dst);
    count_launch(h);
  }
}
static const SmootherMatrix* level_numbering(H* h, int level) {   // nullptr: natural numbering
  return level >= 0 && level < (int)h->levels.size() && !h->levels[level]->remote ? &h->levels[level]->M : nullptr;
}

static void check_not_partitioned(H* h, const char* what) {
  REQUIRE(!h->part, B200AMG_ERR_UNSUPPORTED, "%s is not available on a row-partitioned handle (use solve / cycle / precond)", what);
}
static void to_dev(H* h, double* dst, const double* src, int64_t n, int memkind) {
  if (n == 0) return;
  CUDA_OK(cudaMemcpyAsync(dst, src, sizeof(double) * n, memkind == B200AMG_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                          h->stream));
}
static void from_dev(H* h, double* dst, const double* src, int64_t n, int memkind) {
  if (n == 0) return;
  CUDA_OK(cudaMemcpyAsync(dst, src, sizeof(double) * n, memkind == B200AMG_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice,
                          h->stream));
}
// scratch device vector big enough for any level-sized temporary used by the entry points
struct Scratch {
  double* p = nullptr;
  explicit Scratch(int64_t n) { p = dev_alloc<double>(n + 8); }   // +8: TMA row-slice copies round up
  ~Scratch() { cudaFree(p); }
};

// device buffers of one b200amg_spgemm_begin call: freed on every exit path
struct SpgemmDevPool {
  std::vector<void*> p;
  ~SpgemmDevPool() { for (void* q : p) cudaFree(q); }
  template <class T> T* alloc(int64_t count) {
    T* q = nullptr;
    CUDA_OK(cudaMalloc(&q, sizeof(T) * (size_t)std::max<int64_t>(count, 1)));
    p.push_back(q);
    return q;
  }
  template <class T> T* upload(const T* src, int64_t count) {
    T* q = alloc<T>(count);
    if (count) CUDA_OK(cudaMemcpy(q, src, sizeof(T) * (size_t)count, cudaMemcpyHostToDevice));
    return q;
  }
  void release(void* q) {
    cudaFree(q);
    p.erase(std::find(p.begin(), p.end(), q));
  }
};

extern "C" {

const char* b200amg_last_error(void) { return g_err.c_str(); }
int32_t b200amg_version(void) { return B200AMG_VERSION; }
int32_t b200amg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int32_t b200amg_create(b200amg_handle_t* out, int32_t device) {
  API_BEGIN
  REQUIRE(out, B200AMG_ERR_BAD_ARG, "null out pointer");
  *out = nullptr;
  const int ndev = b200amg_device_count();
  REQUIRE(ndev > 0, B200AMG_ERR_NO_DEVICE, "no CUDA device visible: the solve phase has no CPU fallback");
  REQUIRE(device >= 0 && device < ndev, B200AMG_ERR_BAD_ARG, "device %d out of range (0..%d)", device, ndev - 1);
  std::unique_ptr<H> h(new H());
  h->device = device;
  CUDA_OK(cudaSetDevice(device));
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  stream_kernels_init();
  gs_cta_kernels_init();
  dsm_kernels_init();
  gs_block_kernels_init();
  gs_pass_kernels_init();
  CUDA_OK(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  gs_tile_ctas<1>(); gs_tile_ctas<2>(); gs_tile_ctas<4>(); gs_tile_ctas<8>(); gs_tile_ctas<16>(); gs_tile_ctas<32>();
  h->stream_chunk = env_int("B200AMG_STREAM_CHUNK", 4);
  h->gs_mode = env_int("B200AMG_GS_MODE", 2);
  h->gs_acquire = env_int("B200AMG_GS_ACQUIRE", 0);
  h->gs_cta_rows = env_int("B200AMG_GS_CTA_ROWS", 12288);
  h->gs_mail_min_width = env_int("B200AMG_GS_MAIL_MIN_WIDTH", 1024);
  h->gs_tile_any_lanes = env_int("B200AMG_GS_TILE_ANY_LANES", 1);
  h->gs_counter_mail = env_int("B200AMG_GS_COUNTER_MAIL", 1);
  h->gs_cluster = env_int("B200AMG_GS_CLUSTER", 0);
  h->gs_dsm = env_int("B200AMG_GS_DSM", 1);
  h->gs_dsm2 = env_int("B200AMG_GS_DSM2", 1);
  h->gs_dsm_max_log_nc = env_int("B200AMG_GS_DSM_MAX_CTAS_LOG2", h->gs_dsm2 ? 4 : 2);
  h->gs_dsm_fence = env_int("B200AMG_GS_DSM_FENCE", 0);

  h->gs_cluster_rows = env_int("B200AMG_GS_CLUSTER_ROWS", 380000);
  h->gs_poll_sleep = env_int("B200AMG_GS_POLL_SLEEP", 0);
  h->gs_poll_masked = env_int("B200AMG_GS_POLL_MASKED", -1);
  h->gs_tile_cta_limit = env_int("B200AMG_GS_TILE_CTAS", 0);
  h->gs_gate_dist = env_int("B200AMG_GS_GATE_DIST", 2);
  h->gs_gate_sleep = env_int("B200AMG_GS_GATE_SLEEP", 100);
  h->partial = dev_alloc<double>(kRedBlocks);
  h->scalars = dev_alloc<double>(16);
  h->gs_fault = dev_alloc<int>(4);
  CUDA_OK(cudaMemset(h->gs_fault, 0, 4 * sizeof(int)));
  CUDA_OK(cudaMallocHost(&h->h_scalars, sizeof(double) * 16));
  *out = h.release();
  API_END
}

// upload P and R of `prev` once the numbering of the level below it is known (nullptr / identity: unchanged)
static void finish_transfer_operators(Level& prev, const HostPerm* coarse) {
  if (!prev.pending) return;
  if (coarse && !coarse->identity()) {
    map_cols(prev.pendP, *coarse);
    prev.pendR = permute_rows(prev.pendR, *coarse);
  }
  prev.P.upload(prev.pendP);
  prev.R.upload(prev.pendR);
  prev.pendP = HostCsr();
  prev.pendR = HostCsr();
  prev.pending = false;
}

// One partitioned level: local blocks of A, A', R, the halo plan and the local vectors.  P waits (pendP) until the
// next add_level / set_coarse call says whether the level below is partitioned as well.
static void build_part_level(H* h, Level& L, const HostCsr& hAt, HostCsr& hP, const HostCsr& hR, int symmetry) {
  std::unique_ptr<Part> part(new Part());
  Part& P = *part;
  P.level = (int)h->parts.size();
  P.n = L.n; P.nc = L.nc; P.symmetry = symmetry; P.pre = L.pre; P.post = L.post;
  auto ok = [](const SmootherCfg& c) { return c.kind == B200AMG_SMOOTHER_NONE || c.kind == B200AMG_SMOOTHER_JACOBI; };
  REQUIRE(ok(L.pre) && ok(L.post), B200AMG_ERR_UNSUPPORTED,
          "Gauss-Seidel / SOR do not shard (the sweep is sequential over the whole index range): use Jacobi on a partitioned level");
  HostCsr hA = transpose(hAt);
  const bool sym = bit_equal(hA, hAt);
  Part* parent = P.level > 0 ? h->parts[P.level - 1].get() : nullptr;
  if (parent)
    P.plan = make_part_plan(h->rank, h->world, hA, sym ? nullptr : &hAt, hR, hP, &parent->plan.coarse_split, &parent->pendP,
                            &parent->plan.row_split);
  else
    P.plan = make_part_plan(h->rank, h->world, hA, sym ? nullptr : &hAt, hR, hP);
  const PartPlan& pl = P.plan;
  const int64_t lo = pl.row_split[pl.rank], hi = pl.row_split[pl.rank + 1];
  P.A.upload(part_local_block(hA, lo, hi, lo, hi, pl.halo_cols), pl.nloc);
  if (sym) P.At.alias(P.A);
  else P.At.upload(part_local_block(hAt, lo, hi, lo, hi, pl.halo_cols), pl.nloc);
  P.R.upload(part_local_block(hR, pl.coarse_split[pl.rank], pl.coarse_split[pl.rank + 1], lo, hi, pl.halo_cols), pl.nloc);
  {
    const HostCsr& w = symmetry == B200AMG_SYMMETRY_HERMITIAN ? hAt : hA;
    std::vector<double> d((size_t)pl.nloc, 0.0);
    for (int64_t i = lo; i < hi; ++i)
      for (int k = w.ptr[i]; k < w.ptr[i + 1]; ++k)
        if (w.idx[k] == i) d[i - lo] = w.val[k];
    P.diag = dev_upload(d);
  }
  P.send_idx = dev_upload(pl.send_idx);
  P.sendbuf = dev_alloc<double>((int64_t)pl.send_idx.size());
  const int64_t nv = pl.nloc + pl.nhalo + 8;
  P.x = dev_alloc<double>(nv); P.b = dev_alloc<double>(nv); P.res = dev_alloc<double>(nv); P.temp = dev_alloc<double>(nv);
  for (double* v : {P.x, P.b, P.res, P.temp}) CUDA_OK(cudaMemset(v, 0, sizeof(double) * (size_t)nv));
  P.pendP = std::move(hP);
  P.pendingP = true;
  if (parent) {   // the parent restricts into my b and prolongs from my x ([owned | halo] ids)
    const PartPlan& pp = parent->plan;
    parent->child = &P;
    parent->cb = P.b;
    parent->cx = P.x;
    parent->P.upload(part_local_block(parent->pendP, pp.row_split[pp.rank], pp.row_split[pp.rank + 1], lo, hi, pl.halo_cols), pl.nloc);
    parent->pendP = HostCsr();
    parent->pendingP = false;
  }
  h->parts.push_back(std::move(part));
  h->part = h->parts[0].get();
}
// the level below the LAST partitioned level lives on rank 0: coarse_b slices are gathered there, coarse_x windows sent back
static void finish_last_part_level(H* h) {
  if (h->parts.empty() || !h->parts.back()->pendingP) return;
  Part& P = *h->parts.back();
  Level& L = *h->levels[P.level];
  const PartPlan& pl = P.plan;
  if (pl.rank == 0) {   // my coarse rows / window are slices of the full vectors
    L.coarse_x = dev_alloc<double>(L.nc + 8);
    L.coarse_b = dev_alloc<double>(L.nc + 8);
    P.cb = L.coarse_b + pl.coarse_split[0];
    P.cx = L.coarse_x + pl.cx_lo;
  } else {
    P.cb = dev_alloc<double>(pl.ncloc + 8); P.own_cb = true;
    P.cx = dev_alloc<double>(pl.cx_hi - pl.cx_lo + 8); P.own_cx = true;
  }
  P.P.upload(part_shifted_block(P.pendP, pl.row_split[pl.rank], pl.row_split[pl.rank + 1], pl.cx_lo, pl.cx_hi - pl.cx_lo));
  P.pendP = HostCsr();
  P.pendingP = false;
}

int32_t b200amg_add_level(b200amg_handle_t h, const b200amg_csc_t* A, const b200amg_csc_t* P, const b200amg_csc_t* R,
                          const b200amg_smoother_t* pre, const b200amg_smoother_t* post, int32_t symmetry) {
  API_BEGIN
  REQUIRE(h && A && P && R, B200AMG_ERR_BAD_ARG, "null argument");
  REQUIRE(!h->finalized, B200AMG_ERR_STATE, "hierarchy already finalized");
  REQUIRE(symmetry == B200AMG_SYMMETRY_HERMITIAN || symmetry == B200AMG_SYMMETRY_NONE, B200AMG_ERR_BAD_ARG, "bad symmetry tag");
  REQUIRE(A->m == A->n, B200AMG_ERR_DIM_MISMATCH, "A must be square (%lld x %lld)", (long long)A->m, (long long)A->n);
  set_device(h);
  std::unique_ptr<Level> L(new Level());
  L->n = A->n;
  L->pre = to_cfg(pre);
  L->post = to_cfg(post);
  if (!h->levels.empty())
    REQUIRE(h->levels.back()->nc == L->n, B200AMG_ERR_DIM_MISMATCH, "level has %lld rows but the previous level coarsens to %lld",
            (long long)L->n, (long long)h->levels.back()->nc);
  const bool fine_of_partition = h->world > 1 && (int)h->levels.size() < h->part_levels;   // this level is split by rows
  if (h->world > 1 && !fine_of_partition) finish_last_part_level(h);
  const bool remote = h->world > 1 && !fine_of_partition && h->rank != 0;   // the other levels live on rank 0 only
  {
    HostCsr hP, hR;
    { UploadTimer t("stage P, R by rows"); hP = stage_operator_by_rows(P); hR = stage_operator_by_rows(R); }
    REQUIRE(hP.nrows == L->n, B200AMG_ERR_DIM_MISMATCH, "P has %lld rows, A has %lld", (long long)hP.nrows, (long long)L->n);
    REQUIRE(hR.ncols == L->n, B200AMG_ERR_DIM_MISMATCH, "R has %lld columns, A has %lld", (long long)hR.ncols, (long long)L->n);
    REQUIRE(hR.nrows == hP.ncols, B200AMG_ERR_DIM_MISMATCH, "R has %lld rows but P has %lld columns", (long long)hR.nrows,
            (long long)hP.ncols);
    L->nc = hR.nrows;
    L->nnz_p = hP.nnz();
    HostCsr hAt;
    { UploadTimer t("stage A"); hAt = stage_csc_as_rows_of_transpose(A); }
    L->nnz_a = hAt.nnz();
    UploadTimer t_level("level total (after staging)");
    if (fine_of_partition) {
      L->remote = true;   // no full device copy of this level on any rank
      build_part_level(h, *L, hAt, hP, hR, symmetry);
    } else if (remote) {
      L->remote = true;
    } else {
      L->M.build(hAt, symmetry, cfg_needs_fwd(L->pre) || cfg_needs_fwd(L->post), cfg_needs_bwd(L->pre) || cfg_needs_bwd(L->post), true);
      REQUIRE(!(h->part && h->levels.size() == h->parts.size() && !L->M.perm.identity()), B200AMG_ERR_UNSUPPORTED,
              "the level below a partitioned fine level must use Jacobi smoothing (its numbering is shared with the other ranks)");
      // this level's numbering: rows of P, columns of R now; the coarse side when the next level arrives
      L->pendP = permute_rows(hP, L->M.perm);
      map_cols(hR, L->M.perm);
      L->pendR = std::move(hR);
      L->pending = true;
      L->res = dev_alloc<double>(L->n + 8);
      L->temp = dev_alloc<double>(L->n + 8);
      L->coarse_x = dev_alloc<double>(L->nc + 8);
      L->coarse_b = dev_alloc<double>(L->nc + 8);
    }
  }
  if (!h->levels.empty()) finish_transfer_operators(*h->levels.back(), L->remote ? nullptr : &L->M.perm);
  h->levels.push_back(std::move(L));
  API_END
}

static void set_coarse_impl(H* h, const b200amg_csc_t* final_A, int64_t n, const double* inv, b200amg_coarse_fn fn, void* user) {
  REQUIRE(h && final_A, B200AMG_ERR_BAD_ARG, "null argument");
  REQUIRE(!h->finalized, B200AMG_ERR_STATE, "hierarchy already finalized");
  REQUIRE(final_A->m == n && final_A->n == n, B200AMG_ERR_DIM_MISMATCH, "final_A is %lld x %lld, coarse operator is %lld",
          (long long)final_A->m, (long long)final_A->n, (long long)n);
  REQUIRE(n == 0 || inv || fn, B200AMG_ERR_BAD_ARG, "null coarse operator");
  REQUIRE(fn || n <= 16384, B200AMG_ERR_UNSUPPORTED,
          "dense coarse operator limited to 16384 rows (got %lld): use b200amg_set_coarse_callback for a larger coarsest level", (long long)n);
  if (!h->levels.empty())
    REQUIRE(h->levels.back()->nc == n, B200AMG_ERR_DIM_MISMATCH, "coarsest matrix has %lld rows, last level coarsens to %lld",
            (long long)n, (long long)h->levels.back()->nc);
  set_device(h);
  REQUIRE(h->world == 1 || !h->levels.empty(), B200AMG_ERR_UNSUPPORTED, "a partitioned hierarchy needs at least one level");
  if (!h->levels.empty()) finish_transfer_operators(*h->levels.back(), nullptr);   // the coarsest level keeps its numbering
  if (h->world > 1) finish_last_part_level(h);
  h->nfinal = n;
  if (h->world == 1 || h->rank == 0) {
    HostCsr hAt = stage_csc_as_rows_of_transpose(final_A);
    h->finalA.upload(transpose(hAt));
    if (fn) {
      h->coarse_fn = fn;
      h->coarse_user = user;
      CUDA_OK(cudaMallocHost(&h->coarse_hb, sizeof(double) * (size_t)std::max<int64_t>(n, 1)));
      CUDA_OK(cudaMallocHost(&h->coarse_hx, sizeof(double) * (size_t)std::max<int64_t>(n, 1)));
    } else {
      std::vector<double> m(inv, inv + n * n);
      h->coarse_inv = dev_upload(m);
    }
    h->res_final = dev_alloc<double>(n);
  }
  h->have_coarse = true;
}

int32_t b200amg_set_coarse(b200amg_handle_t h, const b200amg_csc_t* final_A, int64_t n, const double* inv) {
  API_BEGIN
  set_coarse_impl(h, final_A, n, inv, nullptr, nullptr);
  API_END
}

int32_t b200amg_set_coarse_callback(b200amg_handle_t h, const b200amg_csc_t* final_A, int64_t n, b200amg_coarse_fn fn, void* user) {
  API_BEGIN
  REQUIRE(fn, B200AMG_ERR_BAD_ARG, "null coarse-solver callback");
  set_coarse_impl(h, final_A, n, nullptr, fn, user);
  API_END
}

int32_t b200amg_set_partition(b200amg_handle_t h, int32_t rank, int32_t world_size, const void* id, int64_t id_bytes) {
  API_BEGIN
  REQUIRE(h, B200AMG_ERR_BAD_ARG, "null handle");
  REQUIRE(!h->finalized && h->levels.empty(), B200AMG_ERR_STATE, "set_partition must precede add_level");
  REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, B200AMG_ERR_BAD_ARG, "bad rank %d / world size %d", rank, world_size);
  if (world_size > 1) {
    REQUIRE(id && id_bytes == (int64_t)sizeof(ncclUniqueId), B200AMG_ERR_BAD_ARG, "nccl unique id must be %d bytes", (int)sizeof(ncclUniqueId));
    nccl_api();   // fail early if NCCL cannot be loaded
    std::memcpy(&h->nccl_id, id, sizeof(ncclUniqueId));
  }
  h->rank = rank;
  h->world = world_size;
  h->part_levels = std::max(1, env_int("B200AMG_PART_LEVELS", h->part_levels));
  h->part_whole_graph = env_int("B200AMG_PART_WHOLE_GRAPH", 1) != 0;
  API_END
}

int32_t b200amg_nccl_unique_id(void* out, int64_t cap) {
  API_BEGIN
  REQUIRE(out && cap >= (int64_t)sizeof(ncclUniqueId), B200AMG_ERR_BAD_ARG, "buffer must hold %d bytes", (int)sizeof(ncclUniqueId));
  ncclUniqueId id;
  NCCL_OK(nccl_api().GetUniqueId(&id));
  std::memcpy(out, &id, sizeof id);
  API_END
}

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

int32_t b200amg_partition_info(b200amg_handle_t h, int64_t* row_lo, int64_t* row_hi, int64_t* nhalo, int64_t* nsend,
                               int64_t* coarse_lo, int64_t* coarse_hi, int64_t* cx_lo, int64_t* cx_hi) {
  API_BEGIN
  REQUIRE(h && h->part, B200AMG_ERR_STATE, "handle is not partitioned");
  const PartPlan& pl = h->part->plan;
  if (row_lo) *row_lo = pl.row_split[pl.rank];
  if (row_hi) *row_hi = pl.row_split[pl.rank + 1];
  if (nhalo) *nhalo = pl.nhalo;
  if (nsend) *nsend = (int64_t)pl.send_idx.size();
  if (coarse_lo) *coarse_lo = pl.coarse_split[pl.rank];
  if (coarse_hi) *coarse_hi = pl.coarse_split[pl.rank + 1];
  if (cx_lo) *cx_lo = pl.cx_lo;
  if (cx_hi) *cx_hi = pl.cx_hi;
  API_END
}

// Export my exchangeable vectors and flag words, map the neighbours' (CUDA IPC over NVLink), build the per-channel tables of
// peer_halo.cuh.  Collective: every rank calls it at finalize; the path is switched on only if EVERY rank succeeded.
struct PeerBlob {
  cudaIpcMemHandle_t sync;
  cudaIpcMemHandle_t buf[kPeerMaxLevels][kPeerChannels];
  int recv_off[kPeerMaxLevels][kPeerMaxWorld + 1];
  long long nloc[kPeerMaxLevels];
  int ok;
};
static void peer_setup(H* h) {
  const int world = h->world, me = h->rank, nl = (int)h->parts.size();
  int ok = env_int("B200AMG_PEER_HALO", 1) != 0 && world <= kPeerMaxWorld && nl <= kPeerMaxLevels;
  std::vector<PeerBlob> all((size_t)world);
  PeerBlob mine;
  memset(&mine, 0, sizeof mine);
  h->peer.sync = dev_alloc<unsigned long long>(kPeerSyncWords);
  CUDA_OK(cudaMemset(h->peer.sync, 0, sizeof(unsigned long long) * kPeerSyncWords));
  h->peer.tickets = dev_alloc<unsigned>(kPeerMaxLevels * kPeerChannels);
  CUDA_OK(cudaMemset(h->peer.tickets, 0, sizeof(unsigned) * kPeerMaxLevels * kPeerChannels));
  if (ok) {
    ok = cudaIpcGetMemHandle(&mine.sync, h->peer.sync) == cudaSuccess;
    for (int l = 0; l < nl && ok; ++l) {
      Part& P = *h->parts[(size_t)l];
      double* bufs[kPeerChannels] = {P.x, P.temp, P.res};
      for (int c = 0; c < kPeerChannels && ok; ++c) ok = cudaIpcGetMemHandle(&mine.buf[l][c], bufs[c]) == cudaSuccess;
      for (int q = 0; q <= world; ++q) mine.recv_off[l][q] = P.plan.recv_off[(size_t)q];
      mine.nloc[l] = P.plan.nloc;
    }
    (void)cudaGetLastError();
  }
  mine.ok = ok;
  {   // all-gather of the blobs (bytes) over the communicator that already exists
    unsigned char* d_all = dev_alloc<unsigned char>((int64_t)sizeof(PeerBlob) * world);
    CUDA_OK(cudaMemcpy(d_all + sizeof(PeerBlob) * (size_t)me, &mine, sizeof mine, cudaMemcpyHostToDevice));
    NCCL_OK(nccl_api().AllGather(d_all + sizeof(PeerBlob) * (size_t)me, d_all, sizeof(PeerBlob), ncclChar, h->comm, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaMemcpy(all.data(), d_all, sizeof(PeerBlob) * (size_t)world, cudaMemcpyDeviceToHost));
    cudaFree(d_all);
  }
  for (int q = 0; q < world; ++q) ok = ok && all[(size_t)q].ok;
  std::vector<PeerTables> tabs((size_t)(kPeerMaxLevels * kPeerChannels));
  memset(tabs.data(), 0, sizeof(PeerTables) * tabs.size());
  if (ok) {
    std::vector<unsigned long long*> rsync((size_t)world, nullptr);
    auto open = [&](const cudaIpcMemHandle_t& hd) -> void* {
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); ok = 0; return nullptr; }
      h->peer.opened.push_back(p);
      return p;
    };
    for (int l = 0; l < nl && ok; ++l) {
      const PartPlan& pl = h->parts[(size_t)l]->plan;
      for (int q = 0; q < world && ok; ++q) {
        if (q == me) continue;
        const bool sends = pl.send_off[(size_t)q + 1] > pl.send_off[(size_t)q], recvs = pl.recv_off[(size_t)q + 1] > pl.recv_off[(size_t)q];
        if (!sends && !recvs) continue;
        if (!rsync[(size_t)q]) rsync[(size_t)q] = (unsigned long long*)open(all[(size_t)q].sync);
        if (!ok) break;
        for (int c = 0; c < kPeerChannels && ok; ++c) {
          PeerTables& T = tabs[(size_t)(l * kPeerChannels + c)];
          unsigned long long* words = rsync[(size_t)q] + (size_t)(l * kPeerChannels + c) * kPeerWords;
          T.flag_at[q] = words + me;
          T.ack_at[q] = words + kPeerMaxWorld + me;
          if (sends) {
            double* base = (double*)open(all[(size_t)q].buf[l][c]);
            if (!ok) break;
            T.dst[q] = base + all[(size_t)q].nloc[l] + all[(size_t)q].recv_off[l][me];
            // what I send must be exactly what q expects from me
            if (all[(size_t)q].recv_off[l][me + 1] - all[(size_t)q].recv_off[l][me] != pl.send_off[(size_t)q + 1] - pl.send_off[(size_t)q]) ok = 0;
          }
        }
      }
      for (int c = 0; c < kPeerChannels; ++c) {
        PeerTables& T = tabs[(size_t)(l * kPeerChannels + c)];
        T.world = world;
        for (int q = 0; q <= world; ++q) T.send_off[q] = pl.send_off[(size_t)q];
        for (int q = 0; q < world; ++q) T.recv_cnt[q] = pl.recv_off[(size_t)q + 1] - pl.recv_off[(size_t)q];
      }
    }
  }
  {   // agree: all ranks or none
    int* d_ok = dev_alloc<int>(2);
    CUDA_OK(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
    NCCL_OK(nccl_api().AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, h->comm, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_ok);
  }
  if (ok) {
    h->peer.d_tab = dev_alloc<PeerTables>((int64_t)tabs.size());
    CUDA_OK(cudaMemcpy(h->peer.d_tab, tabs.data(), sizeof(PeerTables) * tabs.size(), cudaMemcpyHostToDevice));
  }
  h->peer.on = ok != 0;
  if (env_int("B200AMG_VERBOSE_UPLOAD", 0) || env_int("B200AMG_PEER_VERBOSE", 0))
    fprintf(stderr, "[b200amg] rank %d: halo exchange over %s\n", me, h->peer.on ? "peer memory (CUDA IPC, direct stores into the neighbours' halos)" : "NCCL send/recv");
}

// does launch_sweep take the TMA-fed mailbox sweep (gs_tile_kernel) for this matrix?  (mirrors its selection)
static bool sweep_uses_tile_kernel(const H* h, const SmootherMatrix& M) {
  if (M.pass.ok || M.block.ok || M.n <= 0 || M.nlev <= 0) return false;
  const bool wide = M.n / M.nlev >= h->gs_mail_min_width;
  if (!wide) return false;
  if (h->gs_mode >= 1 && M.n <= h->gs_cta_rows && M.walked().ntiles > 0 && M.d_fwd_lvlptr) return false;
  return h->gs_mode == 2 && M.mail && M.gs_ntiles > 0 && (M.gs_lanes == 1 || h->gs_tile_any_lanes);
}

// How many persistent CTAs the mailbox sweep of a level gets.  MORE tiles in flight is not better: CTAs that hold tiles
// several wavefronts ahead of the sweep's front only poll (issue slots and L2 bandwidth taken from the tiles on the critical
// path of their SM).  Measured on B200, 256^3 RS hierarchy, SGS ms with 296 / 222 / 148 / 74 CTAs: level 0 (one lane per row)
// 3.61 / 3.03 / 3.4 / 5.6, level 1 (4 lanes) 7.30 / 6.57 / 7.04 / 11.6, level 2 (8 lanes) 5.57 / 5.49 / 5.29 / 4.69
// (profiles/r02_tile_cta_limit_256_scan.log) - the best count depends on the level, so it is MEASURED here, once per level,
// on the level's own (zeroed) vectors.  The result of a sweep does not depend on it.  B200AMG_GS_TILE_TUNE=0 switches it off.
static void tune_tile_ctas(H* h) {
  if (h->world != 1 || !env_int("B200AMG_GS_TILE_TUNE", 1) || h->gs_tile_cta_limit > 0) return;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  for (size_t lvl = 0; lvl < h->levels.size(); ++lvl) {
    Level& L = *h->levels[lvl];
    const bool gs_pre = L.pre.kind == B200AMG_SMOOTHER_GS || L.pre.kind == B200AMG_SMOOTHER_SOR;
    const bool gs_post = L.post.kind == B200AMG_SMOOTHER_GS || L.post.kind == B200AMG_SMOOTHER_SOR;
    if (!(gs_pre || gs_post) || !sweep_uses_tile_kernel(h, L.M)) continue;
    double* x = lvl == 0 ? h->x0 : h->levels[lvl - 1]->coarse_x;
    double* b = lvl == 0 ? h->b0 : h->levels[lvl - 1]->coarse_b;
    if (!x || !b || !L.temp) continue;
    if (!e0) { CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1)); }
    const SmootherCfg& cfg = gs_pre ? L.pre : L.post;
    CUDA_OK(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)L.n, h->stream));
    CUDA_OK(cudaMemsetAsync(b, 0, sizeof(double) * (size_t)L.n, h->stream));
    const int full = std::min(L.M.gs_ntiles, 2 * h->num_sms);
    int best = 0;
    float best_ms = 0.f;
    int worse_in_a_row = 0;
    for (int eighths = 8; eighths >= 1 && worse_in_a_row < 2; --eighths) {
      const int ctas = std::max(1, full * eighths / 8);
      L.M.gs_tile_ctas = ctas;
      float ms = 0.f;
      for (int rep = 0; rep < 3; ++rep) {   // first one untimed
        if (rep == 1) CUDA_OK(cudaEventRecord(e0, h->stream));
        smooth(h, L.M, cfg, x, b, L.temp, false);
      }
      CUDA_OK(cudaEventRecord(e1, h->stream));
      CUDA_OK(cudaEventSynchronize(e1));
      CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
      if (best == 0 || ms < best_ms) { best = ctas; best_ms = ms; worse_in_a_row = 0; }
      else ++worse_in_a_row;
    }
    L.M.gs_tile_ctas = best;
    if (env_int("B200AMG_GS_TILE_TUNE_VERBOSE", 0))
      std::fprintf(stderr, "[b200amg] level %zu: mailbox sweep on %d of %d CTAs (%.3f ms per smoother call)\n", lvl, best, full, best_ms / 2);
  }
  if (e0) { cudaEventDestroy(e0); cudaEventDestroy(e1); }
  CUDA_OK(cudaStreamSynchronize(h->stream));
  sync_and_check(h);
}

int32_t b200amg_finalize(b200amg_handle_t h) {
  API_BEGIN
  REQUIRE(h, B200AMG_ERR_BAD_ARG, "null handle");
  REQUIRE(!h->finalized, B200AMG_ERR_STATE, "hierarchy already finalized");
  REQUIRE(h->have_coarse, B200AMG_ERR_STATE, "b200amg_set_coarse has not been called");
  set_device(h);
  h->n0 = h->levels.empty() ? h->nfinal : h->levels[0]->n;
  if (h->world > 1) {
    REQUIRE(h->part, B200AMG_ERR_STATE, "partitioned hierarchy without a fine level");
    NCCL_OK(nccl_api().CommInitRank(&h->comm, h->world, h->nccl_id, h->rank));
    h->use_graphs = false;
    h->part_overlap = env_int("B200AMG_PART_OVERLAP", 1) != 0;
    CUDA_OK(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
    peer_setup(h);
  } else {
    h->x0 = dev_alloc<double>(h->n0 + 8);
    h->b0 = dev_alloc<double>(h->n0 + 8);
    CUDA_OK(cudaMemset(h->x0, 0, sizeof(double) * (size_t)(h->n0 + 8)));
    CUDA_OK(cudaMemset(h->b0, 0, sizeof(double) * (size_t)(h->n0 + 8)));
  }
  h->finalized = true;
  tune_tile_ctas(h);
  API_END
}

int32_t b200amg_destroy(b200amg_handle_t h) {
  if (!h) return B200AMG_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& L : h->levels) {
    L->M.release(); L->P.release(); L->R.release();
    cudaFree(L->res); cudaFree(L->temp); cudaFree(L->coarse_x); cudaFree(L->coarse_b);
  }
  for (auto& pp : h->parts) pp->release();
  // graphs that hold NCCL kernels go first: the communicator cannot be torn down while captured work still refers to it
  for (int c = 0; c < 3; ++c) {
    if (h->part_cycle_graph[c]) cudaGraphExecDestroy(h->part_cycle_graph[c]);
    h->part_cycle_graph[c] = nullptr;
  }
  if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
  for (void* p : h->peer.opened) cudaIpcCloseMemHandle(p);
  h->peer.opened.clear();
  cudaFree(h->peer.sync); cudaFree(h->peer.tickets); cudaFree(h->peer.d_tab);
  if (h->comm) nccl_api().CommDestroy(h->comm);
  if (h->ev_ready) cudaEventDestroy(h->ev_ready);
  if (h->ev_done) cudaEventDestroy(h->ev_done);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  h->finalA.release();
  cudaFree(h->coarse_inv); cudaFree(h->res_final); cudaFree(h->x0); cudaFree(h->b0);
  cudaFree(h->partial); cudaFree(h->scalars); cudaFree(h->gs_fault); cudaFreeHost(h->h_scalars);
  cudaFreeHost(h->coarse_hb); cudaFreeHost(h->coarse_hx);
  cudaFree(h->pcg_u); cudaFree(h->pcg_q); cudaFree(h->pcg_x); cudaFree(h->flush); cudaFree(h->io_tmp);
  for (cudaEvent_t e : h->res_events) cudaEventDestroy(e);
  for (int c = 0; c < 3; ++c)
    if (h->cycle_graph[c]) cudaGraphExecDestroy(h->cycle_graph[c]);
  if (h->resnorm_graph) cudaGraphExecDestroy(h->resnorm_graph);
  for (int c = 0; c < 3; ++c)
    if (h->part_cycle_graph[c]) cudaGraphExecDestroy(h->part_cycle_graph[c]);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return B200AMG_OK;
}

// _solve!  — src/multilevel.jl:158-198
int32_t b200amg_solve(b200amg_handle_t h, double* x, const double* b, int32_t cycle, int32_t maxiter, double abstol,
                      double reltol, int32_t calculate_residual, double* residuals, int32_t cap, int32_t* nres,
                      int32_t* iters, int32_t memkind) {
  API_BEGIN
  check_ready(h);
  REQUIRE(x && b, B200AMG_ERR_BAD_ARG, "null vector");
  REQUIRE(cycle >= 0 && cycle <= 2, B200AMG_ERR_BAD_ARG, "unknown cycle %d", cycle);
  const int64_t n = h->n0;
  if (h->part) {
    part_load(h, h->part->b, b, memkind);
    part_load(h, h->part->x, x, memkind);
  } else {
    vec_in(h, level_numbering(h, 0), h->b0, b, n, memkind);
    vec_in(h, level_numbering(h, 0), h->x0, x, n, memkind);
  }
  int nr = 0;
  h->res_events_used = 0;
  if (h->time_residual) {
    const size_t want = 2 * (size_t)std::min(std::max(maxiter, 0), 2048);
    while (h->res_events.size() < want) {
      cudaEvent_t e;
      CUDA_OK(cudaEventCreate(&e));
      h->res_events.push_back(e);
    }
  }
  double normb;
  if (h->part) {
    sumsq_part(h, h->part->b, h->scalars);
    normb = std::sqrt(read_scalar(h, h->scalars));
  } else {
    norm2_async(h, n, h->b0, h->scalars);
    normb = read_scalar(h, h->scalars);
  }
  double normres = normb;                                                              // :170
  if (normb != 0) abstol = std::max(reltol * normb, abstol);                           // :171-173
  if (residuals && nr < cap) residuals[nr++] = normb;                                  // :174
  int itr = 1;
  while (itr <= maxiter && (!calculate_residual || normres > abstol)) {                // :178
    run_cycle(h, cycle);                                                               // :179-183
    if (calculate_residual) {
      if (h->part) {
        residual_norm_part(h);
        normres = std::sqrt(read_scalar(h, h->scalars));
      } else {
        residual_norm(h);                                                              // :188-190
        normres = read_scalar(h, h->scalars);
      }
      if (residuals && nr < cap) residuals[nr++] = normres;                            // :191
    }
    itr += 1;
  }
  if (h->part) part_store(h, x, h->part->x, memkind);
  else vec_out(h, level_numbering(h, 0), x, h->x0, n, memkind);
  sync_and_check(h);
  if (nres) *nres = nr;
  if (iters) *iters = itr - 1;
  API_END
}

// _solve!(x, ml, b, ...) for MATRIX right-hand sides (the reference's block workspaces, src/multilevel.jl:28-59): x, b are
// n x ncols, column-major with leading dimension ld (a Julia Matrix).  The reference relaxes / restricts / prolongs column by
// column (src/smoother.jl:77,118,195; stdlib mul! over columns) and tests ONE norm over all columns (Frobenius:
// multilevel.jl:170,190).  All columns stay on the device for the whole call: per iteration every column runs the captured
// cycle graph on the level-0 work vectors (device-to-device copies in and out), the residual columns are formed there, and
// only ncols sums of squares cross PCIe per iteration.
int32_t b200amg_solve_block(b200amg_handle_t h, double* x, const double* b, int64_t ncols, int64_t ld, int32_t cycle, int32_t maxiter,
                            double abstol, double reltol, int32_t calculate_residual, double* residuals, int32_t cap, int32_t* nres,
                            int32_t* iters, int32_t memkind) {
  API_BEGIN
  check_ready(h);
  check_not_partitioned(h, "solve_block");
  REQUIRE(x && b, B200AMG_ERR_BAD_ARG, "null vector");
  REQUIRE(cycle >= 0 && cycle <= 2, B200AMG_ERR_BAD_ARG, "unknown cycle %d", cycle);
  const int64_t n = h->n0;
  REQUIRE(ncols >= 1 && ld >= n, B200AMG_ERR_DIM_MISMATCH, "bad block shape (%lld columns, leading dimension %lld, n = %lld)",
          (long long)ncols, (long long)ld, (long long)n);
  const int64_t stride = (n + 8 + 1) & ~(int64_t)1;   // device column stride (16-byte aligned columns, room for the TMA slack)
  Scratch X(stride * ncols), B(stride * ncols), SS(ncols + 8);
  const SmootherMatrix* num = level_numbering(h, 0);
  for (int64_t j = 0; j < ncols; ++j) {
    vec_in(h, num, B.p + j * stride, b + j * ld, n, memkind);
    vec_in(h, num, X.p + j * stride, x + j * ld, n, memkind);
  }
  std::vector<double> ss((size_t)ncols);
  auto frobenius = [&](auto&& column) {   // sqrt of the sum over columns of ||column(j)||^2, columns added in order
    for (int64_t j = 0; j < ncols; ++j) dot_async(h, n, column(j), column(j), SS.p + j);
    CUDA_OK(cudaMemcpyAsync(ss.data(), SS.p, sizeof(double) * (size_t)ncols, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    double t = 0.0;
    for (int64_t j = 0; j < ncols; ++j) t += ss[(size_t)j];
    return std::sqrt(t);
  };
  int nr = 0;
  const double normb = frobenius([&](int64_t j) { return (const double*)(B.p + j * stride); });
  double normres = normb;                                                              // :170
  if (normb != 0) abstol = std::max(reltol * normb, abstol);                           // :171-173
  if (residuals && nr < cap) residuals[nr++] = normb;                                  // :174
  const DevCsr& A = h->levels.empty() ? h->finalA : h->levels[0]->M.A;
  double* res = h->levels.empty() ? h->res_final : h->levels[0]->res;
  Scratch R(calculate_residual ? stride * ncols : 1);
  int itr = 1;
  while (itr <= maxiter && (!calculate_residual || normres > abstol)) {                // :178
    for (int64_t j = 0; j < ncols; ++j) {
      CUDA_OK(cudaMemcpyAsync(h->x0, X.p + j * stride, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
      CUDA_OK(cudaMemcpyAsync(h->b0, B.p + j * stride, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
      run_cycle(h, cycle);                                                             // :179-183
      CUDA_OK(cudaMemcpyAsync(X.p + j * stride, h->x0, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
    }
    if (calculate_residual) {
      for (int64_t j = 0; j < ncols; ++j) residual(h, A, X.p + j * stride, B.p + j * stride, R.p + j * stride);   // :188-189
      normres = frobenius([&](int64_t j) { return (const double*)(R.p + j * stride); });                         // :190
      if (residuals && nr < cap) residuals[nr++] = normres;                            // :191
    }
    itr += 1;
  }
  (void)res;
  for (int64_t j = 0; j < ncols; ++j) vec_out(h, num, x + j * ld, X.p + j * stride, n, memkind);
  sync_and_check(h);
  if (nres) *nres = nr;
  if (iters) *iters = itr - 1;
  API_END
}

int32_t b200amg_cycle(b200amg_handle_t h, double* x, const double* b, int32_t cycle, int32_t memkind) {
  API_BEGIN
  check_ready(h);
  REQUIRE(x && b, B200AMG_ERR_BAD_ARG, "null vector");
  REQUIRE(cycle >= 0 && cycle <= 2, B200AMG_ERR_BAD_ARG, "unknown cycle %d", cycle);
  if (h->part) {
    part_load(h, h->part->b, b, memkind);
    part_load(h, h->part->x, x, memkind);
    run_cycle(h, cycle);
    part_store(h, x, h->part->x, memkind);
  } else {
    vec_in(h, level_numbering(h, 0), h->b0, b, h->n0, memkind);
    vec_in(h, level_numbering(h, 0), h->x0, x, h->n0, memkind);
    run_cycle(h, cycle);
    vec_out(h, level_numbering(h, 0), x, h->x0, h->n0, memkind);
  }
  sync_and_check(h);
  API_END
}

// ldiv!(x, p, b)  — src/preconditioner.jl:12-19
int32_t b200amg_precond(b200amg_handle_t h, double* x, const double* b, int32_t cycle, int32_t init_zero, int32_t memkind) {
  API_BEGIN
  check_ready(h);
  REQUIRE(x && b, B200AMG_ERR_BAD_ARG, "null vector");
  REQUIRE(cycle >= 0 && cycle <= 2, B200AMG_ERR_BAD_ARG, "unknown cycle %d", cycle);
  if (h->part) {
    Part& P = *h->part;
    part_load(h, P.b, b, memkind);
    if (init_zero) CUDA_OK(cudaMemsetAsync(P.x, 0, sizeof(double) * (size_t)std::max<int64_t>(P.plan.nloc, 1), h->stream));
    else CUDA_OK(cudaMemcpyAsync(P.x, P.b, sizeof(double) * (size_t)P.plan.nloc, cudaMemcpyDeviceToDevice, h->stream));
    run_cycle(h, cycle);
    part_store(h, x, P.x, memkind);
    sync_and_check(h);
    return B200AMG_OK;
  }
  vec_in(h, level_numbering(h, 0), h->b0, b, h->n0, memkind);
  if (init_zero) CUDA_OK(cudaMemsetAsync(h->x0, 0, sizeof(double) * (size_t)std::max<int64_t>(h->n0, 1), h->stream));
  else CUDA_OK(cudaMemcpyAsync(h->x0, h->b0, sizeof(double) * h->n0, cudaMemcpyDeviceToDevice, h->stream));
  run_cycle(h, cycle);
  vec_out(h, level_numbering(h, 0), x, h->x0, h->n0, memkind);
  sync_and_check(h);
  API_END
}

int32_t b200amg_smooth(b200amg_handle_t h, int32_t level, int32_t which, double* x, const double* b, int32_t memkind) {
  API_BEGIN
  check_ready(h);
  check_not_partitioned(h, "smooth");
  REQUIRE(level >= 0 && level < (int)h->levels.size(), B200AMG_ERR_BAD_ARG, "level %d out of range", level);
  REQUIRE(x && b, B200AMG_ERR_BAD_ARG, "null vector");
  Level& L = *h->levels[level];
  Scratch sx(L.n), sb(L.n);
  vec_in(h, &L.M, sx.p, x, L.n, memkind);
  vec_in(h, &L.M, sb.p, b, L.n, memkind);
  smooth(h, L.M, which == B200AMG_PRE ? L.pre : L.post, sx.p, sb.p, L.temp, false);
  vec_out(h, &L.M, x, sx.p, L.n, memkind);
  sync_and_check(h);
  API_END
}

int32_t b200amg_apply(b200amg_handle_t h, int32_t level, int32_t op, double* y, const double* x, int32_t memkind) {
  API_BEGIN
  check_ready(h);
  check_not_partitioned(h, "apply");
  REQUIRE(y && x, B200AMG_ERR_BAD_ARG, "null vector");
  const int nl = (int)h->levels.size();
  const DevCsr* A = nullptr;
  const SmootherMatrix *num_in = nullptr, *num_out = nullptr;   // numbering of the input / output vector
  if (level == nl && op == B200AMG_OP_A) A = &h->finalA;
  else {
    REQUIRE(level >= 0 && level < nl, B200AMG_ERR_BAD_ARG, "level %d out of range", level);
    Level& L = *h->levels[level];
    A = op == B200AMG_OP_A ? &L.M.A : op == B200AMG_OP_P ? &L.P : op == B200AMG_OP_R ? &L.R : nullptr;
    REQUIRE(A, B200AMG_ERR_BAD_ARG, "unknown operator %d", op);
    const SmootherMatrix* fine = level_numbering(h, level);
    const SmootherMatrix* coarse = level_numbering(h, level + 1);
    num_in = op == B200AMG_OP_A ? fine : op == B200AMG_OP_P ? coarse : fine;
    num_out = op == B200AMG_OP_A ? fine : op == B200AMG_OP_P ? fine : coarse;
  }
  Scratch sx(A->ncols), sy(A->nrows);
  vec_in(h, num_in, sx.p, x, A->ncols, memkind);
  spmv(h, *A, sx.p, sy.p);
  vec_out(h, num_out, y, sy.p, A->nrows, memkind);
  CUDA_OK(cudaStreamSynchronize(h->stream));
  API_END
}

int32_t b200amg_residual(b200amg_handle_t h, int32_t level, double* r, const double* b, const double* x, int32_t memkind) {
  API_BEGIN
  check_ready(h);
  check_not_partitioned(h, "residual");
  REQUIRE(r && b && x, B200AMG_ERR_BAD_ARG, "null vector");
  const int nl = (int)h->levels.size();
  REQUIRE(level >= 0 && level <= nl, B200AMG_ERR_BAD_ARG, "level %d out of range", level);
  const DevCsr& A = level == nl ? h->finalA : h->levels[level]->M.A;
  Scratch sx(A.ncols), sb(A.nrows), sr(A.nrows);
  const SmootherMatrix* num = level_numbering(h, level);
  vec_in(h, num, sx.p, x, A.ncols, memkind);
  vec_in(h, num, sb.p, b, A.nrows, memkind);
  residual(h, A, sx.p, sb.p, sr.p);
  vec_out(h, num, r, sr.p, A.nrows, memkind);
  CUDA_OK(cudaStreamSynchronize(h->stream));
  API_END
}

int32_t b200amg_coarse_solve(b200amg_handle_t h, double* x, const double* b, int32_t memkind) {
  API_BEGIN
  check_ready(h);
  check_not_partitioned(h, "coarse_solve");
  REQUIRE(x && b, B200AMG_ERR_BAD_ARG, "null vector");
  Scratch sx(h->nfinal), sb(h->nfinal);
  to_dev(h, sb.p, b, h->nfinal, memkind);
  coarse_solve(h, sx.p, sb.p);
  from_dev(h, x, sx.p, h->nfinal, memkind);
  sync_and_check(h);
  API_END
}

int32_t b200amg_norm(b200amg_handle_t h, int64_t n, const double* v, double* out, int32_t memkind) {
  API_BEGIN
  check_ready(h);
  REQUIRE(v && out && n >= 0, B200AMG_ERR_BAD_ARG, "bad argument");
  if (memkind == B200AMG_MEM_HOST) {
    Scratch s(n);
    to_dev(h, s.p, v, n, memkind);
    norm2_async(h, n, s.p, h->scalars);
    *out = read_scalar(h, h->scalars);
  } else {
    norm2_async(h, n, v, h->scalars);
    *out = read_scalar(h, h->scalars);
  }
  API_END
}

// Device-resident left-preconditioned CG (IterativeSolvers' PCGIterable with Pl = one AMG cycle):
//   c = Pl \ r ; rho = c.r ; u = c + (rho/rho_prev) u ; q = A u ; alpha = rho/(u.q) ; x += alpha u ; r -= alpha q
// r lives in b0 and c in x0, so the preconditioner is zero-fill + the captured cycle graph with no copies.
int32_t b200amg_pcg(b200amg_handle_t h, double* x, const double* b, int32_t cycle, int32_t maxiter, double abstol,
                    double reltol, double* residuals, int32_t cap, int32_t* nres, int32_t* iters, int32_t memkind) {
  API_BEGIN
  check_ready(h);
  REQUIRE(x && b, B200AMG_ERR_BAD_ARG, "null vector");
  REQUIRE(cycle >= 0 && cycle <= 2, B200AMG_ERR_BAD_ARG, "unknown cycle %d", cycle);
  if (h->part) {
    // Row-partitioned handle: every rank holds its row block of x, r, u, q; z = Pl \ r is the partitioned cycle; q = A u needs u's
    // halo (u travels in the level's exported `res` vector); dots = local partial sums + one all-reduce of a double (the same
    // bits on every rank, so every rank takes the same loop decisions).  Same recurrences as the single-GPU loop below.
    Part& P = *h->part;
    const int64_t nl = P.plan.nloc, nalloc = std::max<int64_t>(nl, 1);
    if (!h->pcg_u) { h->pcg_u = dev_alloc<double>(nalloc); h->pcg_q = dev_alloc<double>(nalloc); h->pcg_x = dev_alloc<double>(nalloc); }
    double* S = h->scalars;
    auto dot_part = [&](const double* u, const double* v, double* out) {
      dot_partial_kernel<<<kRedBlocks, kThreads, 0, h->stream>>>(nl, u, v, h->partial);
      count_launch(h);
      reduce_final_kernel<<<1, kThreads, 0, h->stream>>>(kRedBlocks, h->partial, out, 0);
      count_launch(h);
      allreduce_scalar(h, out);
    };
    part_load(h, P.b, b, memkind);                                                         // r = b (x starts at zero)
    CUDA_OK(cudaMemsetAsync(h->pcg_x, 0, sizeof(double) * (size_t)nalloc, h->stream));
    CUDA_OK(cudaMemsetAsync(h->pcg_u, 0, sizeof(double) * (size_t)nalloc, h->stream));
    set_scalar_kernel<<<1, 32, 0, h->stream>>>(S + 1, 1.0);
    count_launch(h);
    dot_part(P.b, P.b, S);
    double residual = std::sqrt(read_scalar(h, S));
    const double tol = std::max(reltol * residual, abstol);
    int nr = 0, it = 0;
    if (residuals && nr < cap) residuals[nr++] = residual;
    const unsigned g = grid_for(nalloc);
    while (!(it >= maxiter || residual <= tol)) {
      CUDA_OK(cudaMemsetAsync(P.x, 0, sizeof(double) * (size_t)nalloc, h->stream));         // ldiv!: x .= 0
      run_cycle(h, cycle);                                                                 // c = Pl \ r   (c == P.x, r == P.b)
      copy_scalar_kernel<<<1, 32, 0, h->stream>>>(S + 2, S + 1);
      count_launch(h);
      dot_part(P.x, P.b, S + 1);                                                           // rho = c.r
      if (nl > 0) {
        pcg_update_u_kernel<<<g, kThreads, 0, h->stream>>>(nl, P.x, h->pcg_u, S + 1, S + 2);
        count_launch(h);
        CUDA_OK(cudaMemcpyAsync(P.res, h->pcg_u, sizeof(double) * (size_t)nl, cudaMemcpyDeviceToDevice, h->stream));
      }
      halo_exchange(h, P, P.res);
      spmv(h, P.A, P.res, h->pcg_q);                                                       // q = A u
      halo_consumed(h);
      dot_part(h->pcg_u, h->pcg_q, S + 3);                                                 // u.q
      if (nl > 0) {
        pcg_update_xr_kernel<<<g, kThreads, 0, h->stream>>>(nl, h->pcg_x, P.b, h->pcg_u, h->pcg_q, S + 1, S + 3);
        count_launch(h);
      }
      dot_part(P.b, P.b, S);
      residual = std::sqrt(read_scalar(h, S));
      if (residuals && nr < cap) residuals[nr++] = residual;
      ++it;
    }
    part_store(h, x, h->pcg_x, memkind);
    sync_and_check(h);
    if (nres) *nres = nr;
    if (iters) *iters = it;
    return B200AMG_OK;
  }
  const int64_t n = h->n0;
  const DevCsr& A = h->levels.empty() ? h->finalA : h->levels[0]->M.A;
  if (!h->pcg_u) { h->pcg_u = dev_alloc<double>(n); h->pcg_q = dev_alloc<double>(n); h->pcg_x = dev_alloc<double>(n); }
  double* S = h->scalars;  // [0] norm [1] rho [2] rho_prev [3] uq
  vec_in(h, level_numbering(h, 0), h->b0, b, n, memkind);                                 // r = b (x starts at zero)
  CUDA_OK(cudaMemsetAsync(h->pcg_x, 0, sizeof(double) * (size_t)std::max<int64_t>(n, 1), h->stream));
  CUDA_OK(cudaMemsetAsync(h->pcg_u, 0, sizeof(double) * (size_t)std::max<int64_t>(n, 1), h->stream));
  set_scalar_kernel<<<1, 32, 0, h->stream>>>(S + 1, 1.0);
  count_launch(h);
  norm2_async(h, n, h->b0, S);
  double residual = read_scalar(h, S);
  const double tol = std::max(reltol * residual, abstol);
  int nr = 0, it = 0;
  if (residuals && nr < cap) residuals[nr++] = residual;
  const unsigned g = grid_for(n);
  while (!(it >= maxiter || residual <= tol)) {
    CUDA_OK(cudaMemsetAsync(h->x0, 0, sizeof(double) * (size_t)std::max<int64_t>(n, 1), h->stream));  // ldiv!: x .= 0
    run_cycle(h, cycle);                                                                 // c = Pl \ r   (c == x0, r == b0)
    copy_scalar_kernel<<<1, 32, 0, h->stream>>>(S + 2, S + 1);
    count_launch(h);
    dot_async(h, n, h->x0, h->b0, S + 1);                                                // rho = c.r
    pcg_update_u_kernel<<<g, kThreads, 0, h->stream>>>(n, h->x0, h->pcg_u, S + 1, S + 2);
    count_launch(h);
    spmv(h, A, h->pcg_u, h->pcg_q);                                                      // q = A u
    dot_async(h, n, h->pcg_u, h->pcg_q, S + 3);                                          // u.q
    pcg_update_xr_kernel<<<g, kThreads, 0, h->stream>>>(n, h->pcg_x, h->b0, h->pcg_u, h->pcg_q, S + 1, S + 3);
    count_launch(h);
    norm2_async(h, n, h->b0, S);
    residual = read_scalar(h, S);
    if (residuals && nr < cap) residuals[nr++] = residual;
    ++it;
  }
  vec_out(h, level_numbering(h, 0), x, h->pcg_x, n, memkind);
  sync_and_check(h);
  if (nres) *nres = nr;
  if (iters) *iters = it;
  API_END
}

// ---- standalone smoothers -----------------------------------------------------------------
struct b200amg_smoother_obj {
  H* h = nullptr;   // private mini-handle (stream, counters)
  SmootherMatrix M;
  SmootherCfg cfg;
  double *x = nullptr, *b = nullptr, *temp = nullptr;
};

int32_t b200amg_smoother_create(b200amg_smoother_handle_t* out, int32_t device, const b200amg_csc_t* A,
                                const b200amg_smoother_t* config, int32_t symmetry) {
  API_BEGIN
  REQUIRE(out && A && config, B200AMG_ERR_BAD_ARG, "null argument");
  *out = nullptr;
  REQUIRE(A->m == A->n, B200AMG_ERR_DIM_MISMATCH, "A must be square");
  REQUIRE(symmetry == B200AMG_SYMMETRY_HERMITIAN || symmetry == B200AMG_SYMMETRY_NONE, B200AMG_ERR_BAD_ARG, "bad symmetry tag");
  b200amg_handle_t hh = nullptr;
  int32_t rc = b200amg_create(&hh, device);
  if (rc) throw AmgError{rc, g_err};
  std::unique_ptr<b200amg_smoother_obj> s(new b200amg_smoother_obj());
  s->h = hh;
  try {
    s->cfg = to_cfg(config);
    HostCsr hAt = stage_csc_as_rows_of_transpose(A);
    s->M.build(hAt, symmetry, cfg_needs_fwd(s->cfg), cfg_needs_bwd(s->cfg), false);
    s->x = dev_alloc<double>(A->n + 8);
    s->b = dev_alloc<double>(A->n + 8);
    s->temp = dev_alloc<double>(A->n + 8);
  } catch (...) {
    s->M.release(); cudaFree(s->x); cudaFree(s->b); cudaFree(s->temp);
    b200amg_destroy(hh);
    throw;
  }
  *out = s.release();
  API_END
}

int32_t b200amg_smoother_apply(b200amg_smoother_handle_t s, double* x, const double* b, int32_t memkind) {
  API_BEGIN
  REQUIRE(s && x && b, B200AMG_ERR_BAD_ARG, "null argument");
  H* h = s->h;
  set_device(h);
  vec_in(h, &s->M, s->x, x, s->M.n, memkind);
  vec_in(h, &s->M, s->b, b, s->M.n, memkind);
  smooth(h, s->M, s->cfg, s->x, s->b, s->temp, false);
  vec_out(h, &s->M, x, s->x, s->M.n, memkind);
  sync_and_check(h);
  API_END
}

int32_t b200amg_smoother_destroy(b200amg_smoother_handle_t s) {
  if (!s) return B200AMG_OK;
  cudaSetDevice(s->h->device);
  cudaStreamSynchronize(s->h->stream);
  s->M.release();
  cudaFree(s->x); cudaFree(s->b); cudaFree(s->temp);
  b200amg_destroy(s->h);
  delete s;
  return B200AMG_OK;
}

// ---- introspection / measurement ----------------------------------------------------------
int32_t b200amg_num_levels(b200amg_handle_t h) { return h ? (int32_t)h->levels.size() + 1 : 0; }

int32_t b200amg_level_info(b200amg_handle_t h, int32_t level, int64_t* n, int64_t* nnz_a, int64_t* nnz_p, int64_t* wavefronts) {
  API_BEGIN
  REQUIRE(h, B200AMG_ERR_BAD_ARG, "null handle");
  const int nl = (int)h->levels.size();
  REQUIRE(level >= 0 && level <= nl, B200AMG_ERR_BAD_ARG, "level %d out of range", level);
  if (level == nl) {
    if (n) *n = h->nfinal;
    if (nnz_a) *nnz_a = h->finalA.nnz;
    if (nnz_p) *nnz_p = 0;
    if (wavefronts) *wavefronts = 0;
  } else {
    Level& L = *h->levels[level];
    if (n) *n = L.n;
    if (nnz_a) *nnz_a = L.nnz_a;
    if (nnz_p) *nnz_p = L.nnz_p;
    if (wavefronts) *wavefronts = L.M.block.ok ? L.M.block.wavefronts : (L.M.fwd.built ? L.M.fwd.nlev : (L.M.bwd.built ? L.M.bwd.nlev : 0));
  }
  API_END
}

// bytes per stored VALUE the bandwidth kernels read on a level: out[0] A, out[1] P, out[2] R (4 = the lossless binary32 copy is
// in use, 8 = fp64; 0 = the operator does not exist on this rank)
int32_t b200amg_storage_info(b200amg_handle_t h, int32_t level, int32_t* out, int32_t cap) {
  API_BEGIN
  REQUIRE(h && out && cap >= 3, B200AMG_ERR_BAD_ARG, "bad argument");
  const int nl = (int)h->levels.size();
  REQUIRE(level >= 0 && level <= nl, B200AMG_ERR_BAD_ARG, "level %d out of range", level);
  auto vb = [&](const DevCsr& M) { return M.nnz == 0 && !M.val ? 0 : (M.val32 && h->fp32_storage ? 4 : 8); };
  out[0] = out[1] = out[2] = 0;
  if (level == nl) {
    out[0] = vb(h->finalA);
  } else if (level < (int)h->parts.size()) {
    const Part& P = *h->parts[(size_t)level];
    out[0] = vb(P.A); out[1] = vb(P.P); out[2] = vb(P.R);
  } else {
    const Level& L = *h->levels[(size_t)level];
    out[0] = vb(L.M.A); out[1] = vb(L.P); out[2] = vb(L.R);
  }
  API_END
}

int64_t b200amg_launch_count(b200amg_handle_t h) { return h ? h->launches : 0; }

// counters of a partitioned handle: [0] NCCL groups / collectives enqueued so far, [1] halo exchanges over peer memory so far,
// [2] 1 if the halo exchanges go over peer memory (CUDA IPC), 0 if over NCCL send/recv, [3] partitioned levels
int32_t b200amg_comm_stats(b200amg_handle_t h, int64_t* out, int32_t cap) {
  API_BEGIN
  REQUIRE(h && out && cap >= 4, B200AMG_ERR_BAD_ARG, "bad argument");
  out[0] = h->collectives;
  out[1] = h->peer_exchanges;
  out[2] = h->peer.on ? 1 : 0;
  out[3] = (int64_t)h->parts.size();
  API_END
}

int32_t b200amg_time_kernel(b200amg_handle_t h, int32_t level, int32_t what, int32_t cycle, int32_t reps, int32_t flush_l2,
                            double* ms) {
  API_BEGIN
  check_ready(h);
  REQUIRE(ms && reps > 0, B200AMG_ERR_BAD_ARG, "bad argument");
  const int nl = (int)h->levels.size();
  REQUIRE(what == 5 || what == 6 || what == 7 || (level >= 0 && level < nl), B200AMG_ERR_BAD_ARG, "level %d out of range", level);
  REQUIRE(what != 7 || h->part, B200AMG_ERR_BAD_ARG, "selector 7 (halo exchange) needs a partitioned handle");
  if (flush_l2 && !h->flush) {
    h->flush_bytes = (size_t)512 << 20;
    CUDA_OK(cudaMalloc(&h->flush, h->flush_bytes));
  }
  if (what == 5) ensure_cycle_graph(h, cycle);
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  double total = 0.0;
  for (int r = 0; r < reps; ++r) {
    if (flush_l2) CUDA_OK(cudaMemsetAsync(h->flush, r & 0xff, h->flush_bytes, h->stream));
    CUDA_OK(cudaEventRecord(e0, h->stream));
    if (what == 5) {
      run_cycle(h, cycle);
    } else if (h->part) {
      // per-rank view of the partitioned fine level: 0 local SpMV, 1 local residual, 2 pre-smoother
      // (halo exchanges included), 6 local sum of squares + all-reduce, 7 one halo exchange
      Part& P = *h->part;
      REQUIRE(level == 0 || what == 6 || what == 7, B200AMG_ERR_UNSUPPORTED, "only the fine level of a partitioned handle can be timed");
      switch (what) {
        case 0: spmv(h, P.A, P.x, P.res); break;
        case 1: residual(h, P.A, P.x, P.b, P.res); break;
        case 2: smooth_part(h, P, P.pre); break;
        case 3: spmv(h, P.R, P.res, P.cb); break;
        case 4: spmv_add(h, P.P, P.cx, P.x); break;
        case 6: sumsq_part(h, P.b, h->scalars + 4); break;
        case 7: halo_exchange(h, P, P.x); halo_consumed(h); break;
        default: REQUIRE(false, B200AMG_ERR_BAD_ARG, "unknown kernel selector %d", what);
      }
    } else if (what == 6) {
      norm2_async(h, h->n0, h->b0, h->scalars + 4);
    } else {
      Level& L = *h->levels[level];
      // level > 0: the level's x / b are the parent's coarse_x / coarse_b
      double* x = level == 0 ? h->x0 : h->levels[level - 1]->coarse_x;
      const double* b = level == 0 ? h->b0 : h->levels[level - 1]->coarse_b;
      switch (what) {
        case 0: spmv(h, L.M.A, x, L.res); break;
        case 1: residual(h, L.M.A, x, b, L.res); break;
        case 2: smooth(h, L.M, L.pre, x, b, L.temp, false); break;
        case 3: spmv(h, L.R, L.res, L.coarse_b); break;
        case 4: spmv_add(h, L.P, L.coarse_x, x); break;
        default: REQUIRE(false, B200AMG_ERR_BAD_ARG, "unknown kernel selector %d", what);
      }
    }
    CUDA_OK(cudaEventRecord(e1, h->stream));
    CUDA_OK(cudaEventSynchronize(e1));
    float t = 0;
    CUDA_OK(cudaEventElapsedTime(&t, e0, e1));
    total += t;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms = total / reps;
  API_END
}

int32_t b200amg_profile_cycle(b200amg_handle_t h, int32_t cycle, double* ms, int32_t cap) {
  API_BEGIN
  check_ready(h);
  check_not_partitioned(h, "profile_cycle");
  REQUIRE(ms && cap >= 6 * ((int)h->levels.size() + 1), B200AMG_ERR_BAD_ARG, "ms buffer too small");
  std::vector<double> acc((size_t)cap, 0.0);
  h->profiling = true;
  h->prof_ms = &acc;
  try {
    cycle_body(h, cycle, false);
    CUDA_OK(cudaStreamSynchronize(h->stream));
  } catch (...) {
    h->profiling = false; h->prof_ms = nullptr;
    throw;
  }
  h->profiling = false;
  h->prof_ms = nullptr;
  for (int i = 0; i < cap; ++i) ms[i] = acc[i];
  API_END
}

// graphs captured with the old option values would keep launching the old kernels: drop them, they are re-captured on demand
static void drop_cycle_graphs(H* h) {
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (int c = 0; c < 3; ++c) {
    if (h->cycle_graph[c]) cudaGraphExecDestroy(h->cycle_graph[c]);
    h->cycle_graph[c] = nullptr;
    h->cycle_graph_launches[c] = 0;
  }
  if (h->resnorm_graph) { cudaGraphExecDestroy(h->resnorm_graph); h->resnorm_graph = nullptr; }
  for (int c = 0; c < 3; ++c) {
    if (h->part_cycle_graph[c]) cudaGraphExecDestroy(h->part_cycle_graph[c]);
    h->part_cycle_graph[c] = nullptr;
  }
}

int32_t b200amg_set_option(b200amg_handle_t h, int32_t option, double value) {
  API_BEGIN
  REQUIRE(h, B200AMG_ERR_BAD_ARG, "null handle");
  if (option != B200AMG_OPT_TIME_RESIDUAL && option != B200AMG_OPT_USE_GRAPHS && option != B200AMG_OPT_PART_LEVELS) drop_cycle_graphs(h);
  switch (option) {
    case B200AMG_OPT_USE_GRAPHS:
      h->use_graphs = value != 0.0;
      if (h->part) h->part_whole_graph = value != 0.0;   // (finalize switches use_graphs off for partitioned handles)
      break;
    case B200AMG_OPT_TIME_RESIDUAL: h->time_residual = value != 0.0; break;
    case B200AMG_OPT_STREAM_CHUNK: h->stream_chunk = (int)value; break;
    case B200AMG_OPT_GS_MODE: h->gs_mode = (int)value; break;
    case B200AMG_OPT_GS_ACQUIRE: h->gs_acquire = (int)value; break;
    case B200AMG_OPT_GS_POLL_SLEEP: h->gs_poll_sleep = (int)value; break;
    case B200AMG_OPT_GS_CTA_ROWS: h->gs_cta_rows = (int64_t)value; break;
    case B200AMG_OPT_GS_MAIL_MIN_WIDTH: h->gs_mail_min_width = (int64_t)value; break;
    case B200AMG_OPT_GS_CLUSTER: h->gs_cluster = (int)value; break;
    case B200AMG_OPT_GS_DSM: h->gs_dsm = (int)value; break;
    case B200AMG_OPT_GS_DSM_FENCE: h->gs_dsm_fence = (int)value; break;
    case B200AMG_OPT_GS_DSM_MAX_CTAS_LOG2: h->gs_dsm_max_log_nc = (int)value; break;
    case B200AMG_OPT_FP32_STORAGE: h->fp32_storage = value != 0; break;
    case 16: h->gs_poll_masked = (int)value; break;   // experiment knobs (tools/tune_kernels.py, tools/tile_knobs.py)
    case 19: h->gs_tile_cta_limit = (int)value; break;
    case 20: h->gs_gate_dist = (int)value; break;
    case B200AMG_OPT_GS_DSM2: h->gs_dsm2 = (int)value; break;
    case B200AMG_OPT_PART_LEVELS:
      REQUIRE(h->levels.empty(), B200AMG_ERR_STATE, "PART_LEVELS must be set before the first add_level");
      h->part_levels = std::max(1, (int)value);
      break;
    case 10: h->gs_cluster_log_nc = (int)value; break;     // experiment knobs (tools/tune_kernels.py)
    case 11: h->gs_cluster_threads = (int)value; break;
    case B200AMG_OPT_GS_GATE_SLEEP: h->gs_gate_sleep = (int)value; break;
    default: REQUIRE(false, B200AMG_ERR_BAD_ARG, "unknown option %d", option);
  }
  API_END
}

int32_t b200amg_residual_timings(b200amg_handle_t h, double* ms, int32_t cap, int32_t* n) {
  API_BEGIN
  REQUIRE(h && ms && n, B200AMG_ERR_BAD_ARG, "null argument");
  set_device(h);
  int k = 0;
  for (int i = 0; i + 1 < h->res_events_used && k < cap; i += 2, ++k) {
    float t = 0;
    CUDA_OK(cudaEventSynchronize(h->res_events[i + 1]));
    CUDA_OK(cudaEventElapsedTime(&t, h->res_events[i], h->res_events[i + 1]));
    ms[k] = t;
  }
  *n = k;
  API_END
}

int32_t b200amg_debug_gs_timeline(b200amg_handle_t h, int32_t level, int32_t backward, uint64_t* out, int64_t cap,
                                  int64_t* ntasks) {
  API_BEGIN
  check_ready(h);
  check_not_partitioned(h, "gs timeline");
  REQUIRE(level >= 0 && level < (int)h->levels.size() && out && ntasks, B200AMG_ERR_BAD_ARG, "bad argument");
  Level& L = *h->levels[level];
  const DevSchedule& sc = backward ? L.M.bwd : L.M.fwd;
  if (L.M.block.ok) {   // blocked sweep: 8 words per tile (block_gs.cuh)
    const int64_t words = (int64_t)L.M.block.ntiles * 8;
    REQUIRE(cap >= words, B200AMG_ERR_BAD_ARG, "timeline buffer too small (%lld needed)", (long long)words);
    unsigned long long* d = dev_alloc<unsigned long long>(words);
    CUDA_OK(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * (size_t)words, h->stream));
    h->gs_debug = d;
    double* xb = level == 0 ? h->x0 : h->levels[level - 1]->coarse_x;
    const double* bb = level == 0 ? h->b0 : h->levels[level - 1]->coarse_b;
    if (L.M.pass.ok) launch_gs_pass(h, L.M, sc, xb, bb, L.pre.omega, L.pre.kind == B200AMG_SMOOTHER_SOR);
    else launch_gs_block(h, L.M, L.M.walked(), sc, xb, bb, L.pre.omega, L.pre.kind == B200AMG_SMOOTHER_SOR);
    h->gs_debug = nullptr;
    CUDA_OK(cudaMemcpyAsync(out, d, sizeof(unsigned long long) * (size_t)words, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    cudaFree(d);
    *ntasks = L.M.block.ntiles;
    return B200AMG_OK;
  }
  REQUIRE(sc.built, B200AMG_ERR_STATE, "no Gauss-Seidel schedule on this level");
  const bool wide = sc.nlev > 0 && L.M.n / sc.nlev >= h->gs_mail_min_width;
  const bool dsm = !wide && h->gs_dsm && L.M.dsm_ntiles > 0 && L.M.dsm_log_nc <= h->gs_dsm_max_log_nc;   // 8 stamps per tile (dsm_gs.cuh)
  const bool tile = !dsm && wide && h->gs_mode == 2 && L.M.mail && L.M.gs_ntiles > 0;                    // 8 stamps per tile (gs_tile_kernel)
  const int64_t words = (int64_t)(dsm ? L.M.dsm_ntiles : tile ? L.M.gs_ntiles : sc.ntasks) * 8;
  REQUIRE(cap >= words, B200AMG_ERR_BAD_ARG, "timeline buffer too small (%lld needed)", (long long)words);
  unsigned long long* d = dev_alloc<unsigned long long>(words);
  CUDA_OK(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * (size_t)words, h->stream));
  h->gs_debug = d;
  const int sor = L.pre.kind == B200AMG_SMOOTHER_SOR;
  double* x = level == 0 ? h->x0 : h->levels[level - 1]->coarse_x;
  const double* b = level == 0 ? h->b0 : h->levels[level - 1]->coarse_b;
  if (dsm) REQUIRE(launch_gs_dsm(h, L.M, L.M.walked(), sc, x, b, L.pre.omega, sor), B200AMG_ERR_CUDA, "dsm sweep not launchable");
  else if (tile) launch_gs_tile(h, L.M, L.M.walked(), sc, x, b, L.pre.omega, sor);
  else launch_dataflow(h, L.M.walked(), sc, x, b, L.pre.omega, sor);
  h->gs_debug = nullptr;
  CUDA_OK(cudaMemcpyAsync(out, d, sizeof(unsigned long long) * (size_t)words, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  cudaFree(d);
  *ntasks = dsm ? L.M.dsm_ntiles : tile ? L.M.gs_ntiles : sc.ntasks;
  API_END
}

// ------------------------------------------------------------------------------------------
// device Galerkin product (spgemm.cuh): C = A * B, CSC int32 0-based in and out — the same contract as the host
// restatement's amgsetup_spgemm_begin / _fetch, bit-identical results.  Not re-entrant (one pending result).
// ------------------------------------------------------------------------------------------
struct SpgemmPending {
  int64_t n = 0;
  std::vector<int> colcount;
  std::vector<int> rows;
  std::vector<double> vals;
};
static SpgemmPending* g_spgemm_pending = nullptr;
// scratch of the hash tables and the per-batch output: kept between calls (a setup phase multiplies 2 x levels times),
// grown on demand, released by b200amg_spgemm_release or at process exit
struct SpgemmScratch {
  int device = -1;
  long long slots = 0;
  int* keys = nullptr;
  double* vals = nullptr;
  int* Cj = nullptr;
  double* Cx = nullptr;
  void release() {
    if (device >= 0) cudaSetDevice(device);
    cudaFree(keys); cudaFree(vals); cudaFree(Cj); cudaFree(Cx);
    keys = nullptr; vals = nullptr; Cj = nullptr; Cx = nullptr;
    slots = 0;
    device = -1;
  }
  void ensure(int dev, long long want) {
    if (device == dev && slots >= want) return;
    release();
    CUDA_OK(cudaSetDevice(dev));
    device = dev;
    CUDA_OK(cudaMalloc(&keys, sizeof(int) * (size_t)want));
    CUDA_OK(cudaMalloc(&vals, sizeof(double) * (size_t)want));
    CUDA_OK(cudaMalloc(&Cj, sizeof(int) * (size_t)(want / 2 + 1)));
    CUDA_OK(cudaMalloc(&Cx, sizeof(double) * (size_t)(want / 2 + 1)));
    slots = want;
  }
};
static SpgemmScratch g_spgemm_scratch;

int32_t b200amg_spgemm_begin(int32_t device, int64_t m, int64_t k, int64_t n, const int32_t* Ap, const int32_t* Aj, const double* Ax,
                             const int32_t* Bp, const int32_t* Bj, const double* Bx, int64_t* nnz_out) {
  API_BEGIN
  REQUIRE(Ap && Bp && nnz_out && m >= 0 && k >= 0 && n >= 0, B200AMG_ERR_BAD_ARG, "bad argument");
  const int ndev = b200amg_device_count();
  REQUIRE(ndev > 0, B200AMG_ERR_NO_DEVICE, "no CUDA device visible: the device Galerkin product has no CPU fallback");
  REQUIRE(device >= 0 && device < ndev, B200AMG_ERR_BAD_ARG, "device %d out of range (0..%d)", device, ndev - 1);
  CUDA_OK(cudaSetDevice(device));
  delete g_spgemm_pending;
  g_spgemm_pending = nullptr;
  std::unique_ptr<SpgemmPending> R(new SpgemmPending());
  R->n = n;
  R->colcount.assign((size_t)n, 0);
  const int64_t nnzA = Ap[k], nnzB = Bp[n];
  REQUIRE(nnzA == 0 || (Aj && Ax), B200AMG_ERR_BAD_ARG, "null A arrays");
  REQUIRE(nnzB == 0 || (Bj && Bx), B200AMG_ERR_BAD_ARG, "null B arrays");
  SpgemmDevPool D;
  UploadTimer t_all("spgemm total");
  double t_mark = UploadTimer::now(), t_up = 0.0, t_hash = 0.0, t_emit = 0.0;
  auto lap = [&](double& acc) { const double t = UploadTimer::now(); acc += t - t_mark; t_mark = t; };
  int* dAp = D.upload(Ap, k + 1);
  int* dAj = D.upload(Aj, nnzA);
  double* dAx = D.upload(Ax, nnzA);
  int* dBp = D.upload(Bp, n + 1);
  int* dBj = D.upload(Bj, nnzB);
  double* dBx = D.upload(Bx, nnzB);
  lap(t_up);
  // ---- products per column -> table capacities ----
  std::vector<long long> prod((size_t)n, 0);
  if (n) {
    long long* dprod = D.alloc<long long>(n);
    spgemm_products_kernel<<<(unsigned)((n + 255) / 256), 256>>>(n, dAp, dBp, dBj, dprod);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpy(prod.data(), dprod, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost));
  }
  const long long slot_budget = (long long)env_int("B200AMG_SPGEMM_SLOTS_M", 192) << 20;   // table slots per batch (12 bytes each)
  std::vector<int> cap((size_t)n, 0);
  long long max_cap = 0;
  for (int64_t j = 0; j < n; ++j) {
    if (prod[j] == 0) continue;
    long long c = 4;
    while (c < 2 * prod[j]) c <<= 1;
    REQUIRE(c <= (1ll << 30), B200AMG_ERR_UNSUPPORTED, "a column of the product has %lld partial products", prod[j]);
    cap[j] = (int)c;
    max_cap = std::max(max_cap, c);
  }
  REQUIRE(max_cap <= slot_budget, B200AMG_ERR_UNSUPPORTED, "one column needs %lld table slots (budget %lld)", max_cap, slot_budget);
  long long total_cap = 0;
  for (int64_t j = 0; j < n; ++j) total_cap += cap[j];
  const long long slot_alloc = std::max<long long>(std::min(slot_budget, total_cap), 2);   // small products: small scratch
  // ---- batches of columns that fit the scratch budget ----
  g_spgemm_scratch.ensure(device, slot_alloc);     // distinct rows <= products <= capacity / 2: the output needs half the slots
  int* dkeys = g_spgemm_scratch.keys;
  double* dvals = g_spgemm_scratch.vals;
  int* dCj = g_spgemm_scratch.Cj;
  double* dCx = g_spgemm_scratch.Cx;
  int64_t j0 = 0, total = 0;
  std::vector<long long> off, cptr;
  std::vector<int> ucount;
  while (j0 < n) {
    int64_t j1 = j0;
    long long slots = 0;
    off.clear();
    while (j1 < n && slots + cap[j1] <= slot_budget && j1 - j0 < (1 << 24)) {
      off.push_back(slots);
      slots += cap[j1];
      ++j1;
    }
    const int64_t count = j1 - j0;
    lap(t_up);
    long long* doff = D.upload(off.data(), count);
    int* dcap = D.upload(cap.data() + j0, count);
    int* du = D.alloc<int>(count);
    CUDA_OK(cudaMemset(dkeys, 0xff, sizeof(int) * (size_t)slots));
    const unsigned grid = (unsigned)((count + kSpgemmThreads - 1) / kSpgemmThreads);
    spgemm_hash_kernel<<<grid, kSpgemmThreads>>>(j0, count, dAp, dAj, dAx, dBp, dBj, dBx, doff, dcap, dkeys, dvals, du);
    CUDA_OK(cudaGetLastError());
    ucount.resize((size_t)count);
    CUDA_OK(cudaMemcpy(ucount.data(), du, sizeof(int) * (size_t)count, cudaMemcpyDeviceToHost));
    lap(t_hash);
    cptr.resize((size_t)count);
    long long bn = 0;
    for (int64_t t = 0; t < count; ++t) {
      cptr[t] = bn;
      bn += ucount[t];
      R->colcount[j0 + t] = ucount[t];
    }
    long long* dcptr = D.upload(cptr.data(), count);
    spgemm_emit_kernel<<<grid, kSpgemmThreads>>>(count, doff, dcap, dkeys, dvals, dcptr, dCj, dCx);
    CUDA_OK(cudaGetLastError());
    const size_t old = R->rows.size();
    R->rows.resize(old + (size_t)bn);
    R->vals.resize(old + (size_t)bn);
    if (bn) {
      CUDA_OK(cudaMemcpy(R->rows.data() + old, dCj, sizeof(int) * (size_t)bn, cudaMemcpyDeviceToHost));
      CUDA_OK(cudaMemcpy(R->vals.data() + old, dCx, sizeof(double) * (size_t)bn, cudaMemcpyDeviceToHost));
    }
    total += bn;
    lap(t_emit);
    REQUIRE(total < INT32_MAX, B200AMG_ERR_UNSUPPORTED, "the product has more than 2^31 entries");
    // the per-batch descriptors are small; release them now (the big buffers are reused)
    for (void* q : {(void*)doff, (void*)dcap, (void*)du, (void*)dcptr}) D.release(q);
    j0 = j1;
  }
  CUDA_OK(cudaDeviceSynchronize());
  if (t_all.on)
    fprintf(stderr, "[b200amg] spgemm %lld x %lld x %lld nnz(C)=%lld: upload+plan %.3f s, hash %.3f s, emit+download %.3f s\n", (long long)m,
            (long long)k, (long long)n, (long long)total, t_up, t_hash, t_emit);
  *nnz_out = total;
  g_spgemm_pending = R.release();
  API_END
}

int32_t b200amg_spgemm_release(void) {
  API_BEGIN
  delete g_spgemm_pending;
  g_spgemm_pending = nullptr;
  g_spgemm_scratch.release();
  API_END
}

int32_t b200amg_spgemm_fetch(int32_t* Cp, int32_t* Cj, double* Cx) {
  API_BEGIN
  REQUIRE(g_spgemm_pending, B200AMG_ERR_STATE, "no pending product: call b200amg_spgemm_begin first");
  REQUIRE(Cp, B200AMG_ERR_BAD_ARG, "null colptr");
  std::unique_ptr<SpgemmPending> R(g_spgemm_pending);
  g_spgemm_pending = nullptr;
  Cp[0] = 0;
  for (int64_t j = 0; j < R->n; ++j) Cp[j + 1] = Cp[j] + R->colcount[j];
  if (!R->rows.empty()) {
    REQUIRE(Cj && Cx, B200AMG_ERR_BAD_ARG, "null output arrays");
    std::memcpy(Cj, R->rows.data(), sizeof(int) * R->rows.size());
    std::memcpy(Cx, R->vals.data(), sizeof(double) * R->vals.size());
  }
  API_END
}

int32_t b200amg_get_stream(b200amg_handle_t h, void** stream) {
  API_BEGIN
  REQUIRE(h && stream, B200AMG_ERR_BAD_ARG, "null argument");
  *stream = (void*)h->stream;
  API_END
}

int32_t b200amg_device_vectors(b200amg_handle_t h, double** x, double** b) {
  API_BEGIN
  check_ready(h);
  check_not_partitioned(h, "device_vectors");
  if (x) *x = h->x0;
  if (b) *b = h->b0;
  API_END
}

}  // extern "C"
