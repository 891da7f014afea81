// Helpers of the entry points: device selection, fault checks, host <-> device vector staging, level construction.
// Part of engine.cu (one translation unit).
#pragma once

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

