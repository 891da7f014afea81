// Shared by every translation unit of libb200amg.so: error reporting across the C ABI (thread-local message, status
// codes, exceptions never leave an entry point), CUDA / argument checks, environment knobs, upload stage timers and the
// small device-memory helpers.
#pragma once
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
#include "b200amg.h"

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
inline thread_local std::string g_err;   // one per thread, shared by all translation units (b200amg_last_error)
inline int32_t fail(int32_t code, const char* fmt, ...) {
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

