// Device Galerkin products of the setup phase (b200amg_spgemm_begin / _fetch / _release; kernels in spgemm.cuh): its own
// translation unit, independent of the hierarchy handle.
#include "engine_base.h"
#include "spgemm.cuh"

using namespace b200amg;

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


}  // extern "C"
