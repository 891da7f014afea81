// Host-side staging of the matrices that cross the C ABI (Julia's CSC, 1-based Int64 or 0-based int32) into the engine's
// 0-based int32 compressed rows, with validation.  Host-only: shared by the device engine and the host-only plan entry points.
#pragma once
#include "engine_base.h"
#include "host_csr.h"

namespace b200amg {}
using namespace b200amg;

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


