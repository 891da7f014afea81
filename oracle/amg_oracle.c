/*
 * amg_oracle.c — CPU restatement of AlgebraicMultigrid.jl's SOLVE PHASE.  TEST INFRASTRUCTURE.
 *
 * This file is the parity oracle for the CUDA engine behind include/b200amg.h.  It is NOT
 * part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it.  Nothing under algebraicmultigrid.jl_b200/ imports it.
 *
 * Why a restatement: the reference is pure Julia and there is no Julia runtime in this image
 * (nor on the GPU box), so the reference itself cannot be executed.  Each function below
 * follows the cited reference lines (paths relative to /root/reference) statement by
 * statement — same loop order, same accumulation order, separate multiply and add (build with
 * -ffp-contract=off: the reference requests no fused multiply-add anywhere).  The sparse
 * matrix-vector products live in Julia's stdlib SparseArrays, which is not vendored in
 * /root/reference; they are restated from the stdlib's published algorithm (column scatter for
 * mul!(y, ::SparseMatrixCSC, x), per-column dot for the Adjoint method) and anchored on the
 * reference's own call sites and golden vectors.
 *
 * Pinning (tests/test_oracle_goldens.py): every known-answer the reference's tests hold for
 * this path — exact-rational Gauss-Seidel answers (test/sa_tests.jl:316-379), the SGSx4
 * 10-vector (test/test_regression.jl:14-23), the 46-value single-cycle vectors and the CG
 * goldens on test/thing.jl (test/runtests.jl:143-224), convergence on poisson(1000)/randlap
 * (:112-141), V/W/F on poisson((50,50)) (test/cycle_tests.jl), issue #56 (test_regression.jl:59-69),
 * fast == general smoothers (test/test_smoothers.jl:29-45).
 *
 * Data convention: CSC, 0-based int32 indices, fp64 (the reference's Int64 1-based arrays minus
 * the offset).  "column i treated as row i" is kept exactly as the reference does it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t idx_t;

typedef struct {
  int64_t m, n;
  const idx_t* colptr;
  const idx_t* rowval;
  const double* nzval;
  int adjoint; /* operator = stored' */
} ocsc_t;

typedef struct {
  int kind;  /* 1 GS, 2 Jacobi, 3 SOR */
  int sweep; /* 1 fwd, 2 bwd, 3 sym */
  int iter;
  double omega;
} osmoother_t;

typedef struct {
  ocsc_t A, P, R;
  osmoother_t pre, post;
  int symmetry; /* 0 Hermitian (fast smoothers), 1 NoSymmetry */
  idx_t* diagidx; /* NoSymmetry GS/SOR: DiagonalIndices */
  double* diagvals; /* NoSymmetry Jacobi */
  double *res, *coarse_x, *coarse_b, *temp;
} olevel_t;

typedef struct {
  int nlevels;
  int cap;
  olevel_t* levels;
  ocsc_t final_A;
  int64_t nc;
  const double* coarse_inv; /* nc x nc, column-major */
  double* res_final;
} ohier_t;

/* ------------------------------------------------------------------------------------------
 * Julia stdlib SparseArrays (not vendored): mul!(y, A::SparseMatrixCSC, x), 5-arg form with
 * alpha = 1, beta = 0: fill!(y, 0); for col: axj = x[col]; for k in nzrange: y[rowval[k]] += nzval[k]*axj
 * Call sites: src/multilevel.jl:188,219,223(RS),233(SA).
 * ------------------------------------------------------------------------------------------ */
static void csc_mul(const ocsc_t* A, const double* x, double* y) {
  for (int64_t i = 0; i < A->m; ++i) y[i] = 0.0;
  for (int64_t col = 0; col < A->n; ++col) {
    const double axj = x[col];
    for (idx_t k = A->colptr[col]; k < A->colptr[col + 1]; ++k) y[A->rowval[k]] += A->nzval[k] * axj;
  }
}
/* mul!(y, A'::Adjoint{SparseMatrixCSC}, x): for col: tmp = 0; for k: tmp += nzval[k]'*x[rowval[k]]; y[col] = tmp
 * Call sites: src/multilevel.jl:223 (SA restriction), :233 (RS prolongation). */
static void csc_adjoint_mul(const ocsc_t* A, const double* x, double* y) {
  for (int64_t col = 0; col < A->n; ++col) {
    double tmp = 0.0;
    for (idx_t k = A->colptr[col]; k < A->colptr[col + 1]; ++k) tmp += A->nzval[k] * x[A->rowval[k]];
    y[col] = tmp;
  }
}
static void op_mul(const ocsc_t* A, const double* x, double* y) {
  if (A->adjoint) csc_adjoint_mul(A, x, y); else csc_mul(A, x, y);
}
static int64_t op_rows(const ocsc_t* A) { return A->adjoint ? A->n : A->m; }

void oracle_mul(int64_t m, int64_t n, const idx_t* colptr, const idx_t* rowval, const double* nzval,
                int adjoint, const double* x, double* y) {
  ocsc_t A = {m, n, colptr, rowval, nzval, adjoint};
  op_mul(&A, x, y);
}

/* norm(v): LinearAlgebra.norm -> BLAS dnrm2 for Vector{Float64} (src/multilevel.jl:170,190).
 * Restated as the scaled-free textbook form; order of accumulation is unspecified in BLAS,
 * which is why residual histories are compared with a relative tolerance. */
double oracle_norm(int64_t n, const double* v) {
  long double s = 0.0L;
  for (int64_t i = 0; i < n; ++i) s += (long double)v[i] * (long double)v[i];
  return (double)sqrtl(s);
}

/* ------------------------------------------------------------------------------------------
 * gs!  — src/smoother.jl:73-90
 * ------------------------------------------------------------------------------------------ */
static void gs(const ocsc_t* A, const double* b, double* x, int64_t start, int64_t step, int64_t stop) {
  for (int64_t i = start; step > 0 ? i <= stop : i >= stop; i += step) {
    double rsum = 0.0, d = 0.0;
    for (idx_t j = A->colptr[i]; j < A->colptr[i + 1]; ++j) {
      const idx_t row = A->rowval[j];
      const double val = A->nzval[j];
      if (i == row) d = val; else rsum += val * x[row];
    }
    x[i] = (d == 0) ? x[i] : (b[i] - rsum) / d;
  }
}
/* smooth!(x, ::FastGSSmoother, b)  — src/smoother.jl:61-71 */
static void smooth_fast_gs(const ocsc_t* A, const osmoother_t* s, double* x, const double* b) {
  const int64_t n = A->m;
  for (int it = 0; it < s->iter; ++it) {
    if (s->sweep == 1 || s->sweep == 3) gs(A, b, x, 0, 1, n - 1);
    if (s->sweep == 2 || s->sweep == 3) gs(A, b, x, n - 1, -1, 0);
  }
}
/* smooth!(x, ::FastJacobiSmoother, b)  — src/smoother.jl:113-141 */
static void smooth_fast_jacobi(const ocsc_t* A, const osmoother_t* s, double* x, const double* b, double* temp) {
  const int64_t n = A->m;
  const double one = 1.0, w = s->omega;
  for (int it = 0; it < s->iter; ++it) {
    for (int64_t i = 0; i < n; ++i) temp[i] = x[i];
    for (int64_t i = 0; i < n; ++i) {
      double rsum = 0.0, diag = 0.0;
      for (idx_t j = A->colptr[i]; j < A->colptr[i + 1]; ++j) {
        const idx_t row = A->rowval[j];
        const double val = A->nzval[j];
        if (row == i) diag = val; else rsum += val * temp[row];
      }
      const double xcand = (one - w) * temp[i] + w * ((b[i] - rsum) / diag);
      x[i] = (diag == 0.0) ? x[i] : xcand;
    }
  }
}
/* smooth!(x, ::JacobiSmoother, b) (NoSymmetry)  — src/smoother.jl:157-171; diagvals = diag(A) :152-155 */
static void smooth_general_jacobi(const ocsc_t* A, const osmoother_t* s, double* x, const double* b,
                                  double* temp, const double* diagvals) {
  const int64_t n = A->m;
  for (int it = 0; it < s->iter; ++it) {
    csc_mul(A, x, temp);
    for (int64_t i = 0; i < n; ++i) temp[i] -= b[i];
    for (int64_t i = 0; i < n; ++i) {
      const double d = diagvals[i];
      if (d != 0.0) x[i] -= s->omega * temp[i] / d;
    }
  }
}
/* sor_step!  — src/smoother.jl:205-221 ; smooth!(x, ::FastSORSmoother, b) :193-203 */
static void sor_step(const ocsc_t* A, const double* b, double* x, double w, int64_t start, int64_t step, int64_t stop) {
  for (int64_t i = start; step > 0 ? i <= stop : i >= stop; i += step) {
    double rsum = 0.0, d = 0.0;
    for (idx_t j = A->colptr[i]; j < A->colptr[i + 1]; ++j) {
      const idx_t row = A->rowval[j];
      const double val = A->nzval[j];
      if (i == row) d = val; else rsum += val * x[row];
    }
    x[i] = (d == 0) ? x[i] : (1 - w) * x[i] + (w / d) * (b[i] - rsum);
  }
}
static void smooth_fast_sor(const ocsc_t* A, const osmoother_t* s, double* x, const double* b) {
  const int64_t n = A->m;
  for (int it = 0; it < s->iter; ++it) {
    if (s->sweep == 1 || s->sweep == 3) sor_step(A, b, x, s->omega, 0, 1, n - 1);
    if (s->sweep == 2 || s->sweep == 3) sor_step(A, b, x, s->omega, n - 1, -1, 0);
  }
}

/* ------------------------------------------------------------------------------------------
 * NoSymmetry Gauss-Seidel / SOR family  — src/smoother.jl:226-582
 * ------------------------------------------------------------------------------------------ */
/* DiagonalIndices(A)  :233-248 ; returns 0 or -(col+1) for SingularException(col) */
static int diagonal_indices(const ocsc_t* A, idx_t* diag) {
  for (int64_t col = 0; col < A->n; ++col) {
    idx_t r1 = A->colptr[col], r2 = A->colptr[col + 1] - 1;
    /* searchsortedfirst(rowval, col, r1, r2) */
    idx_t lo = r1, hi = r2 + 1;
    while (lo < hi) { idx_t mid = lo + (hi - lo) / 2; if (A->rowval[mid] < col) lo = mid + 1; else hi = mid; }
    r1 = lo;
    if (r1 > r2 || A->rowval[r1] != col || A->nzval[r1] == 0.0) return -(int)(col + 1);
    diag[col] = r1;
  }
  return 0;
}
/* forward_sub!(F::FastLowerTriangular, x)  :282-300 */
static void forward_sub(const ocsc_t* A, const idx_t* diag, double* x) {
  for (int64_t col = 0; col < A->n; ++col) {
    const idx_t idx = diag[col];
    x[col] /= A->nzval[idx];
    for (idx_t i = idx + 1; i < A->colptr[col + 1]; ++i) x[A->rowval[i]] -= A->nzval[i] * x[col];
  }
}
/* forward_sub!(alpha, F, x, beta, y)  :305-323 */
static void forward_sub5(double alpha, const ocsc_t* A, const idx_t* diag, double* x, double beta, const double* y) {
  for (int64_t col = 0; col < A->n; ++col) {
    const idx_t idx = diag[col];
    x[col] = alpha * x[col] / A->nzval[idx] + beta * y[col];
    for (idx_t i = idx + 1; i < A->colptr[col + 1]; ++i) x[A->rowval[i]] -= A->nzval[i] * x[col];
  }
}
/* backward_sub!(F::FastUpperTriangular, x)  :329-347 */
static void backward_sub(const ocsc_t* A, const idx_t* diag, double* x) {
  for (int64_t col = A->n - 1; col >= 0; --col) {
    const idx_t idx = diag[col];
    x[col] /= A->nzval[idx];
    for (idx_t i = A->colptr[col]; i < idx; ++i) x[A->rowval[i]] -= A->nzval[i] * x[col];
  }
}
/* backward_sub!(alpha, F, x, beta, y)  :349-367 */
static void backward_sub5(double alpha, const ocsc_t* A, const idx_t* diag, double* x, double beta, const double* y) {
  for (int64_t col = A->n - 1; col >= 0; --col) {
    const idx_t idx = diag[col];
    x[col] = alpha * x[col] / A->nzval[idx] + beta * y[col];
    for (idx_t i = A->colptr[col]; i < idx; ++i) x[A->rowval[i]] -= A->nzval[i] * x[col];
  }
}
/* gauss_seidel_multiply!(alpha, U::StrictlyUpperTriangular, x, beta, y, z)  :373-388 */
static void gsmul_upper(double alpha, const ocsc_t* A, const idx_t* diag, const double* x, double beta,
                        const double* y, double* z) {
  for (int64_t col = 0; col < A->n; ++col) {
    const double ax = alpha * x[col];
    for (idx_t j = A->colptr[col]; j < diag[col]; ++j) z[A->rowval[j]] += A->nzval[j] * ax;
    z[col] = beta * y[col];
  }
}
/* gauss_seidel_multiply!(alpha, L::StrictlyLowerTriangular, x, beta, y, z)  :394-408 */
static void gsmul_lower(double alpha, const ocsc_t* A, const idx_t* diag, const double* x, double beta,
                        const double* y, double* z) {
  for (int64_t col = A->n - 1; col >= 0; --col) {
    const double ax = alpha * x[col];
    z[col] = beta * y[col];
    for (idx_t j = diag[col] + 1; j < A->colptr[col + 1]; ++j) z[A->rowval[j]] += A->nzval[j] * ax;
  }
}
/* smooth! for Forward/Backward/SymmetricGaussSeidelSmoother  :421-483 */
static void smooth_general_gs(const ocsc_t* A, const idx_t* diag, const osmoother_t* s, double* x, const double* b) {
  for (int it = 0; it < s->iter; ++it) {
    if (s->sweep == 1 || s->sweep == 3) { gsmul_upper(-1.0, A, diag, x, 1.0, b, x); forward_sub(A, diag, x); }
    if (s->sweep == 2 || s->sweep == 3) { gsmul_lower(-1.0, A, diag, x, 1.0, b, x); backward_sub(A, diag, x); }
  }
}
/* smooth! for Forward/Backward/SymmetricSORSmoother  :499-582 */
static void smooth_general_sor(const ocsc_t* A, const idx_t* diag, const osmoother_t* s, double* x,
                               const double* b, double* tmp) {
  const int64_t n = A->n;
  const double w = s->omega;
  for (int it = 0; it < s->iter; ++it) {
    if (s->sweep == 1 || s->sweep == 3) {
      gsmul_upper(-1.0, A, diag, x, 1.0, b, tmp);
      forward_sub5(w, A, diag, tmp, 1.0 - w, x);
      memcpy(x, tmp, sizeof(double) * n);
    }
    if (s->sweep == 2 || s->sweep == 3) {
      gsmul_lower(-1.0, A, diag, x, 1.0, b, tmp);
      backward_sub5(w, A, diag, tmp, 1.0 - w, x);
      memcpy(x, tmp, sizeof(double) * n);
    }
  }
}

/* setup_smoother + smooth! dispatch  — src/smoother.jl:56-59,108-111,152-155,188-191,416-419,... */
static int smooth_dispatch(const ocsc_t* A, const osmoother_t* s, int symmetry, double* x, const double* b,
                           double* temp, idx_t* diagidx, double* diagvals) {
  if (s->kind == 0) return 0;
  if (symmetry == 0) {
    if (s->kind == 1) smooth_fast_gs(A, s, x, b);
    else if (s->kind == 2) smooth_fast_jacobi(A, s, x, b, temp);
    else if (s->kind == 3) smooth_fast_sor(A, s, x, b);
    else return -1;
  } else {
    if (s->kind == 2) smooth_general_jacobi(A, s, x, b, temp, diagvals);
    else if (s->kind == 1) smooth_general_gs(A, diagidx, s, x, b);
    else if (s->kind == 3) smooth_general_sor(A, diagidx, s, x, b, temp);
    else return -1;
  }
  return 0;
}

/* standalone: (config::Smoother)(A, x, b, symmetry)  — src/smoother.jl:34-38
 * returns 0, or -(col+1) for SingularException(col). */
int oracle_smooth(int64_t n, const idx_t* colptr, const idx_t* rowval, const double* nzval, int kind,
                  int sweep, int iter, double omega, int symmetry, double* x, const double* b) {
  ocsc_t A = {n, n, colptr, rowval, nzval, 0};
  osmoother_t s = {kind, sweep, iter, omega};
  double* temp = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  idx_t* di = NULL;
  double* dv = NULL;
  int rc = 0;
  if (symmetry == 1) {
    if (kind == 2) {
      dv = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
      for (int64_t c = 0; c < n; ++c)
        for (idx_t k = colptr[c]; k < colptr[c + 1]; ++k)
          if (rowval[k] == c) dv[c] = nzval[k];
    } else {
      di = (idx_t*)malloc(sizeof(idx_t) * (size_t)(n > 0 ? n : 1));
      rc = diagonal_indices(&A, di);
    }
  }
  if (rc == 0) rc = smooth_dispatch(&A, &s, symmetry, x, b, temp, di, dv);
  free(temp); free(di); free(dv);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * hierarchy container (MultiLevel / Level / MultiLevelWorkspace — src/multilevel.jl:1-59)
 * ------------------------------------------------------------------------------------------ */
ohier_t* oracle_create(void) { return (ohier_t*)calloc(1, sizeof(ohier_t)); }

static ocsc_t mk(int64_t m, int64_t n, const idx_t* cp, const idx_t* rv, const double* nz, int adj) {
  ocsc_t a = {m, n, cp, rv, nz, adj};
  return a;
}

int oracle_add_level(ohier_t* h, int64_t n, const idx_t* Acp, const idx_t* Arv, const double* Anz,
                     int64_t Pm, int64_t Pn, const idx_t* Pcp, const idx_t* Prv, const double* Pnz, int Padj,
                     int64_t Rm, int64_t Rn, const idx_t* Rcp, const idx_t* Rrv, const double* Rnz, int Radj,
                     int pre_kind, int pre_sweep, int pre_iter, double pre_omega,
                     int post_kind, int post_sweep, int post_iter, double post_omega, int symmetry) {
  if (h->nlevels == h->cap) {
    h->cap = h->cap ? 2 * h->cap : 8;
    h->levels = (olevel_t*)realloc(h->levels, sizeof(olevel_t) * (size_t)h->cap);
  }
  olevel_t* L = &h->levels[h->nlevels];
  memset(L, 0, sizeof(*L));
  L->A = mk(n, n, Acp, Arv, Anz, 0);
  L->P = mk(Pm, Pn, Pcp, Prv, Pnz, Padj);
  L->R = mk(Rm, Rn, Rcp, Rrv, Rnz, Radj);
  L->pre.kind = pre_kind; L->pre.sweep = pre_sweep; L->pre.iter = pre_iter; L->pre.omega = pre_omega;
  L->post.kind = post_kind; L->post.sweep = post_sweep; L->post.iter = post_iter; L->post.omega = post_omega;
  L->symmetry = symmetry;
  const int64_t nc = op_rows(&L->R);
  L->res = (double*)calloc((size_t)n + 1, sizeof(double));
  L->temp = (double*)calloc((size_t)n + 1, sizeof(double));
  L->coarse_x = (double*)calloc((size_t)nc + 1, sizeof(double));
  L->coarse_b = (double*)calloc((size_t)nc + 1, sizeof(double));
  if (symmetry == 1) {
    int need_di = (pre_kind == 1 || pre_kind == 3 || post_kind == 1 || post_kind == 3);
    int need_dv = (pre_kind == 2 || post_kind == 2);
    if (need_di) {
      L->diagidx = (idx_t*)malloc(sizeof(idx_t) * ((size_t)n + 1));
      int rc = diagonal_indices(&L->A, L->diagidx);
      if (rc) return rc;
    }
    if (need_dv) {
      L->diagvals = (double*)calloc((size_t)n + 1, sizeof(double));
      for (int64_t c = 0; c < n; ++c)
        for (idx_t k = Acp[c]; k < Acp[c + 1]; ++k)
          if (Arv[k] == c) L->diagvals[c] = Anz[k];
    }
  }
  h->nlevels++;
  return 0;
}

int oracle_set_coarse(ohier_t* h, int64_t n, const idx_t* cp, const idx_t* rv, const double* nz,
                      const double* coarse_inv_colmajor) {
  h->final_A = mk(n, n, cp, rv, nz, 0);
  h->nc = n;
  h->coarse_inv = coarse_inv_colmajor;
  free(h->res_final);
  h->res_final = (double*)calloc((size_t)n + 1, sizeof(double));
  return 0;
}

void oracle_destroy(ohier_t* h) {
  if (!h) return;
  for (int i = 0; i < h->nlevels; ++i) {
    olevel_t* L = &h->levels[i];
    free(L->res); free(L->temp); free(L->coarse_x); free(L->coarse_b); free(L->diagidx); free(L->diagvals);
  }
  free(h->levels); free(h->res_final); free(h);
}

/* coarse solver apply: (p::Pinv)(x, b) = mul!(x, p.pinvA, b)  — src/coarse_solver.jl:16.
 * The default QRSolver (:66-81) computes F\b; for the nonsingular coarse matrices that is the
 * same vector as inv(A)*b, which the host hands over as a dense matrix in both cases. */
static void coarse_solve(const ohier_t* h, double* x, const double* b) {
  const int64_t n = h->nc;
  for (int64_t i = 0; i < n; ++i) x[i] = 0.0;
  for (int64_t j = 0; j < n; ++j) {
    const double bj = b[j];
    const double* col = h->coarse_inv + j * n;
    for (int64_t i = 0; i < n; ++i) x[i] += col[i] * bj;
  }
}
void oracle_coarse_solve(const ohier_t* h, double* x, const double* b) { coarse_solve(h, x, b); }

/* __solve!(x, ml, cycle, b, lvl)  — src/multilevel.jl:214-239 ; cycle: 0 V, 1 W, 2 F */
static void solve_level(ohier_t* h, double* x, int cycle, const double* b, int lvl) {
  olevel_t* L = &h->levels[lvl];
  const int64_t n = L->A.m;
  smooth_dispatch(&L->A, &L->pre, L->symmetry, x, b, L->temp, L->diagidx, L->diagvals);   /* :216 */
  double* res = L->res;
  csc_mul(&L->A, x, res);                                                                  /* :219 */
  for (int64_t i = 0; i < n; ++i) res[i] = b[i] - res[i];                                   /* :220 */
  double* coarse_b = L->coarse_b;
  op_mul(&L->R, res, coarse_b);                                                             /* :223 */
  double* coarse_x = L->coarse_x;
  const int64_t nc = op_rows(&L->R);
  for (int64_t i = 0; i < nc; ++i) coarse_x[i] = 0.0;                                       /* :226 */
  if (lvl == h->nlevels - 1) {
    coarse_solve(h, coarse_x, coarse_b);                                                    /* :228 */
  } else {                                                                                  /* :200-212 */
    if (cycle == 0) solve_level(h, coarse_x, 0, coarse_b, lvl + 1);
    else if (cycle == 1) { solve_level(h, coarse_x, 1, coarse_b, lvl + 1); solve_level(h, coarse_x, 1, coarse_b, lvl + 1); }
    else { solve_level(h, coarse_x, 2, coarse_b, lvl + 1); solve_level(h, coarse_x, 0, coarse_b, lvl + 1); }
  }
  op_mul(&L->P, coarse_x, res);                                                             /* :233 */
  for (int64_t i = 0; i < n; ++i) x[i] += res[i];                                           /* :234 */
  smooth_dispatch(&L->A, &L->post, L->symmetry, x, b, L->temp, L->diagidx, L->diagvals);    /* :236 */
}

int oracle_cycle(ohier_t* h, double* x, const double* b, int cycle) {
  if (h->nlevels == 0) coarse_solve(h, x, b); else solve_level(h, x, cycle, b, 0);
  return 0;
}

/* _solve!(x, ml, b, cycle; maxiter, abstol, reltol, log, calculate_residual)  — src/multilevel.jl:158-198 */
int oracle_solve(ohier_t* h, double* x, const double* b, int cycle, int maxiter, double abstol, double reltol,
                 int calculate_residual, double* residuals, int cap, int* nres, int* iters) {
  const ocsc_t* A = h->nlevels == 0 ? &h->final_A : &h->levels[0].A;
  const int64_t n = A->m;
  int nr = 0;
  double normres, normb;
  normres = normb = oracle_norm(n, b);                                                      /* :170 */
  if (normb != 0) abstol = fmax(reltol * normb, abstol);                                    /* :171-173 */
  if (residuals && nr < cap) residuals[nr++] = normb;                                       /* :174 */
  double* res = h->nlevels == 0 ? h->res_final : h->levels[0].res;                          /* :176 */
  int itr = 1;
  while (itr <= maxiter && (!calculate_residual || normres > abstol)) {                     /* :178 */
    if (h->nlevels == 0) coarse_solve(h, x, b); else solve_level(h, x, cycle, b, 0);        /* :179-183 */
    if (calculate_residual) {
      csc_mul(A, x, res);                                                                   /* :188 */
      for (int64_t i = 0; i < n; ++i) res[i] = b[i] - res[i];                               /* :189 */
      normres = oracle_norm(n, res);                                                        /* :190 */
      if (residuals && nr < cap) residuals[nr++] = normres;                                 /* :191 */
    }
    itr += 1;
  }
  if (nres) *nres = nr;
  if (iters) *iters = itr - 1;
  return 0;
}

/* ldiv!(x, p::Preconditioner, b)  — src/preconditioner.jl:12-19 */
int oracle_precond(ohier_t* h, double* x, const double* b, int cycle, int init_zero) {
  const int64_t n = h->nlevels == 0 ? h->final_A.m : h->levels[0].A.m;
  if (init_zero) for (int64_t i = 0; i < n; ++i) x[i] = 0.0;
  else for (int64_t i = 0; i < n; ++i) x[i] = b[i];
  return oracle_solve(h, x, b, cycle, 1, 0.0, 0.0, 0, NULL, 0, NULL, NULL);
}

/* IterativeSolvers.cg(A, b; Pl = p, abstol, reltol, maxiter) as the reference's tests call it
 * (test/runtests.jl:186,204; test/cycle_tests.jl:25).  IterativeSolvers.jl is a test-only,
 * un-vendored dependency (Project.toml:27, version unpinned); restated from its published
 * PCGIterable: c = Pl\r; rho = c.r; u = c + (rho/rho_prev) u; c = A u; alpha = rho/(u.c);
 * x += alpha u; r -= alpha c; stop when ||r|| <= max(reltol*||r0||, abstol) or iteration == maxiter.
 * x starts at zero.  use_precond = 0 gives plain CG. */
int oracle_pcg(ohier_t* h, double* x, const double* b, int cycle, int use_precond, int maxiter, double abstol,
               double reltol, double* residuals, int cap, int* nres, int* iters) {
  const ocsc_t* A = h->nlevels == 0 ? &h->final_A : &h->levels[0].A;
  const int64_t n = A->m;
  double* r = (double*)malloc(sizeof(double) * (size_t)(n + 1));
  double* c = (double*)calloc((size_t)(n + 1), sizeof(double));
  double* u = (double*)calloc((size_t)(n + 1), sizeof(double));
  for (int64_t i = 0; i < n; ++i) { x[i] = 0.0; r[i] = b[i]; }
  double residual = oracle_norm(n, r);
  const double tol = fmax(reltol * residual, abstol);
  double rho = 1.0;
  int nr = 0, it = 0;
  if (residuals && nr < cap) residuals[nr++] = residual;
  while (!(it >= maxiter || residual <= tol)) {
    if (use_precond) oracle_precond(h, c, r, cycle, 1); else memcpy(c, r, sizeof(double) * (size_t)n);
    const double rho_prev = rho;
    rho = 0.0;
    for (int64_t i = 0; i < n; ++i) rho += c[i] * r[i];
    const double beta = rho / rho_prev;
    for (int64_t i = 0; i < n; ++i) u[i] = c[i] + beta * u[i];
    csc_mul(A, u, c);
    double uc = 0.0;
    for (int64_t i = 0; i < n; ++i) uc += u[i] * c[i];
    const double alpha = rho / uc;
    for (int64_t i = 0; i < n; ++i) x[i] += alpha * u[i];
    for (int64_t i = 0; i < n; ++i) r[i] -= alpha * c[i];
    residual = oracle_norm(n, r);
    if (residuals && nr < cap) residuals[nr++] = residual;
    ++it;
  }
  if (nres) *nres = nr;
  if (iters) *iters = it;
  free(r); free(c); free(u);
  return 0;
}
