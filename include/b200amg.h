/*
 * b200amg.h — C-ABI of the B200-native AMG solve-phase engine (libb200amg.so).
 *
 * This is the drop-in boundary for the ONE hot path of AlgebraicMultigrid.jl that moves to
 * the device: the multigrid cycle loop (src/multilevel.jl:152-239) and the relaxations
 * (src/smoother.jl:1-582), plus the stdlib SpMV / restriction / prolongation / norm calls they
 * make (call sites src/multilevel.jl:170,188-190,219-220,223,226,233-234), the coarse-solver
 * APPLY (src/coarse_solver.jl:16,75-81) and the preconditioner entry (src/preconditioner.jl:12-24).
 * Hierarchy SETUP stays on the host; its result is handed over level by level.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  Every function returns an int32 status
 *     (0 = ok, negative = error); b200amg_last_error() gives a thread-local message.  Nothing
 *     throws across the boundary.  The Julia shim turns a non-zero status into error(...).
 *   - matrices arrive exactly as Julia holds them: CSC arrays (colptr, rowval, nzval), fp64
 *     values, Int64 1-based indices (index_bits = 64, index_base = 1).  0-based and/or int32
 *     indices are accepted too (what this repo's own host side keeps).  Row indices must be
 *     sorted inside each column (SparseMatrixCSC invariant).  The arrays are BORROWED for the
 *     duration of the call only; the library converts to its own device layouts.
 *   - `adjoint != 0` means the operator is the lazy `Adjoint` of the stored CSC matrix — how the
 *     reference stores P for Ruge-Stuben (src/classical.jl:64-65) and R for smoothed
 *     aggregation (src/aggregation.jl:158-159).  m, n are ALWAYS the stored parent's size.
 *   - vectors: contiguous fp64, length n of the level.  memkind says whether x/b/y pointers are
 *     host memory (the library does H2D/D2H around the call) or device memory on the handle's GPU.
 *   - one in-flight call per handle (the reference's workspace is shared mutable state too:
 *     src/multilevel.jl:176,218-225); distinct handles are independent.  Calls return after the
 *     result is complete (stream-synchronised).
 */
#ifndef B200AMG_H
#define B200AMG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200AMG_VERSION 100

typedef struct b200amg_hierarchy* b200amg_handle_t;
typedef struct b200amg_smoother_obj* b200amg_smoother_handle_t;

/* status codes */
enum {
  B200AMG_OK = 0,
  B200AMG_ERR_BAD_ARG = -1,
  B200AMG_ERR_DIM_MISMATCH = -2,      /* Julia: DimensionMismatch */
  B200AMG_ERR_SINGULAR = -3,          /* Julia: SingularException (NoSymmetry GS/SOR setup, smoother.jl:239-241) */
  B200AMG_ERR_CUDA = -4,
  B200AMG_ERR_NCCL = -5,
  B200AMG_ERR_OOM = -6,
  B200AMG_ERR_STATE = -7,             /* call order violated (e.g. solve before finalize) */
  B200AMG_ERR_UNSUPPORTED = -8,
  B200AMG_ERR_NO_DEVICE = -9,         /* no CUDA device: there is NO CPU fallback */
  B200AMG_ERR_CALLBACK = -10          /* the coarse-solver callback reported a failure */
};

/* smoother kinds — src/smoother.jl:18-23 (GaussSeidel), :92-99 (Jacobi), :173-180 (SOR) */
enum { B200AMG_SMOOTHER_NONE = 0, B200AMG_SMOOTHER_GS = 1, B200AMG_SMOOTHER_JACOBI = 2, B200AMG_SMOOTHER_SOR = 3 };
/* sweeps — src/smoother.jl:11-17 */
enum { B200AMG_SWEEP_FORWARD = 1, B200AMG_SWEEP_BACKWARD = 2, B200AMG_SWEEP_SYMMETRIC = 3 };
/* symmetry tags — src/utils.jl:1-5.  HERMITIAN selects the "fast" column-as-row smoothers,
 * NONE the true-A variants (smoother.jl:144-171, :226-582). */
enum { B200AMG_SYMMETRY_HERMITIAN = 0, B200AMG_SYMMETRY_NONE = 1 };
/* cycles — src/multilevel.jl:116-124, :200-212 */
enum { B200AMG_CYCLE_V = 0, B200AMG_CYCLE_W = 1, B200AMG_CYCLE_F = 2 };
enum { B200AMG_MEM_HOST = 0, B200AMG_MEM_DEVICE = 1 };
enum { B200AMG_OP_A = 0, B200AMG_OP_P = 1, B200AMG_OP_R = 2 };
enum { B200AMG_PRE = 0, B200AMG_POST = 1 };

/* A SparseMatrixCSC (optionally wrapped in a lazy Adjoint). */
typedef struct {
  int64_t m, n;            /* size of the STORED matrix */
  const void* colptr;      /* n+1 entries */
  const void* rowval;      /* nnz entries, sorted per column */
  const double* nzval;     /* nnz entries */
  int32_t index_bits;      /* 32 or 64 */
  int32_t index_base;      /* 0 or 1 */
  int32_t adjoint;         /* 1: operator = stored' */
  int32_t reserved;
} b200amg_csc_t;

/* A smoother configuration: GaussSeidel(sweep; iter) / Jacobi(ω; iter) / SOR(ω, sweep; iter). */
typedef struct {
  int32_t kind;
  int32_t sweep;
  int32_t iter;
  int32_t reserved;
  double omega;
} b200amg_smoother_t;

const char* b200amg_last_error(void);
int32_t b200amg_version(void);
/* number of CUDA devices visible (0 => every compute entry point returns B200AMG_ERR_NO_DEVICE) */
int32_t b200amg_device_count(void);

/* ---- hierarchy lifecycle : replaces the MultiLevel / Level / MultiLevelWorkspace objects
 *      (src/multilevel.jl:1-59) as the thing the cycle runs on ------------------------------- */
int32_t b200amg_create(b200amg_handle_t* out, int32_t device);
/* push!(levels, Level(A, P, R, pre, post))  — src/classical.jl:48-52, src/aggregation.jl:147-151.
 * A: n x n.  P: n x nc operator, R: nc x n operator (either may be given as adjoint of the other's
 * storage, or both as plain CSC as test/gmg.jl:40-46 does). */
int32_t b200amg_add_level(b200amg_handle_t h, const b200amg_csc_t* A, const b200amg_csc_t* P,
                          const b200amg_csc_t* R, const b200amg_smoother_t* pre,
                          const b200amg_smoother_t* post, int32_t symmetry);
/* final_A + coarse solver (src/multilevel.jl:16-17).  The factorisation is setup work and
 * stays on the host: the caller passes the dense n x n matrix M (column-major) whose product
 * M*b is the coarse solve — pinv(A) for Pinv (src/coarse_solver.jl:9-16), A^-1 for the
 * QR/LU solvers (:66-81).  final_A is used when the hierarchy has no levels (multilevel.jl:167). */
int32_t b200amg_set_coarse(b200amg_handle_t h, const b200amg_csc_t* final_A, int64_t n,
                           const double* coarse_inverse_colmajor);
/* The same with the coarse solver as a HOST CALLABLE — the reference's `coarse_solver(x, b)` is any callable
 * (src/multilevel.jl:180,228; src/coarse_solver.jl:24-58 wraps LinearSolve factorisations, :66-81 a sparse QR), which is
 * what a coarsest level too large for a dense operator needs (coarsening stopped at max_levels).  Per coarse solve the
 * library copies coarse_b into pinned host memory, calls fn(user, n, 1, x_host, b_host) from a CUDA runtime thread in
 * stream order (a host node when the cycle is a captured graph) and copies x back; fn must not call CUDA or this library
 * on the same handle, returns 0 on success; a non-zero status surfaces as B200AMG_ERR_CALLBACK from the entry point that
 * ran the cycle.  Use instead of b200amg_set_coarse. */
typedef int32_t (*b200amg_coarse_fn)(void* user, int64_t n, int64_t ncols, double* x_host, const double* b_host);
int32_t b200amg_set_coarse_callback(b200amg_handle_t h, const b200amg_csc_t* final_A, int64_t n,
                                    b200amg_coarse_fn fn, void* user);
/* Optional, BEFORE the first add_level: make this handle one rank (one process, one GPU) of a
 * hierarchy whose FINE level is split by contiguous row blocks over world_size ranks; coarser
 * levels live on rank 0.  Every rank then passes the SAME full hierarchy to add_level / set_coarse
 * (each keeps only its part) and the same full-length x, b to solve / cycle / precond; x comes
 * back assembled on every rank.  Fine-level smoothers must be Jacobi (Gauss-Seidel does not shard).
 * nccl_unique_id: the 128-byte ncclUniqueId created on rank 0 (b200amg_nccl_unique_id) and
 * broadcast by the host (torch.distributed / MPI / a file). */
int32_t b200amg_set_partition(b200amg_handle_t h, int32_t rank, int32_t world_size,
                              const void* nccl_unique_id, int64_t id_bytes);
/* rank 0 creates the id (ncclGetUniqueId) that every rank passes to b200amg_set_partition; cap >= 128. */
int32_t b200amg_nccl_unique_id(void* out, int64_t cap);
/* what this rank holds of the partitioned fine level: owned rows [row_lo,row_hi), halo entries received /
 * entries sent per exchange, restricted coarse rows [coarse_lo,coarse_hi), coarse_x window [cx_lo,cx_hi). */
int32_t b200amg_partition_info(b200amg_handle_t h, int64_t* row_lo, int64_t* row_hi, int64_t* nhalo, int64_t* nsend,
                               int64_t* coarse_lo, int64_t* coarse_hi, int64_t* cx_lo, int64_t* cx_hi);
/* Host-only (no device needed): the partition plan rank `rank` of `world` derives for a fine level —
 * row blocks, restricted coarse rows, the sorted halo column list with its per-owner segments
 * (recv_off), the owned entries to send per destination (send_idx / send_off) and every rank's
 * coarse_x window.  Capacities: *_split, recv_off, send_off: world+1; cx_lo, cx_hi: world;
 * halo_cols, send_idx: cap. */
int32_t b200amg_partition_plan(const b200amg_csc_t* A, const b200amg_csc_t* P, const b200amg_csc_t* R, int32_t rank, int32_t world,
                               int64_t* row_split, int64_t* coarse_split, int32_t* halo_cols, int64_t* nhalo, int32_t* recv_off,
                               int32_t* send_idx, int64_t* nsend, int32_t* send_off, int64_t* cx_lo, int64_t* cx_hi, int64_t cap);
/* The same for a level BELOW a partitioned one (B200AMG_OPT_PART_LEVELS > 1): its row blocks are the parent's
 * coarse_split (so the parent's restriction writes straight into this level's owned b), and the entries of this
 * level's x that the parent's owned rows of P reference are part of the halo (the prolongation reads them after one
 * halo exchange).  parent_P: the prolongation of the level above; parent_row_split / parent_coarse_split: world + 1
 * entries from the parent's plan. */
int32_t b200amg_partition_plan_child(const b200amg_csc_t* A, const b200amg_csc_t* P, const b200amg_csc_t* R, const b200amg_csc_t* parent_P,
                                     const int64_t* parent_row_split, const int64_t* parent_coarse_split, int32_t rank, int32_t world,
                                     int64_t* row_split, int64_t* coarse_split, int32_t* halo_cols, int64_t* nhalo, int32_t* recv_off,
                                     int32_t* send_idx, int64_t* nsend, int32_t* send_off, int64_t* cx_lo, int64_t* cx_hi, int64_t cap);
/* Host-only (no device needed): the plan of the blocked exact-order Gauss-Seidel / SOR sweep for matrix A (structurally
 * symmetric) — tiles, stages, local steps, cross-tile requirements (csrc/device/block_plan.h) — built, validated against
 * every invariant the kernel relies on, and optionally EXECUTED by the host emulation of the kernel: x_out = the sweep(s)
 * `sweep` (1 forward, 2 backward, 3 symmetric) applied to x with right-hand side b (all in the caller's numbering).
 * params (may be NULL; 0 = default): [0] contiguous tile rows, [1] block a, [2] block b, [3] stage entries, [4] stage rows,
 * [5] window, [6] depth, [7] verbose, [8] 1: emulate the PASS sweep (csrc/device/pass_plan.h, pass_gs.cuh: slabs, near / far
 * codes, window) on the same plan instead (stats[0] -2: no pass layout, -3: its addressing rules violated; [2] / [3] / [4] then
 * report chunks / passes / lanes of the pass layout).  params holds 9 entries.  stats[16]: [0] 1 ok / 0 not applicable / -1 invariant violated (msg says which),
 * [1] tiles, [2] stages, [3] steps, [4] lanes per row, [5] wavefronts of the whole level, [6] 1000 x mean rows per step,
 * [7] theta, [8] a, [9] b, [10] rows of the largest tile, [11] steps of the longest tile, [12] / [13] forward / backward
 * requirements, [14] / [15] extents of the K / J coordinates.  new_of_old (may be NULL): n entries. */
int32_t b200amg_block_plan_check(const b200amg_csc_t* A, const int64_t* params, int64_t* stats, int32_t* new_of_old, const double* x,
                                 const double* b, double* x_out, double omega, int32_t sor, int32_t sweep, char* msg, int64_t msg_cap);
/* builds device layouts (row-major operators, transposes, wavefront schedules), workspaces and
 * the captured cycle graphs. */
int32_t b200amg_finalize(b200amg_handle_t h);
int32_t b200amg_destroy(b200amg_handle_t h);

/* ---- the solve phase ----------------------------------------------------------------------- */
/* _solve!(x, ml, b, cycle; maxiter, abstol, reltol, log, calculate_residual)
 * src/multilevel.jl:158-198.  x: in = initial guess, out = solution.  residuals (may be NULL):
 * caller-allocated, cap entries; receives ||b|| then one ||b - A x|| per iteration (the `log`
 * history); *nres = entries written.  *iters = cycles executed. */
int32_t b200amg_solve(b200amg_handle_t h, double* x, const double* b, int32_t cycle,
                      int32_t maxiter, double abstol, double reltol, int32_t calculate_residual,
                      double* residuals, int32_t cap, int32_t* nres, int32_t* iters,
                      int32_t memkind);
/* The same for MATRIX right-hand sides (block workspaces, src/multilevel.jl:28-59; LinearSolve precs with bs > 1,
 * src/precs.jl:12-17): x, b are n x ncols, column-major with leading dimension ld (a Julia Matrix).  Columns are relaxed /
 * restricted / prolonged one by one (src/smoother.jl:77,118,195) and ONE Frobenius norm over all columns is tested
 * (multilevel.jl:170,190).  Every column stays on the device for the whole call. */
int32_t b200amg_solve_block(b200amg_handle_t h, double* x, const double* b, int64_t ncols, int64_t ld, int32_t cycle,
                            int32_t maxiter, double abstol, double reltol, int32_t calculate_residual, double* residuals,
                            int32_t cap, int32_t* nres, int32_t* iters, int32_t memkind);
/* __solve!(x, ml, cycle, b, 1): exactly one cycle on the caller's x (src/multilevel.jl:214-239). */
int32_t b200amg_cycle(b200amg_handle_t h, double* x, const double* b, int32_t cycle, int32_t memkind);
/* ldiv!(x, p::Preconditioner, b): x .= 0 (init_zero) or x .= b, then one cycle, no residual
 * (src/preconditioner.jl:12-19). */
int32_t b200amg_precond(b200amg_handle_t h, double* x, const double* b, int32_t cycle,
                        int32_t init_zero, int32_t memkind);
/* smooth!(x, levels[level].presmoother|postsmoother, b)  — src/multilevel.jl:216,236. level is 0-based. */
int32_t b200amg_smooth(b200amg_handle_t h, int32_t level, int32_t which, double* x, const double* b,
                       int32_t memkind);
/* mul!(y, levels[level].{A|P|R}, x)  — src/multilevel.jl:219,223,233; src/preconditioner.jl:20.
 * level == number of levels addresses final_A (op A only). */
int32_t b200amg_apply(b200amg_handle_t h, int32_t level, int32_t op, double* y, const double* x,
                      int32_t memkind);
/* res = b - A_level x  — src/multilevel.jl:188-189, :219-220 fused */
int32_t b200amg_residual(b200amg_handle_t h, int32_t level, double* r, const double* b,
                         const double* x, int32_t memkind);
/* coarse_solver(x, b)  — src/multilevel.jl:180,228 */
int32_t b200amg_coarse_solve(b200amg_handle_t h, double* x, const double* b, int32_t memkind);
/* norm(v) over a level-sized vector — src/multilevel.jl:170,190 */
int32_t b200amg_norm(b200amg_handle_t h, int64_t n, const double* v, double* out, int32_t memkind);

/* Preconditioned CG that stays on the device (the caller of ldiv! in the reference's tests is
 * IterativeSolvers.cg: test/cycle_tests.jl:25, test/runtests.jl:186,204).  Left-preconditioned
 * CG on A_1 with one `cycle` per iteration as the preconditioner; stops when
 * ||r|| <= max(reltol*||r0||, abstol) or after maxiter iterations.  residuals as in solve. */
int32_t b200amg_pcg(b200amg_handle_t h, double* x, const double* b, int32_t cycle, int32_t maxiter,
                    double abstol, double reltol, double* residuals, int32_t cap, int32_t* nres,
                    int32_t* iters, int32_t memkind);

/* ---- standalone smoother objects: setup_smoother(config, A, symmetry) / smooth!(x, s, b)
 *      (src/smoother.jl:1-10,25-49) ---------------------------------------------------------- */
int32_t b200amg_smoother_create(b200amg_smoother_handle_t* out, int32_t device, const b200amg_csc_t* A,
                                const b200amg_smoother_t* config, int32_t symmetry);
int32_t b200amg_smoother_apply(b200amg_smoother_handle_t s, double* x, const double* b, int32_t memkind);
int32_t b200amg_smoother_destroy(b200amg_smoother_handle_t s);

/* ---- introspection / measurement ----------------------------------------------------------- */
int32_t b200amg_num_levels(b200amg_handle_t h);  /* length(ml) = levels + 1 */
/* level 0..length-1: rows, nnz of A_level; nnz of P (0 for the last); number of forward
 * wavefronts of its Gauss-Seidel schedule (0 if not built). */
int32_t b200amg_level_info(b200amg_handle_t h, int32_t level, int64_t* n, int64_t* nnz_a,
                           int64_t* nnz_p, int64_t* wavefronts);
/* kernels launched by this handle since creation (graph replays count their kernel nodes). */
/* Bytes per stored matrix VALUE the bandwidth kernels read on a level (cap >= 3): out[0] A, out[1] P, out[2] R — 4 when the
 * lossless binary32 copy is in use (B200AMG_OPT_FP32_STORAGE), 8 for fp64, 0 when the operator does not exist on this rank. */
int32_t b200amg_storage_info(b200amg_handle_t h, int32_t level, int32_t* out, int32_t cap);
int64_t b200amg_launch_count(b200amg_handle_t h);
/* Communication counters of a row-partitioned handle (cap >= 4): [0] NCCL groups / collectives enqueued so far, [1] halo
 * exchanges done over peer memory so far, [2] 1 if halo exchanges use peer memory (CUDA IPC mappings of the neighbours'
 * vectors, direct NVLink stores: csrc/device/peer_halo.cuh), 0 if NCCL send/recv, [3] partitioned levels. */
int32_t b200amg_comm_stats(b200amg_handle_t h, int64_t* out, int32_t cap);
/* Time `reps` back-to-back launches of one hot-path kernel on the handle's stream with CUDA
 * events; returns the average milliseconds per launch in *ms.  what: 0 spmv y=A x, 1 residual,
 * 2 pre-smoother, 3 restriction, 4 prolongation+correction, 5 one full cycle, 6 norm,
 * 7 one halo exchange of the fine-level x (partitioned handles only). */
int32_t b200amg_time_kernel(b200amg_handle_t h, int32_t level, int32_t what, int32_t cycle, int32_t reps,
                            int32_t flush_l2, double* ms);
/* The six phases the reference times with @timeit_debug (src/multilevel.jl:216-236):
 * 0 Presmoother, 1 Residual eval, 2 Restriction, 3 Coarse solve, 4 Prolongation, 5 Postsmoother.
 * Runs one un-captured cycle with an event pair around every phase; ms[level*6 + phase]
 * accumulates (cap >= 6*length). */
int32_t b200amg_profile_cycle(b200amg_handle_t h, int32_t cycle, double* ms, int32_t cap);
/* Engine options (before or after finalize).  USE_GRAPHS (default 1): replay each cycle type as a
 * captured CUDA graph.  TIME_RESIDUAL (default 0): b200amg_solve brackets the fine-level
 * convergence-residual kernel (multilevel.jl:188-189) of every iteration with a CUDA event pair on
 * the launching stream; read the per-iteration milliseconds back with b200amg_residual_timings.
 * STREAM_CHUNK (default 1): consecutive tiles each CTA of the TMA stream kernels takes per run
 * (0 = one contiguous range per CTA).  GS_MODE (default 2) selects the Gauss-Seidel / SOR sweep:
 * 0 one launch per wavefront, 1 wavefront-counter dataflow kernel, 2 TMA-fed per-row mailbox kernel
 * on wide levels (mean wavefront >= GS_MAIL_MIN_WIDTH rows), one CTA on levels of <= GS_CTA_ROWS rows,
 * counter kernel in between, 3 like 2 with the ticket mailbox kernel.  GS_ACQUIRE / GS_POLL_SLEEP /
 * GS_GATE_SLEEP tune the hand-off.  PART_LEVELS (default 1; before the first add_level of a partitioned handle):
 * how many of the finest levels are split by rows over the ranks (a level below a partitioned one inherits its parent's
 * coarse-row ownership, so restriction needs no communication and prolongation one more halo exchange).
 * GS_DSM (default 1): levels whose wavefronts are narrower than GS_MAIL_MIN_WIDTH and whose x fits the shared memory of a
 * thread-block cluster of <= 2^GS_DSM_MAX_CTAS_LOG2 CTAs (default 2: 4 CTAs, ~68 000 rows; up to 4: 16 CTAs, ~260 000 rows)
 * are swept by ONE cluster with x in distributed shared memory (csrc/device/dsm_gs.cuh); 0 = the GS_MODE kernels on every
 * level.  GS_DSM_FENCE bit 0 / bit 1 add a cluster-scope fence on the producer / consumer side of its hand-off.
 * FP32_STORAGE (default: the environment variable B200AMG_FP32_STORAGE at upload time, 0 if unset): an operator all of whose
 * values are exactly representable in binary32 (stencils, aggregation prolongators, any hierarchy built from Float32 input —
 * test/runtests.jl:244-259) is ALSO stored with 4-byte values when B200AMG_FP32_STORAGE=1 is set while the hierarchy is
 * uploaded, and the bandwidth kernels (SpMV, residual, restriction / prolongation, Jacobi) read that copy: 8 instead of 12
 * bytes per entry, the same fp64 products and sums, bit-identical results; 0 = read the 8-byte values.  Off by default
 * because it measured slower on B200.
 * Cycle graphs already captured keep the values they were captured with. */
enum { B200AMG_OPT_USE_GRAPHS = 0, B200AMG_OPT_TIME_RESIDUAL = 1, B200AMG_OPT_STREAM_CHUNK = 2, B200AMG_OPT_GS_MODE = 3, B200AMG_OPT_GS_ACQUIRE = 4, B200AMG_OPT_GS_POLL_SLEEP = 5,
       B200AMG_OPT_GS_GATE_SLEEP = 6, B200AMG_OPT_GS_CTA_ROWS = 7,
       B200AMG_OPT_GS_MAIL_MIN_WIDTH = 8, B200AMG_OPT_GS_CLUSTER = 9,
       B200AMG_OPT_PART_LEVELS = 12, B200AMG_OPT_GS_DSM = 13, B200AMG_OPT_GS_DSM_FENCE = 14, B200AMG_OPT_GS_DSM_MAX_CTAS_LOG2 = 15,
       B200AMG_OPT_FP32_STORAGE = 17,
       B200AMG_OPT_GS_DSM2 = 18 /* gs_dsm2_kernel: two consumer groups alternate the tiles of the one-cluster sweep */ };
int32_t b200amg_set_option(b200amg_handle_t h, int32_t option, double value);
int32_t b200amg_residual_timings(b200amg_handle_t h, double* ms, int32_t cap, int32_t* n);
/* Diagnostics: run one dataflow Gauss-Seidel sweep (forward / backward) of `level` on the level's
 * current vectors and return 8 %globaltimer stamps (ns) per task: claimed, prefetch issued, wait
 * done, gathers may start, row written, CTA barrier passed, fence done, count published. */
int32_t b200amg_debug_gs_timeline(b200amg_handle_t h, int32_t level, int32_t backward, uint64_t* out, int64_t cap,
                                  int64_t* ntasks);
/* the CUDA stream (cudaStream_t) every kernel of this handle is launched on */
/* Setup phase, Galerkin products on the device: C = A * B for A (m x k) and B (k x n) given as
 * compressed-sparse-column arrays with int32 0-based indices and sorted rows per column — the products `R*A` and
 * `(R*A)*P` of `extend_hierarchy!` (src/classical.jl:46, src/aggregation.jl:145).  Semantics of the stdlib product the
 * reference calls: sorted rows per column, structural zeros kept; the accumulation order per entry is the sequential
 * column-by-column one, so the result is bit-identical to the host product.  _begin computes and returns nnz(C);
 * _fetch copies colptr (n + 1), rowval and nzval (nnz) out and drops the pending result.  Not re-entrant.
 * B200AMG_SPGEMM_SLOTS_M (environment, default 192): hash-table slots per batch of columns, in units of 2^20. */
int32_t b200amg_spgemm_begin(int32_t device, int64_t m, int64_t k, int64_t n, const int32_t* Ap, const int32_t* Aj, const double* Ax,
                             const int32_t* Bp, const int32_t* Bj, const double* Bx, int64_t* nnz_out);
int32_t b200amg_spgemm_fetch(int32_t* Cp, int32_t* Cj, double* Cx);
/* frees the hash-table scratch the products keep between calls (and any result not fetched) */
int32_t b200amg_spgemm_release(void);
int32_t b200amg_get_stream(b200amg_handle_t h, void** stream);
/* raw device pointers of the level-0 work vectors (x, b) for zero-copy callers (torch / CUDA.jl) */
int32_t b200amg_device_vectors(b200amg_handle_t h, double** x, double** b);

#ifdef __cplusplus
}
#endif
#endif /* B200AMG_H */
