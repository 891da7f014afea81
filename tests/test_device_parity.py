"""GPU parity: the CUDA engine (through the C-ABI, via the host mirror of the reference API) against
the CPU oracle on the same seeded inputs, against the reference's golden vectors, and through
size-independent properties at larger sizes.

Stated fp64 tolerances (relative, infinity norm unless noted):
  single kernel (SpMV / residual / restriction / prolongation / Jacobi)   1e-13
  Gauss-Seidel / SOR sweeps (wavefront order == lexicographic order)        1e-12
  one V/W/F cycle                                                           1e-11
  _solve: identical iteration count, residual history 1e-6 per entry, final x 1e-9 (2-norm)
The only source of difference is the summation order inside a row (lane-strided partial sums +
shuffle tree, FMA contraction) — the update order of Gauss-Seidel is the reference's.
"""
import numpy as np
import pytest

import goldens
import oracle

pytestmark = pytest.mark.gpu

TOL_KERNEL, TOL_SWEEP, TOL_CYCLE, TOL_SOLVE_X, TOL_HIST = 1e-13, 1e-12, 1e-11, 1e-9, 1e-6


def relinf(a, b):
    d = np.abs(np.asarray(a) - np.asarray(b)).max() if len(a) else 0.0
    s = np.abs(b).max() if len(b) else 0.0
    return d / s if s > 0 else d


def _rng(seed=0):
    return np.random.default_rng(seed)


def _thing_b():
    b = np.zeros(46)
    b[0], b[1] = 1, -1
    return b


@pytest.fixture(scope="module")
def jac(amg):
    return amg.Jacobi(2.0 / 3.0)


# ---- single kernels --------------------------------------------------------------------------
@pytest.mark.parametrize("dims", [(1000,), (50, 50), (17, 9, 13), (64, 64, 64)])
def test_spmv_residual_restrict_prolong(amg, dims):
    A = amg.poisson(dims if len(dims) > 1 else dims[0])
    for ml in (amg.ruge_stuben(A), amg.smoothed_aggregation(A)):
        dev = ml.device()
        r = _rng(1)
        for lv, level in enumerate(ml.levels):
            n, nc = level.A.n, level.R.shape[0]
            x, b, xc = r.standard_normal(n), r.standard_normal(n), r.standard_normal(nc)
            y = dev.apply(lv, 0, np.empty(n), x)
            assert relinf(y, oracle.mul(level.A, x)) <= TOL_KERNEL
            res = dev.residual(lv, np.empty(n), b, x)
            assert relinf(res, b - oracle.mul(level.A, x)) <= TOL_KERNEL
            Rst, Radj = oracle.oracle._storage(level.R)
            Pst, Padj = oracle.oracle._storage(level.P)
            assert relinf(dev.apply(lv, 2, np.empty(nc), x), oracle.mul(Rst, x, adjoint=bool(Radj))) <= TOL_KERNEL
            assert relinf(dev.apply(lv, 1, np.empty(n), xc), oracle.mul(Pst, xc, adjoint=bool(Padj))) <= TOL_KERNEL
        nf = ml.final_A.n
        xf = r.standard_normal(nf)
        assert relinf(dev.apply(len(ml.levels), 0, np.empty(nf), xf), oracle.mul(ml.final_A, xf)) <= TOL_KERNEL
        assert relinf(dev.coarse_solve(np.empty(nf), xf), ml.coarse_solver.dense_operator() @ xf) <= 1e-12
        assert abs(dev.norm(x) - oracle.norm(x)) <= 1e-14 * oracle.norm(x)
        ml.release()


def test_spmv_nonsymmetric_uses_true_A(amg, fx):
    A = fx.sprand_plus_diag(300, 0.03, 5.0, seed=7)
    ml = amg.ruge_stuben(A, symmetry=amg.NoSymmetry())
    x = _rng(2).standard_normal(300)
    y = ml.device().apply(0, 0, np.empty(300), x)
    assert relinf(y, A.to_scipy() @ x) <= TOL_KERNEL
    ml.release()


# ---- smoothers ---------------------------------------------------------------------------------
def test_gauss_seidel_known_answers_on_device(amg):
    # test/sa_tests.jl:316-379 — exact rationals, bit-for-bit
    fwd, bwd = amg.GaussSeidel(amg.ForwardSweep()), amg.GaussSeidel(amg.BackwardSweep())
    A1, A3 = amg.poisson(1), amg.poisson(3)

    def run(s, A, x, b):
        x = np.array(x, dtype=float)
        s(A, x, np.array(b, dtype=float))
        return x

    assert np.array_equal(run(fwd, A1, [0.0], [0.0]), [0.0])
    assert np.array_equal(run(fwd, A3, [0.0, 1, 2], [0.0, 0, 0]), [1 / 2, 5 / 4, 5 / 8])
    assert np.array_equal(run(bwd, A3, [0.0, 1, 2], [0.0, 0, 0]), [1 / 8, 1 / 4, 1 / 2])
    assert np.array_equal(run(fwd, A1, [0.0], [10.0]), [5.0])
    assert np.array_equal(run(fwd, A3, [0.0, 1, 2], [10.0, 20, 30]), [11 / 2, 55 / 4, 175 / 8])
    A100 = amg.poisson(100)
    x1 = run(amg.GaussSeidel(amg.ForwardSweep(), 200), A100, np.ones(100), np.zeros(100))
    x2 = run(amg.GaussSeidel(amg.BackwardSweep(), 200), A100, np.ones(100), np.zeros(100))
    r1, r2 = np.linalg.norm(A100.matvec(x1)), np.linalg.norm(A100.matvec(x2))
    assert r1 < 0.01 and r2 < 0.01 and np.isclose(r1, r2)


def test_regression_26_on_device(amg):
    x = np.ones(10)
    amg.GaussSeidel(amg.SymmetricSweep(), 4)(amg.poisson(10), x, np.zeros(10))
    assert np.sum((x - goldens.SGS4_POISSON10) ** 2) < 1e-6


def _smoother_zoo(amg):
    F_, B_, S_ = amg.ForwardSweep(), amg.BackwardSweep(), amg.SymmetricSweep()
    return [amg.Jacobi(4 / 5, iter=2), amg.Jacobi(2 / 3), amg.GaussSeidel(F_), amg.GaussSeidel(B_), amg.GaussSeidel(S_, 2),
            amg.SOR(0.5, F_), amg.SOR(0.5, B_), amg.SOR(1.3, S_, 2)]


@pytest.mark.parametrize("case", ["poisson1d", "poisson2d", "poisson3d", "rs_coarse", "thing", "nonsym"])
@pytest.mark.parametrize("symmetry", ["hermitian", "none"])
def test_smoothers_vs_oracle(amg, fx, case, symmetry):
    if case == "poisson1d":
        A = amg.poisson(777)
    elif case == "poisson2d":
        A = amg.poisson((40, 33))
    elif case == "poisson3d":
        A = amg.poisson((20, 17, 12))
    elif case == "rs_coarse":
        A = amg.ruge_stuben(amg.poisson((24, 24, 24))).levels[1].A     # irregular, ~18 nnz/row, symmetric to rounding only
    elif case == "thing":
        A = fx.matrix("thing")
    else:
        A = fx.sprand_plus_diag(400, 0.02, 5.0, seed=11)
    sym = amg.HermitianSymmetry() if symmetry == "hermitian" else amg.NoSymmetry()
    r = _rng(5)
    x0, b = r.standard_normal(A.n), r.standard_normal(A.n)
    for sm in _smoother_zoo(amg):
        x = x0.copy()
        sm(A, x, b, sym)
        ref = oracle.smooth(A, sm, x0.copy(), b, symmetry=symmetry)
        tol = TOL_KERNEL if sm.kind == "jacobi" else TOL_SWEEP
        assert relinf(x, ref) <= tol, (case, symmetry, sm)


def test_smoother_block_right_hand_sides_column_by_column(amg):
    """smooth!(x, s, b) with n x m blocks relaxes every column (src/smoother.jl:77,118,195 `for col in 1:size(x, 2)`):
    the standalone device smoother must do the same, column against column of the oracle."""
    A = amg.poisson((12, 11, 10))
    r = _rng(9)
    X0, B = r.standard_normal((A.n, 2)), r.standard_normal((A.n, 2))
    for sm in [amg.GaussSeidel(), amg.Jacobi(2 / 3, iter=2), amg.SOR(1.1, amg.ForwardSweep())]:
        X = X0.copy()
        sm(A, X, B, amg.HermitianSymmetry())
        for j in range(2):
            ref = oracle.smooth(A, sm, X0[:, j].copy(), B[:, j].copy())
            assert relinf(X[:, j], ref) <= TOL_SWEEP, (sm, j)
    with pytest.raises((AssertionError, ValueError)):
        amg.GaussSeidel()(A, X0.copy(), B[:, 0].copy(), amg.HermitianSymmetry())


def test_fast_equals_general_on_symmetric_device(amg):
    # test/test_smoothers.jl:29-45
    A = amg.poisson(50)
    x0, b = _rng(3).random(50), np.ones(50)
    for sm in [amg.Jacobi(4 / 5, iter=2), amg.GaussSeidel(amg.SymmetricSweep(), iter=2), amg.SOR(0.5, iter=2)]:
        xf, xg = x0.copy(), x0.copy()
        sm(A, xf, b, amg.HermitianSymmetry())
        sm(A, xg, b, amg.NoSymmetry())
        assert np.allclose(xf, xg), sm


def test_nosymmetry_smoothers_converge_device(amg, fx):
    # test/test_smoothers.jl:15-27
    N = 50
    A = fx.sprand_plus_diag(N, 0.05, 5.0, seed=1)
    x0, b = _rng(2).random(N), np.ones(N)
    for sm in [amg.Jacobi(1 / 6, iter=500), amg.GaussSeidel(amg.ForwardSweep(), 100), amg.GaussSeidel(amg.BackwardSweep(), 100),
               amg.GaussSeidel(amg.SymmetricSweep(), 100), amg.SOR(0.5, amg.ForwardSweep(), 100),
               amg.SOR(0.5, amg.BackwardSweep(), 100), amg.SOR(0.5, amg.SymmetricSweep(), 100)]:
        x = x0.copy()
        sm(A, x, b, amg.NoSymmetry())
        assert np.allclose(A.matvec(x), b), sm


def test_singular_exception_device(amg):
    import scipy.sparse as sp

    A = amg.SparseMatrixCSC.from_scipy(sp.csc_matrix(np.array([[1.0, 2.0], [3.0, 0.0]])))
    with pytest.raises(amg.SingularException):
        amg.GaussSeidel(amg.ForwardSweep())(A, np.ones(2), np.ones(2), amg.NoSymmetry())
    # the fast variants silently skip zero diagonals (smoother.jl:87)
    x = np.ones(2)
    amg.GaussSeidel(amg.ForwardSweep())(A, x, np.ones(2))
    assert np.array_equal(x, oracle.smooth(A, amg.GaussSeidel(amg.ForwardSweep()), np.ones(2), np.ones(2)))


def test_level_smooth_entry(amg):
    ml = amg.ruge_stuben(amg.poisson((30, 30)))
    r = _rng(8)
    for lv, level in enumerate(ml.levels):
        x0, b = r.standard_normal(level.A.n), r.standard_normal(level.A.n)
        x = ml.device().smooth(lv, 0, x0.copy(), b)
        assert relinf(x, oracle.smooth(level.A, level.presmoother.config, x0.copy(), b)) <= TOL_SWEEP
    ml.release()


# ---- cycles ------------------------------------------------------------------------------------
def _hierarchies(amg, A, jac):
    F_ = amg.GaussSeidel(amg.ForwardSweep())
    yield "rs-sgs", amg.ruge_stuben(A)
    yield "rs-fgs-pinv", amg.ruge_stuben(A, presmoother=F_, postsmoother=F_, coarse_solver=amg.Pinv)
    yield "rs-jacobi", amg.ruge_stuben(A, presmoother=jac, postsmoother=jac)
    yield "sa-sgs", amg.smoothed_aggregation(A)
    yield "sa-jacobi", amg.smoothed_aggregation(A, presmoother=jac, postsmoother=jac)
    yield "sa-sor", amg.smoothed_aggregation(A, presmoother=amg.SOR(1.2), postsmoother=amg.SOR(0.9, amg.BackwardSweep(), 2))


@pytest.mark.parametrize("dims", [(1000,), (50, 50), (20, 20, 20)])
def test_one_cycle_vs_oracle(amg, jac, dims):
    A = amg.poisson(dims if len(dims) > 1 else dims[0])
    r = _rng(4)
    b, x0 = r.random(A.n), r.standard_normal(A.n)
    for name, ml in _hierarchies(amg, A, jac):
        H = oracle.OracleHierarchy(ml)
        for cyc, cname in ((amg.V(), "V"), (amg.W(), "W"), (amg.F(), "F")):
            x = amg._solve_(x0.copy(), ml, b, cyc, maxiter=1, calculate_residual=False)
            ref = H.solve(b, cycle=cname, x0=x0, maxiter=1, calculate_residual=False)
            assert relinf(x, ref) <= TOL_CYCLE, (name, cname)
            # ldiv!: zero + one cycle
            p = amg.aspreconditioner(ml, cyc)
            assert relinf(amg.backslash(p, b), H.precond(b, cycle=cname)) <= TOL_CYCLE, (name, cname)
        ml.release()


def test_solve_vs_oracle(amg, jac):
    A = amg.poisson((50, 50))
    b = A.matvec(np.ones(A.n))
    for name, ml in _hierarchies(amg, A, jac):
        H = oracle.OracleHierarchy(ml)
        for cyc, cname in ((amg.V(), "V"), (amg.W(), "W"), (amg.F(), "F")):
            x, hist = amg._solve(ml, b, cyc, reltol=1e-8, log=True)
            xr, histr = H.solve(b, cycle=cname, reltol=1e-8, log=True)
            assert len(hist) == len(histr), (name, cname, len(hist), len(histr))
            assert np.allclose(hist, histr, rtol=TOL_HIST, atol=0), (name, cname)
            assert np.linalg.norm(x - xr) <= TOL_SOLVE_X * np.linalg.norm(xr), (name, cname)
            if len(hist) <= 100:                       # converged within maxiter (test/cycle_tests.jl:16)
                assert np.linalg.norm(b - A.matvec(x)) < 1e-8 * np.linalg.norm(b)
        ml.release()


@pytest.mark.parametrize("method", ["rs", "sa"])
def test_julia_index_convention_crosses_the_abi(amg, jac, method):
    """The arrays a Julia caller holds are Int64 and 1-based, with the lazy Adjoint flag on P (Ruge-Stuben,
    src/classical.jl:64-65) or on R (smoothed aggregation, src/aggregation.jl:158-159).  Uploading the same hierarchy in that
    form must give bit-for-bit the results of this package's own int32 / 0-based upload."""
    from algebraicmultigrid_jl_b200 import _devlib

    A = amg.poisson((14, 13, 12))
    ml = amg.ruge_stuben(A) if method == "rs" else amg.smoothed_aggregation(A, presmoother=jac, postsmoother=jac)
    adj = [hasattr(lv.P, "parent") for lv in ml.levels], [hasattr(lv.R, "parent") for lv in ml.levels]
    assert all(adj[0]) if method == "rs" else all(adj[1])          # the lazy-Adjoint descriptors are really exercised
    b = _rng(4).standard_normal(A.n)
    d0 = _devlib.DeviceHierarchy(ml)
    d1 = _devlib.DeviceHierarchy(ml, julia_indices=True)
    try:
        x0, x1 = np.zeros(A.n), np.zeros(A.n)
        h0, it0 = d0.solve(x0, b, 0, 30, 0.0, 1e-10, True)
        h1, it1 = d1.solve(x1, b, 0, 30, 0.0, 1e-10, True)
        assert it0 == it1 and np.array_equal(h0, h1) and np.array_equal(x0, x1)
        for cyc in (1, 2):
            y0, y1 = np.zeros(A.n), np.zeros(A.n)
            d0.cycle(y0, b, cyc)
            d1.cycle(y1, b, cyc)
            assert np.array_equal(y0, y1)
        H = oracle.OracleHierarchy(ml)
        xr, hr = H.solve(b, maxiter=30, reltol=1e-10, log=True)
        assert len(hr) == len(h1) and relinf(x1, xr) <= TOL_SOLVE_X
    finally:
        d0.close()
        d1.close()


def test_thing_goldens_on_device(amg, fx):
    # test/runtests.jl:143-224 — the reference's own golden vectors, Σdiff² < 1e-8
    A = fx.matrix("thing")
    sm = amg.GaussSeidel(amg.ForwardSweep())
    ml = amg.ruge_stuben(A, presmoother=sm, postsmoother=sm, coarse_solver=amg.Pinv)
    b = _thing_b()
    x = amg._solve(ml, A.matvec(np.ones(46)), maxiter=1, abstol=1e-12)
    assert np.sum((x - goldens.THING_ZERO_GOLDEN) ** 2) < 1e-8
    x = amg.solve(A, b, amg.RugeStubenAMG(), presmoother=sm, postsmoother=sm, maxiter=1, abstol=1e-12, coarse_solver=amg.Pinv)
    assert np.sum((x - goldens.THING_FWDGS_ONE_CYCLE) ** 2) < 1e-8
    x = amg.cg(A, b, Pl=amg.aspreconditioner(ml))
    assert np.sum((x - goldens.THING_CG_FWDGS) ** 2) < 1e-8
    ml.release()
    ml = amg.ruge_stuben(A, coarse_solver=amg.Pinv)
    x = amg.cg(A, b, Pl=amg.aspreconditioner(ml), maxiter=100_000, reltol=1e-6)
    assert np.sum((x - goldens.THING_CG_SGS) ** 2) < 1e-8
    x = amg._solve(ml, b, maxiter=1, reltol=1e-12)
    assert np.sum((x - goldens.THING_SGS_ONE_CYCLE) ** 2) < 1e-8
    ml.release()


def test_solver_poisson1000_device(amg, fx):
    # test/runtests.jl:112-141 (config C1)
    A = amg.poisson(1000)
    b = A.matvec(np.ones(1000))
    assert np.sum((amg._solve(amg.ruge_stuben(A), b) - 1) ** 2) < 1e-8
    fs = amg.GaussSeidel(amg.ForwardSweep())
    assert np.sum((amg._solve(amg.ruge_stuben(A, presmoother=fs, postsmoother=fs), b) - 1) ** 2) < 1e-8
    A = fx.matrix("randlap")
    b = A.matvec(np.ones(100))
    assert np.sum(amg._solve(amg.ruge_stuben(A, presmoother=fs, postsmoother=fs), b) ** 2) < 1e-8
    assert np.sum(amg._solve(amg.ruge_stuben(A), b) ** 2) < 1e-6


def test_cycles_device(amg):
    # test/cycle_tests.jl:6-30
    A = amg.poisson((50, 50))
    b = A.matvec(np.ones(A.n))
    for method in (amg.ruge_stuben, amg.smoothed_aggregation):
        ml = method(A)
        for cyc in (amg.V(), amg.W(), amg.F()):
            x = amg._solve(ml, b, cyc, reltol=1e-8)
            assert np.linalg.norm(b - A.matvec(x)) < 1e-8 * np.linalg.norm(b)
            x = amg.cg(A, b, Pl=amg.aspreconditioner(ml, cyc), reltol=1e-8)
            assert np.linalg.norm(b - A.matvec(x)) <= 1e-8 * np.linalg.norm(b)
        ml.release()


def test_pcg_vs_oracle(amg, fx):
    A = amg.poisson((40, 40))
    b = _rng(9).random(A.n)
    ml = amg.smoothed_aggregation(A)
    H = oracle.OracleHierarchy(ml)
    x, info = amg.cg(A, b, Pl=amg.aspreconditioner(ml), reltol=1e-10, log=True)
    xr, histr = H.pcg(b, reltol=1e-10, log=True)
    assert info["iters"] == H.iters
    assert np.allclose(info["resnorm"], histr, rtol=1e-5, atol=0)
    assert np.linalg.norm(x - xr) <= 1e-9 * np.linalg.norm(xr)
    ml.release()


def test_elasticity_nns_device(amg, fx):
    # test/nns_test.jl:214-223 (config C5): converges with B, not without
    A, b, B = fx.matrix("elastic"), fx.array("elastic_b"), fx.array("elastic_B")
    ml = amg.smoothed_aggregation(A, B=B)
    x, res = amg._solve(ml, b, log=True, reltol=1e-10)
    xr, resr = oracle.OracleHierarchy(ml).solve(b, log=True, reltol=1e-10)
    assert len(res) == len(resr) and np.linalg.norm(x - xr) <= TOL_SOLVE_X * np.linalg.norm(xr)
    assert np.linalg.norm(A.matvec(x) - b) <= np.sqrt(np.finfo(float).eps) * np.linalg.norm(b)     # Julia's `≈`
    xc = amg.cg(A, b, Pl=amg.aspreconditioner(ml), reltol=1e-10)
    assert np.linalg.norm(A.matvec(xc) - b) <= 1e-9 * np.linalg.norm(b)
    ml2 = amg.smoothed_aggregation(A, coarse_solver=amg.Pinv)
    x2, res2 = amg._solve(ml2, b, log=True, reltol=1e-10)
    assert not np.linalg.norm(A.matvec(x2) - b) <= np.sqrt(np.finfo(float).eps) * np.linalg.norm(b)
    assert res2[0] > res2[-1]


def test_regression_95_nonsymmetric_device(amg, fx):
    # test/test_regression.jl:71-83
    N = 10000
    A = fx.sprand_plus_diag(N, 0.001, 5.0, seed=5)
    b = np.ones(N)
    for f in (amg.ruge_stuben, amg.smoothed_aggregation):
        ml = f(A, symmetry=amg.NoSymmetry())
        x = amg._solve(ml, b)
        assert np.allclose(A.matvec(x), b, rtol=1e-8)
        xr = oracle.OracleHierarchy(ml).solve(b)
        assert np.linalg.norm(x - xr) <= TOL_SOLVE_X * np.linalg.norm(xr)
        ml.release()


def test_no_level_hierarchy_and_quirks(amg):
    # test/test_regression.jl:41-57 + SURVEY appendix C
    for n in (1, 3, 10):
        A = amg.poisson(n)
        ml = amg.ruge_stuben(A)
        assert len(ml) == 1
        b = A.matvec(np.ones(n))
        assert np.allclose(amg._solve(ml, b), np.ones(n))
    A = amg.poisson(100)
    ml = amg.ruge_stuben(A)
    x = np.ones(100)
    x, res = amg._solve_(x, ml, np.zeros(100), log=True)
    assert np.array_equal(x, np.ones(100)) and list(res) == [0.0]          # ||b|| = 0: loop never runs
    before = ml.device().launch_count()
    amg._solve(ml, np.ones(100), maxiter=3, calculate_residual=False)
    assert ml.device().launch_count() > before
    x32 = amg._solve(ml, np.ones(100, dtype=np.float32))
    assert x32.dtype == np.float64                                           # promote(eltype(A), eltype(b)) runtests.jl:244-259


def test_precision_and_fp32_storage(amg, jac, monkeypatch):
    """test/runtests.jl:244-259 ("Precision"): eltype(_solve(ml, b)) = promote(eltype(A), eltype(b)) for every Float32 / Float64
    mix — and the device side of it: operators whose values are exactly representable in binary32 (a stencil; EVERY level of a
    hierarchy built from a Float32 matrix) are read as 4-byte values by the bandwidth kernels, with bit-identical results."""
    a = amg.poisson(100)
    b = _rng(5).random(100)
    for T, V in ((np.float64, np.float64), (np.float32, np.float32), (np.float64, np.float32), (np.float32, np.float64)):
        ml = amg.smoothed_aggregation(a.astype(T))
        x = amg._solve(ml, b.astype(V))
        assert x.dtype == np.promote_types(T, V), (T, V, x.dtype)
        ml.release()
    # lossless narrow storage on a Float64 hierarchy: the stencil level qualifies, the Galerkin products do not
    monkeypatch.setenv("B200AMG_FP32_STORAGE", "1")              # (read when a hierarchy is uploaded; off by default)
    A = amg.poisson((40, 40, 40))
    ml = amg.ruge_stuben(A, presmoother=jac, postsmoother=jac)
    dev = ml.device()
    assert dev.storage_info(0)["A"] == 4 and dev.storage_info(1)["A"] == 8
    r = _rng(6)
    x0, bb = r.standard_normal(A.n), r.standard_normal(A.n)
    got = {}
    for on in (1, 0):
        dev.set_option(17, on)                                   # B200AMG_OPT_FP32_STORAGE
        assert dev.storage_info(0)["A"] == (4 if on else 8)
        y = dev.apply(0, 0, np.empty(A.n), x0)
        res = dev.residual(0, np.empty(A.n), bb, x0)
        xs, hist = amg._solve(ml, bb, log=True, maxiter=12)
        got[on] = (y, res, xs, hist)
    for u, v in zip(got[1], got[0]):
        assert np.array_equal(u, v)                              # the same fp64 products and sums: identical bits
    ml.release()
    # a Float32 hierarchy: 4-byte values on every level, fp64 arithmetic; against the oracle on the same (Float32-valued) operators
    A32 = amg.poisson((24, 24, 24)).astype(np.float32)
    for build, kw in ((amg.ruge_stuben, dict(presmoother=jac, postsmoother=jac)), (amg.smoothed_aggregation, dict(presmoother=jac, postsmoother=jac)),
                      (amg.ruge_stuben, {})):
        ml = build(A32, **kw)
        dev = ml.device()
        assert all(dev.storage_info(i)["A"] == 4 and dev.storage_info(i)["P"] == 4 and dev.storage_info(i)["R"] == 4 for i in range(len(ml.levels)))
        b32 = r.random(A32.n).astype(np.float32)
        x, hist = amg._solve(ml, b32, log=True, reltol=1e-7)          # (the default tolerance is sqrt(eps(Float32)) here: multilevel.jl:160)
        assert x.dtype == np.float32
        xr, histr = oracle.OracleHierarchy(ml).solve(b32.astype(np.float64), log=True, reltol=1e-7)
        assert len(hist) == len(histr) and np.allclose(hist, histr, rtol=TOL_HIST)
        assert np.linalg.norm(x - xr) <= 1e-6 * np.linalg.norm(xr)          # x was rounded to Float32 at the boundary
        ml.release()


def test_multicolour_gauss_seidel_is_an_explicit_non_parity_mode(amg, monkeypatch):
    """B200AMG_GS_MULTICOLOR=1 ("offer multicolor as an explicitly non-parity fast mode"): the sweep relaxes colour
    after colour of a greedy colouring instead of in index order.  On the 7-point stencil the colouring is red-black, so one
    forward sweep must equal red-black Gauss-Seidel computed directly; as a smoother it converges to the same solution in a
    comparable number of iterations, but its iterates are NOT the reference's."""
    monkeypatch.setenv("B200AMG_GS_MULTICOLOR", "1")
    dims = (14, 12, 10)
    A = amg.poisson(dims)
    S = A.to_scipy().tocsr()
    r = _rng(21)
    x0, b = r.standard_normal(A.n), r.standard_normal(A.n)
    x = x0.copy()
    amg.GaussSeidel(amg.ForwardSweep())(A, x, b, amg.HermitianSymmetry())
    i, j, k = np.meshgrid(np.arange(dims[0]), np.arange(dims[1]), np.arange(dims[2]), indexing="ij")
    colour = ((i + j + k) % 2).ravel(order="F")                     # first index fastest (gallery.jl:42-63)
    d = S.diagonal()
    ref = x0.copy()
    for c in (0, 1):
        rows = np.nonzero(colour == c)[0]
        ref[rows] = (b[rows] - (S[rows] @ ref - d[rows] * ref[rows])) / d[rows]
    assert relinf(x, ref) <= 1e-13
    parity = oracle.smooth(A, amg.GaussSeidel(amg.ForwardSweep()), x0.copy(), b)
    assert relinf(x, parity) > 1e-3                                  # a different ordering: not the reference's iterates
    A = amg.poisson((28, 28, 28))
    bb = A.matvec(np.ones(A.n))
    ml = amg.ruge_stuben(A)
    xs, hist = amg._solve(ml, bb, log=True)
    assert hist[-1] <= 1.5e-8 * hist[0] and np.abs(xs - 1.0).max() <= 1e-6
    monkeypatch.delenv("B200AMG_GS_MULTICOLOR")
    mlp = amg.ruge_stuben(A)
    _, histp = amg._solve(mlp, bb, log=True)
    assert len(histp) <= len(hist) <= 2 * len(histp)                 # a somewhat weaker smoother than the lexicographic sweep (11 vs 7 entries here)
    ml.release()
    mlp.release()


# ---- larger sizes: properties that need no oracle run -------------------------------------------
def test_properties_at_size(amg, jac):
    A = amg.poisson((96, 96, 96))
    n = A.n
    ml = amg.ruge_stuben(A)
    dev = ml.device()
    r = _rng(12)
    x, y = r.standard_normal(n), r.standard_normal(n)
    # linearity of the SpMV and agreement with the closed-form stencil
    ax, ay, axy = dev.apply(0, 0, np.empty(n), x), dev.apply(0, 0, np.empty(n), y), dev.apply(0, 0, np.empty(n), 2 * x - 3 * y)
    assert relinf(axy, 2 * ax - 3 * ay) <= 1e-12
    X = x.reshape(96, 96, 96)
    st = 6 * X.copy()
    for ax_ in range(3):
        sl_lo = [slice(None)] * 3
        sl_hi = [slice(None)] * 3
        sl_lo[ax_], sl_hi[ax_] = slice(0, -1), slice(1, None)
        st[tuple(sl_hi)] -= X[tuple(sl_lo)]
        st[tuple(sl_lo)] -= X[tuple(sl_hi)]
    assert relinf(ax, st.reshape(-1)) <= TOL_KERNEL
    # A*ones has the known solution ones; the solve converges to it; history is monotone
    b = A.matvec(np.ones(n))
    xs, hist = amg._solve(ml, b, log=True)
    assert hist[-1] <= np.sqrt(np.finfo(float).eps) * hist[0] and np.all(np.diff(hist) < 0)
    assert np.abs(xs - 1).max() < 1e-4          # reltol = sqrt(eps) on the residual, not on the error
    # R = P' : <R r, e> == <r, P e>
    lv = ml.levels[0]
    e = r.standard_normal(lv.R.shape[0])
    Rr, Pe = dev.apply(0, 2, np.empty(lv.R.shape[0]), x), dev.apply(0, 1, np.empty(n), e)
    assert abs(Rr @ e - x @ Pe) <= 1e-10 * abs(x @ Pe)
    # a converged x is a fixed point of every smoother up to the residual
    xx = xs.copy()
    dev.smooth(0, 0, xx, b)
    assert np.abs(xx - xs).max() < 1e-4
    ml.release()


# ---- every Gauss-Seidel sweep protocol gives the reference's sequential sweep ---------------------
@pytest.mark.parametrize("mode,env", [(0, {}), (1, {}), (2, {}), (3, {}), (2, {"cta_rows": 0, "mail_width": 1})])
def test_all_sweep_protocols_agree_with_oracle(amg, monkeypatch, mode, env):
    """gs_mode 0 one launch per wavefront, 1 wavefront-counter dataflow, 2 TMA-fed mailbox (+ single-CTA on small
    levels), 3 ticket mailbox; the last case forces the mailbox sweeps onto every level.  (The blocked sweep, which would
    otherwise take the stencil level and the tiny ones, is switched off: it has its own test below.)"""
    monkeypatch.setenv("B200AMG_GS_BLOCK", "0")
    A = amg.poisson((28, 28, 28))
    ml = amg.ruge_stuben(A)
    dev = ml.device()
    dev.set_option(0, 0)            # no graph replay: options take effect on every launch
    dev.set_option(3, mode)
    if "cta_rows" in env:
        dev.set_option(7, env["cta_rows"])
    if "mail_width" in env:
        dev.set_option(8, env["mail_width"])
    r = _rng(21)
    for lv, level in enumerate(ml.levels):
        x0, b = r.standard_normal(level.A.n), r.standard_normal(level.A.n)
        x = dev.smooth(lv, 0, x0.copy(), b)
        assert relinf(x, oracle.smooth(level.A, level.presmoother.config, x0.copy(), b)) <= TOL_SWEEP, (mode, lv)
    b = r.random(A.n)
    x = amg._solve(ml, b, maxiter=3, calculate_residual=False)
    ref = oracle.OracleHierarchy(ml).solve(b, maxiter=3, calculate_residual=False)
    assert relinf(x, ref) <= TOL_CYCLE
    ml.release()


# ---- BASELINE.json's full size (config C3: 256^3, RS, default symmetric Gauss-Seidel): properties only ----
def test_full_size_256_properties(amg):
    n1 = 256
    A = amg.poisson((n1, n1, n1))
    n = A.n
    assert n == 16777216 and A.nnz == 117047296
    ml = amg.ruge_stuben(A)
    dev = ml.device()
    r = _rng(256)
    x = r.standard_normal(n)
    # SpMV against the closed-form 7-point stencil (first index fastest, gallery.jl:5-63)
    ax = dev.apply(0, 0, np.empty(n), x)
    X = x.reshape(n1, n1, n1)
    st = 6 * X
    for ax_ in range(3):
        lo, hi = [slice(None)] * 3, [slice(None)] * 3
        lo[ax_], hi[ax_] = slice(0, -1), slice(1, None)
        st[tuple(hi)] -= X[tuple(lo)]
        st[tuple(lo)] -= X[tuple(hi)]
    assert relinf(ax, st.reshape(-1)) <= TOL_KERNEL
    del X, st
    # R = P' on the fine level
    nc = ml.levels[0].R.shape[0]
    e = r.standard_normal(nc)
    Rr, Pe = dev.apply(0, 2, np.empty(nc), x), dev.apply(0, 1, np.empty(n), e)
    assert abs(Rr @ e - x @ Pe) <= 1e-10 * abs(x @ Pe)
    # one exact-order symmetric Gauss-Seidel sweep on the fine level: x = ones is a fixed point for b = A*ones, and the
    # sweep is a contraction in the energy norm for any start (SPD matrix)
    b = A.matvec(np.ones(n))
    ones = np.ones(n)
    assert np.abs(dev.smooth(0, 0, ones.copy(), b) - 1.0).max() <= 1e-13
    y = dev.smooth(0, 0, x.copy(), b)
    err0, err1 = x - 1.0, y - 1.0
    e0 = err0 @ dev.apply(0, 0, np.empty(n), err0)
    e1 = err1 @ dev.apply(0, 0, np.empty(n), err1)
    assert 0 < e1 < e0
    # the solve converges to the known solution with a strictly decreasing residual history
    xs, hist = amg._solve(ml, b, log=True, maxiter=60)
    assert hist[-1] <= np.sqrt(np.finfo(float).eps) * hist[0] and np.all(np.diff(hist) < 0)
    assert np.abs(xs - 1).max() < 1e-3
    rr = dev.residual(0, np.empty(n), b, xs)
    assert abs(np.linalg.norm(rr) - hist[-1]) <= 1e-9 * hist[0]
    # ---- BASELINE config C3 at FULL size against the oracle: 3 `_solve!` iterations from x0 = 0 (the size-selected sweep
    # kernels only all run together here; a stale read in a sweep still converges, so properties alone would not catch it)
    H = oracle.OracleHierarchy(ml)
    xo, ho = H.solve(b, maxiter=3, reltol=0.0, log=True)
    xg, hg = amg._solve(ml, b, log=True, maxiter=3, reltol=0.0)
    assert len(hg) == len(ho) == 4
    assert np.max(np.abs(hg - ho) / ho) <= TOL_HIST and relinf(xg, xo) <= TOL_SOLVE_X
    # one symmetric Gauss-Seidel sweep on each of the three largest levels, repeated: every repetition must land on the
    # oracle's sweep (the hand-offs between CTAs are exercised ~50 000 times per repetition)
    for lvl in range(3):
        Al = ml.levels[lvl].A
        xl, bl = _rng(lvl).standard_normal(Al.n), _rng(10 + lvl).standard_normal(Al.n)
        ref = oracle.smooth(Al, amg.GaussSeidel(), xl.copy(), bl)
        for rep in range(8 if lvl == 0 else 4):
            got = dev.smooth(lvl, 0, xl.copy(), bl)
            assert relinf(got, ref) <= TOL_SWEEP, (lvl, rep)
    del H
    ml.release()
    # ---- BASELINE config C4's workload (same matrix, Jacobi) on one GPU: 3 iterations against the oracle
    jac = amg.Jacobi(2.0 / 3.0)
    ml = amg.ruge_stuben(A, presmoother=jac, postsmoother=jac)
    H = oracle.OracleHierarchy(ml)
    xo, ho = H.solve(b, maxiter=3, reltol=0.0, log=True)
    xg, hg = amg._solve(ml, b, log=True, maxiter=3, reltol=0.0)
    assert np.max(np.abs(hg - ho) / ho) <= TOL_HIST and relinf(xg, xo) <= TOL_SOLVE_X
    del H
    ml.release()


def test_full_size_c2_sa_jacobi_1024x1024_vs_oracle(amg):
    """BASELINE config C2 at full size: 2-D poisson((1024, 1024)), smoothed aggregation + Jacobi, against the oracle."""
    A = amg.poisson((1024, 1024))
    assert A.n == 1048576 and A.nnz == 5238784
    jac = amg.Jacobi(2.0 / 3.0)
    ml = amg.smoothed_aggregation(A, presmoother=jac, postsmoother=jac)
    b = A.matvec(np.ones(A.n))
    H = oracle.OracleHierarchy(ml)
    xo, ho = H.solve(b, maxiter=10, reltol=0.0, log=True)
    xg, hg = amg._solve(ml, b, log=True, maxiter=10, reltol=0.0)
    assert len(hg) == len(ho) == 11
    assert np.max(np.abs(hg - ho) / ho) <= TOL_HIST and relinf(xg, xo) <= TOL_SOLVE_X
    ml.release()


def test_block_right_hand_sides(amg):
    # bs > 1 workspaces (multilevel.jl:28-59): columns share one Frobenius-norm convergence test
    A = amg.poisson((30, 30))
    ml = amg.ruge_stuben(A)
    B = _rng(31).random((A.n, 3))
    X, hist = amg._solve(ml, B, log=True, reltol=1e-10)
    assert X.shape == B.shape
    H = oracle.OracleHierarchy(ml)
    Xr = np.zeros_like(B)
    histr = [np.linalg.norm(B)]
    tol = 1e-10 * histr[0]
    while len(histr) <= 100 and histr[-1] > tol:
        for j in range(3):
            Xr[:, j] = H.cycle(Xr[:, j].copy(), B[:, j])
        histr.append(np.linalg.norm(B - np.column_stack([A.matvec(Xr[:, j]) for j in range(3)])))
    assert len(hist) == len(histr) and np.allclose(hist, histr, rtol=TOL_HIST)
    assert np.linalg.norm(X - Xr) <= TOL_SOLVE_X * np.linalg.norm(Xr)
    ml.release()


def test_synthetic_elasticity_with_near_null_space(amg):
    # the reference's elasticity behaviour (test/nns_test.jl:214-223) on the synthetic generator: SA with the rigid-body
    # modes converges and is a good CG preconditioner; the device PCG matches the oracle PCG
    A, b, B = amg.elasticity_2d(40, 24)
    ml = amg.smoothed_aggregation(A, B=B)
    H = oracle.OracleHierarchy(ml)
    x, hist = amg._solve(ml, b, log=True, reltol=1e-10)
    xr, histr = H.solve(b, log=True, reltol=1e-10)
    assert len(hist) == len(histr) and len(hist) < 100
    assert np.linalg.norm(x - xr) <= TOL_SOLVE_X * np.linalg.norm(xr)
    xc, info = amg.cg(A, b, Pl=amg.aspreconditioner(ml), reltol=1e-10, log=True)
    xcr = H.pcg(b, reltol=1e-10)
    assert info["iters"] == H.iters and np.linalg.norm(xc - xcr) <= 1e-8 * np.linalg.norm(xcr)
    assert np.linalg.norm(A.matvec(xc) - b) <= 1e-9 * np.linalg.norm(b)
    ml.release()


def test_synthetic_elasticity_3d_with_rigid_body_modes(amg):
    # 3-D Q1 elasticity (81-entry rows, 3 dofs per node) with the six rigid-body modes as near-null-space: stand-alone solve
    # and device PCG against the oracle, Gauss-Seidel (default) and Jacobi smoothers
    A, b, B = amg.elasticity_3d(12, 10, 8)
    jac = amg.Jacobi(0.5)          # (omega = 2/3 diverges as a stand-alone iteration on this operator)
    for kw in ({}, {"presmoother": jac, "postsmoother": jac}):
        ml = amg.smoothed_aggregation(A, B=B, **kw)
        H = oracle.OracleHierarchy(ml)
        x, hist = amg._solve(ml, b, log=True, reltol=1e-10, maxiter=60)
        xr, histr = H.solve(b, log=True, reltol=1e-10, maxiter=60)
        assert len(hist) == len(histr) and np.allclose(hist, histr, rtol=TOL_HIST)
        assert np.linalg.norm(x - xr) <= TOL_SOLVE_X * np.linalg.norm(xr)
        xc, info = amg.cg(A, b, Pl=amg.aspreconditioner(ml), reltol=1e-10, log=True)
        xcr = H.pcg(b, reltol=1e-10)
        assert info["iters"] == H.iters and np.linalg.norm(xc - xcr) <= 1e-8 * np.linalg.norm(xcr)
        assert np.linalg.norm(A.matvec(xc) - b) <= 1e-9 * np.linalg.norm(b)
        ml.release()


@pytest.mark.parametrize("block", ["1", "2", "pass"])
def test_blocked_sweep_matches_oracle(amg, fx, monkeypatch, block):
    """The blocked exact-order sweep (block_gs.cuh: one CTA per tile of rows, shared-memory window, scouts that stage far
    values, cross-tile progress counters) against the reference's sequential sweeps: Gauss-Seidel and SOR, forward /
    backward / symmetric, on stencils of every dimension, irregular RS coarse operators, a finite-element matrix and
    multi-tile levels; B200AMG_GS_BLOCK=2 forces it onto every level of a hierarchy, =1 is the default per-level choice."""
    if block == "pass":      # the pass sweep (pass_gs.cuh) on every level's blocked plan
        monkeypatch.setenv("B200AMG_GS_PASS", "1")
        block = "2"
    monkeypatch.setenv("B200AMG_GS_BLOCK", block)
    ml0 = amg.ruge_stuben(amg.poisson((40, 40, 40)))
    mats = [amg.poisson(777), amg.poisson((40, 33)), amg.poisson((20, 17, 12)), amg.poisson((40, 40, 40)), fx.matrix("thing")]
    mats += [lv.A for lv in ml0.levels[1:4]]
    F_, B_, S_ = amg.ForwardSweep(), amg.BackwardSweep(), amg.SymmetricSweep()
    zoo = [amg.GaussSeidel(F_), amg.GaussSeidel(B_), amg.GaussSeidel(S_, 2), amg.SOR(0.5, F_), amg.SOR(1.3, S_, 2)]
    r = _rng(77)
    for A in mats:
        x0, b = r.standard_normal(A.n), r.standard_normal(A.n)
        for sm in zoo:
            x = x0.copy()
            sm(A, x, b, amg.HermitianSymmetry())
            ref = oracle.smooth(A, sm, x0.copy(), b)
            assert relinf(x, ref) <= TOL_SWEEP, (A.n, sm, block)
    # whole solves through it (every level relaxed by the blocked sweep when block == "2")
    for dims in [(28, 28, 28), (60, 50)]:
        A = amg.poisson(dims)
        ml = amg.ruge_stuben(A)
        b = r.random(A.n)
        x, hist = amg._solve(ml, b, log=True)
        xr, histr = oracle.OracleHierarchy(ml).solve(b, log=True)
        assert len(hist) == len(histr) and np.max(np.abs(hist - histr) / histr) <= TOL_HIST
        assert np.linalg.norm(x - xr) / np.linalg.norm(xr) <= TOL_SOLVE_X
        for cyc in (amg.W(), amg.F()):
            y = amg._solve(ml, b, cyc, maxiter=2, calculate_residual=False)
            yr = oracle.OracleHierarchy(ml).solve(b, cycle=type(cyc).__name__, maxiter=2, calculate_residual=False)
            assert relinf(y, yr) <= TOL_CYCLE
        ml.release()


def test_cluster_sweep_on_mid_size_level(amg):
    """The optional one-cluster sweep (x in distributed shared memory, cluster_gs.cuh; off by default because it measured
    no faster than the counter sweep) must give the reference's sequential sweep for every cluster shape."""
    ml = amg.ruge_stuben(amg.poisson((48, 48, 48)))
    lv = 1
    level = ml.levels[lv]
    assert 12288 < level.A.n <= 380000 and level.A.nnz / level.A.n >= 16     # (never a blocked-sweep level: long rows)
    dev = ml.device()
    dev.set_option(0, 0)
    r = _rng(48)
    x0, b = r.standard_normal(level.A.n), r.standard_normal(level.A.n)
    ref = oracle.smooth(level.A, level.presmoother.config, x0.copy(), b)
    for cluster, lognc, threads in ((1, 3, 256), (1, 4, 1024), (1, 1, 1024), (0, 3, 256)):
        dev.set_option(9, cluster)
        dev.set_option(10, lognc)
        dev.set_option(11, threads)
        x = dev.smooth(lv, 0, x0.copy(), b)
        assert relinf(x, ref) <= TOL_SWEEP, (cluster, lognc, threads)
    ml.release()


# ---- one-cluster sweep with x in distributed shared memory (csrc/device/dsm_gs.cuh) ----------------------
@pytest.mark.parametrize("log_nc,fence,two_groups", [(0, 0, 0), (1, 0, 0), (2, 0, 0), (3, 0, 0), (4, 0, 0), (2, 3, 0), (-1, 0, 0),
                                                     (0, 0, 1), (1, 0, 1), (2, 0, 1), (3, 0, 1), (4, 0, 1), (2, 3, 1), (-1, 0, 1)])
def test_dsm_cluster_sweep_matches_oracle(amg, fx, monkeypatch, log_nc, fence, two_groups):
    """Every cluster size (1..16 CTAs; -1 = chosen per level) of the distributed-shared-memory sweep — gs_dsm_kernel and the
    two-group gs_dsm2_kernel — gives the reference's sequential Gauss-Seidel / SOR sweeps (forward, backward, symmetric;
    Hermitian and NoSymmetry walks), on stencil rows, irregular RS coarse operators and a nonsymmetric matrix; then whole
    cycles through it."""
    monkeypatch.setenv("B200AMG_GS_BLOCK", "0")
    monkeypatch.setenv("B200AMG_GS_DSM", "2")
    monkeypatch.setenv("B200AMG_GS_DSM2", str(two_groups))
    monkeypatch.setenv("B200AMG_GS_DSM_MAX_CTAS_LOG2", "4")
    monkeypatch.setenv("B200AMG_GS_DSM_FENCE", str(fence))
    if log_nc >= 0:
        monkeypatch.setenv("B200AMG_GS_DSM_LOG_NC", str(log_nc))
    ml0 = amg.ruge_stuben(amg.poisson((24, 24, 24)))
    mats = [amg.poisson((20, 17, 12)), ml0.levels[1].A, ml0.levels[2].A, fx.sprand_plus_diag(400, 0.05, 5.0, seed=11)]
    F_, B_, S_ = amg.ForwardSweep(), amg.BackwardSweep(), amg.SymmetricSweep()
    zoo = [amg.GaussSeidel(F_), amg.GaussSeidel(B_), amg.GaussSeidel(S_, 2), amg.SOR(0.5, F_), amg.SOR(1.3, S_, 2)]
    r = _rng(31)
    for A in mats:
        x0, b = r.standard_normal(A.n), r.standard_normal(A.n)
        for symmetry in ("hermitian", "none"):
            sym = amg.HermitianSymmetry() if symmetry == "hermitian" else amg.NoSymmetry()
            for sm in zoo:
                x = x0.copy()
                sm(A, x, b, sym)
                ref = oracle.smooth(A, sm, x0.copy(), b, symmetry=symmetry)
                assert relinf(x, ref) <= TOL_SWEEP, (A.n, symmetry, sm, log_nc)
    A = amg.poisson((28, 28, 28))
    ml = amg.ruge_stuben(A)
    dev = ml.device()
    launches0 = dev.launch_count()
    b = r.random(A.n)
    x, hist = amg._solve(ml, b, log=True)
    xr, histr = oracle.OracleHierarchy(ml).solve(b, log=True)
    assert len(hist) == len(histr)
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) <= TOL_SOLVE_X
    assert dev.launch_count() > launches0
    ml.release()


# ---- setup phase: Galerkin products on the device (csrc/device/spgemm.cuh) ----------------------
def _same_csc(x, y):
    return (x.shape == y.shape and np.array_equal(x.colptr, y.colptr) and np.array_equal(x.rowval, y.rowval)
            and np.array_equal(x.nzval.view(np.int64), y.nzval.view(np.int64)))


def test_device_galerkin_product_is_bit_identical_to_host(amg, fx, monkeypatch):
    """`R*A` and `(R*A)*P` (classical.jl:46, aggregation.jl:145) through b200amg_spgemm_begin/_fetch: same pattern
    (sorted rows, structural zeros kept) and the same bits as the host product; several batches when the scratch budget
    is small; a hierarchy built with the device backend equals the host-built one level by level."""
    from algebraicmultigrid_jl_b200 import _devlib, _hostlib

    ml = amg.ruge_stuben(amg.poisson((16, 16, 16)))
    sa = amg.smoothed_aggregation(amg.poisson((40, 40)))
    pairs = []
    for lv in ml.levels[:3]:
        R = lv.R.materialize() if hasattr(lv.R, "materialize") else lv.R
        P = lv.P.materialize() if hasattr(lv.P, "materialize") else lv.P
        pairs += [(R, lv.A), (_hostlib.spgemm(R, lv.A), P)]
    lv = sa.levels[0]
    R = lv.R.materialize() if hasattr(lv.R, "materialize") else lv.R
    P = lv.P.materialize() if hasattr(lv.P, "materialize") else lv.P
    pairs += [(R, lv.A), (_hostlib.spgemm(R, lv.A), P)]
    nons = fx.sprand_plus_diag(300, 0.03, 1.0, seed=5)
    pairs += [(nons, nons), (nons, nons.transpose())]
    # exact cancellation: a structural zero must be kept
    import scipy.sparse as sp
    a = amg.SparseMatrixCSC.from_scipy(sp.csc_matrix(np.array([[1.0, -1.0], [2.0, 3.0]])))
    b = amg.SparseMatrixCSC.from_scipy(sp.csc_matrix(np.array([[1.0, 0.0], [1.0, 4.0]])))
    pairs.append((a, b))
    for budget in ("192", "1"):
        monkeypatch.setenv("B200AMG_SPGEMM_SLOTS_M", budget)
        for a_, b_ in pairs:
            h = _hostlib.spgemm(a_, b_)
            d = _devlib.spgemm(a_, b_)
            assert _same_csc(h, d), (a_.shape, b_.shape, budget)
    c = _devlib.spgemm(a, b)
    assert c.nnz == 4 and 0.0 in list(c.nzval)          # (1)(1) + (-1)(1) = 0 stays stored
    monkeypatch.delenv("B200AMG_SPGEMM_SLOTS_M")
    try:
        amg.set_galerkin_backend("device")
        mld = amg.ruge_stuben(amg.poisson((16, 16, 16)))
        sad = amg.smoothed_aggregation(amg.poisson((40, 40)))
    finally:
        amg.set_galerkin_backend("host")
    for h_, d_ in ((ml, mld), (sa, sad)):
        assert len(h_.levels) == len(d_.levels)
        for lh, ld in zip(h_.levels, d_.levels):
            assert _same_csc(lh.A, ld.A)
        assert _same_csc(h_.final_A, d_.final_A)


# ---- coarse solver as a host callable (coarse_solver.jl:24-58,66-81: any callable; sparse factorisation of ANY size) ---------
def test_coarse_solver_as_host_callable(amg, monkeypatch):
    """A coarsest level beyond the dense-operator limit (coarsening stopped at max_levels) is solved by a sparse LU on the
    HOST from inside the cycle (b200amg_set_coarse_callback: D2H copy, host node, H2D copy — also inside the captured cycle
    graph); the result must agree with the oracle run on the same hierarchy with the dense inverse.  A user callable works
    the same way, and a failing callable surfaces as an error, never as a silent result."""
    from algebraicmultigrid_jl_b200 import _devlib, coarse_solver as cs_mod

    A = amg.poisson((24, 24, 24))
    b = oracle.mul(A, np.ones(A.n))
    ml_dense = amg.ruge_stuben(A, max_levels=2)                  # coarsest level: 6912 rows, dense inverse
    nf = ml_dense.final_A.n
    assert len(ml_dense.levels) == 1 and nf > 1000
    xo, ro = oracle.OracleHierarchy(ml_dense).solve(b, log=True, maxiter=8)
    monkeypatch.setattr(cs_mod, "DENSE_LIMIT", 1000)
    for coarse in (None, amg.LinearSolveWrapper(amg.UMFPACKFactorization())):
        kw = {} if coarse is None else {"coarse_solver": coarse}
        ml = amg.ruge_stuben(A, max_levels=2, **kw)
        assert ml.coarse_solver.dense_operator() is None         # sparse factorisation kept on the host
        xd, rd = amg._solve(ml, b, log=True, maxiter=8)
        assert len(rd) == len(ro)
        assert np.abs(rd / ro - 1).max() <= TOL_HIST
        assert np.linalg.norm(xd - xo) / np.linalg.norm(xo) <= TOL_SOLVE_X
        # preconditioner application and the stand-alone coarse solve go through the same callable
        dev = ml.device()
        bf = _rng(4).standard_normal(nf)
        assert relinf(dev.coarse_solve(np.empty(nf), bf), ml_dense.coarse_solver.dense_operator() @ bf) <= 1e-9
        # W cycle: the coarse solve is reached once per visit of the last level
        xw = amg._solve(ml, b, amg.W(), maxiter=3)
        xwo = oracle.OracleHierarchy(ml_dense).solve(b, cycle="W", maxiter=3)
        assert np.linalg.norm(xw - xwo) / np.linalg.norm(xwo) <= TOL_SOLVE_X
    monkeypatch.setattr(cs_mod, "DENSE_LIMIT", 16384)

    # any Python callable cs(x, b)
    dense = np.linalg.inv(ml_dense.final_A.todense())
    calls = []

    class Mine:
        def __init__(self, Ac):
            pass

        def __call__(self, x, bb):
            calls.append(bb.shape)
            x[...] = dense @ bb

    ml = amg.ruge_stuben(A, max_levels=2, coarse_solver=Mine)
    xd = amg._solve(ml, b, maxiter=8)
    assert calls and calls[0] == (nf,)
    assert np.linalg.norm(xd - xo) / np.linalg.norm(xo) <= TOL_SOLVE_X

    class Broken(Mine):
        def __call__(self, x, bb):
            raise RuntimeError("factorisation lost")

    ml = amg.ruge_stuben(A, max_levels=2, coarse_solver=Broken)
    with pytest.raises(_devlib.B200AmgError, match="callback"):
        amg._solve(ml, b, maxiter=2)
