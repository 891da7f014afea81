"""Pin the CPU oracle (oracle/amg_oracle.c) against every known-answer the reference's own tests
hold for the solve phase.  CPU only.  Tolerances are the reference's own (Σdiff² < 1e-8 etc.)."""
import numpy as np
import pytest

import goldens
import oracle


def _isapprox(a, b):
    """Julia's `a ≈ b` for vectors: ||a-b|| <= sqrt(eps) * max(||a||, ||b||)."""
    return np.linalg.norm(a - b) <= np.sqrt(np.finfo(float).eps) * max(np.linalg.norm(a), np.linalg.norm(b))


def _thing_b():
    b = np.zeros(46)
    b[0], b[1] = 1, -1
    return b


# ---- smoothers --------------------------------------------------------------------------------
def test_gauss_seidel_known_answers(amg):
    # test/sa_tests.jl:316-379 — exact rationals
    fwd, bwd = amg.GaussSeidel(amg.ForwardSweep()), amg.GaussSeidel(amg.BackwardSweep())
    A1, A3 = amg.poisson(1), amg.poisson(3)
    assert np.sum(oracle.smooth(A1, fwd, [0.0], [0.0]) ** 2) < 1e-8
    assert np.sum((oracle.smooth(A3, fwd, [0.0, 1, 2], np.zeros(3)) - [1 / 2, 5 / 4, 5 / 8]) ** 2) < 1e-8
    assert np.sum(oracle.smooth(A1, bwd, [0.0], [0.0]) ** 2) < 1e-8
    assert np.sum((oracle.smooth(A3, bwd, [0.0, 1, 2], np.zeros(3)) - [1 / 8, 1 / 4, 1 / 2]) ** 2) < 1e-8
    assert np.sum((oracle.smooth(A1, fwd, [0.0], [10.0]) - [5.0]) ** 2) < 1e-8
    assert np.sum((oracle.smooth(A3, fwd, [0.0, 1, 2], [10.0, 20, 30]) - [11 / 2, 55 / 4, 175 / 8]) ** 2) < 1e-8
    # the rationals are exactly representable: demand bit equality too
    assert np.array_equal(oracle.smooth(A3, fwd, [0.0, 1, 2], np.zeros(3)), [0.5, 1.25, 0.625])
    A100 = amg.poisson(100)
    x1 = oracle.smooth(A100, amg.GaussSeidel(amg.ForwardSweep(), 200), np.ones(100), np.zeros(100))
    x2 = oracle.smooth(A100, amg.GaussSeidel(amg.BackwardSweep(), 200), np.ones(100), np.zeros(100))
    r1, r2 = np.linalg.norm(A100.matvec(x1)), np.linalg.norm(A100.matvec(x2))
    assert r1 < 0.01 and r2 < 0.01 and np.isclose(r1, r2)


def test_regression_26(amg):
    x = oracle.smooth(amg.poisson(10), amg.GaussSeidel(amg.SymmetricSweep(), 4), np.ones(10), np.zeros(10))
    assert np.sum((x - goldens.SGS4_POISSON10) ** 2) < 1e-6


def test_nosymmetry_smoothers_converge(amg, fx):
    # test/test_smoothers.jl:15-27
    N = 50
    A = fx.sprand_plus_diag(N, 0.05, 5.0, seed=1)
    x0 = np.random.default_rng(2).random(N)
    b = np.ones(N)
    for sm in [amg.Jacobi(1 / 6, iter=500), amg.GaussSeidel(amg.ForwardSweep(), 100), amg.GaussSeidel(amg.BackwardSweep(), 100),
               amg.GaussSeidel(amg.SymmetricSweep(), 100), amg.SOR(0.5, amg.ForwardSweep(), 100),
               amg.SOR(0.5, amg.BackwardSweep(), 100), amg.SOR(0.5, amg.SymmetricSweep(), 100)]:
        x = oracle.smooth(A, sm, x0.copy(), b, symmetry="none")
        assert np.allclose(A.matvec(x), b), sm


def test_fast_equals_general_on_symmetric(amg):
    # test/test_smoothers.jl:29-45
    N = 50
    A = amg.poisson(N)
    x0 = np.random.default_rng(3).random(N)
    b = np.ones(N)
    for sm in [amg.Jacobi(4 / 5, iter=2), amg.GaussSeidel(amg.SymmetricSweep(), iter=2), amg.SOR(0.5, iter=2)]:
        xf = oracle.smooth(A, sm, x0.copy(), b, symmetry="hermitian")
        xg = oracle.smooth(A, sm, x0.copy(), b, symmetry="none")
        assert np.allclose(xf, xg), sm


def test_singular_exception(amg):
    import scipy.sparse as sp

    A = amg.SparseMatrixCSC.from_scipy(sp.csc_matrix(np.array([[1.0, 2.0], [3.0, 0.0]])))
    with pytest.raises(ZeroDivisionError):
        oracle.smooth(A, amg.GaussSeidel(amg.ForwardSweep()), np.ones(2), np.ones(2), symmetry="none")


# ---- SpMV second opinion ----------------------------------------------------------------------
def test_mul_vs_scipy(amg, fx):
    A = fx.matrix("test")
    x = np.random.default_rng(0).random(A.n)
    assert np.allclose(oracle.mul(A, x), A.to_scipy() @ x, rtol=1e-14)
    assert np.allclose(oracle.mul(A, x, adjoint=True), A.to_scipy().T @ x, rtol=1e-14)


# ---- cycle -------------------------------------------------------------------------------------
def test_solver_poisson1000(amg, fx):
    # test/runtests.jl:112-141
    A = amg.poisson(1000)
    b = A.matvec(np.ones(1000))
    x = oracle.OracleHierarchy(amg.ruge_stuben(A)).solve(b)
    assert np.sum((x - 1) ** 2) < 1e-8
    fs = amg.GaussSeidel(amg.ForwardSweep())
    x = oracle.OracleHierarchy(amg.ruge_stuben(A, presmoother=fs, postsmoother=fs)).solve(b)
    assert np.sum((x - 1) ** 2) < 1e-8
    x = oracle.OracleHierarchy(amg.ruge_stuben(A, coarse_solver=amg.LinearSolveWrapper(amg.UMFPACKFactorization()))).solve(b)
    assert np.sum((x - 1) ** 2) < 1e-7
    A = fx.matrix("randlap")
    b = A.matvec(np.ones(100))
    x = oracle.OracleHierarchy(amg.ruge_stuben(A, presmoother=fs, postsmoother=fs)).solve(b)
    assert np.sum(x ** 2) < 1e-8
    x = oracle.OracleHierarchy(amg.ruge_stuben(A)).solve(b)
    assert np.sum(x ** 2) < 1e-6


def test_thing_goldens(amg, fx):
    # test/runtests.jl:143-224
    A = fx.matrix("thing")
    n = A.m
    sm = amg.GaussSeidel(amg.ForwardSweep())
    ml = amg.ruge_stuben(A, presmoother=sm, postsmoother=sm, coarse_solver=amg.Pinv)
    H = oracle.OracleHierarchy(ml)
    b = _thing_b()
    x = H.solve(A.matvec(np.ones(n)), maxiter=1, abstol=1e-12)
    assert np.sum((x - goldens.THING_ZERO_GOLDEN) ** 2) < 1e-8
    x = H.solve(b, maxiter=1, abstol=1e-12)
    assert np.sum((x - goldens.THING_FWDGS_ONE_CYCLE) ** 2) < 1e-8
    x = H.pcg(b)
    assert np.sum((x - goldens.THING_CG_FWDGS) ** 2) < 1e-8
    ml = amg.ruge_stuben(A, coarse_solver=amg.Pinv)
    H = oracle.OracleHierarchy(ml)
    x = H.pcg(b, maxiter=100_000, reltol=1e-6)
    assert np.sum((x - goldens.THING_CG_SGS) ** 2) < 1e-8
    x = H.solve(b, maxiter=1, reltol=1e-12)
    assert np.sum((x - goldens.THING_SGS_ONE_CYCLE) ** 2) < 1e-8


def test_cycles(amg):
    # test/cycle_tests.jl:6-30
    A = amg.poisson((50, 50))
    b = A.matvec(np.ones(A.n))
    reltol = 1e-8
    for method in (amg.ruge_stuben, amg.smoothed_aggregation):
        H = oracle.OracleHierarchy(method(A))
        for cyc in "VWF":
            x, hist = H.solve(b, cycle=cyc, reltol=reltol, log=True)
            assert np.linalg.norm(b - A.matvec(x)) < reltol * np.linalg.norm(b)
        for cyc in "VWF":
            x = H.pcg(b, cycle=cyc, reltol=reltol)
            assert np.linalg.norm(b - A.matvec(x)) <= reltol * np.linalg.norm(b)


def test_preconditioned_cg_grids(amg):
    # test/runtests.jl:227-240 (LinearSolve `precs`: RS and SA preconditioned Krylov CG on three grids -> ones, rtol 1e-8);
    # the Krylov solver is third-party there, here the oracle's PCG with the same preconditioner semantics (ldiv!: zero
    # x, one cycle, preconditioner.jl:12-19)
    for sz in ((10, 10), (20, 20), (50, 50)):
        A = amg.poisson(sz)
        u0 = np.ones(A.n)
        b = A.matvec(u0)
        for builder in (amg.RugeStubenPreconBuilder(), amg.SmoothedAggregationPreconBuilder()):   # src/precs.jl
            Pl, Pr = builder(A, None)
            assert isinstance(Pl, amg.Preconditioner) and repr(Pr) == "I" and Pl.ml.levels[0].A is A
            x = oracle.OracleHierarchy(Pl.ml).pcg(b, reltol=1e-10)
            assert np.allclose(x, u0, rtol=1e-8, atol=0.0), (sz, type(builder).__name__, np.abs(x - u0).max())
    # builder keywords reach the setup (precs.jl:12-17,31-36), the block size reaches the workspace
    Pl, _ = amg.RugeStubenPreconBuilder(blocksize=2, max_levels=3)(amg.poisson((20, 20)), None)
    assert len(Pl.ml) <= 3 and Pl.ml.workspace.bs == 2


def test_regression_56(amg):
    # test/test_regression.jl:59-69
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    X = amg.SparseMatrixCSC.from_scipy(amg.poisson(27000).to_scipy() + 24.0 * sp.identity(27000, format="csc"))
    b = np.random.default_rng(0).random(27000)
    ref = spl.spsolve(X.to_scipy(), b)
    x = oracle.OracleHierarchy(amg.ruge_stuben(X)).solve(b, reltol=1e-10)
    assert np.allclose(x, ref, rtol=1e-10, atol=1e-10 * np.linalg.norm(ref, np.inf))
    # SA half: with θ = 0.05 every off-diagonal of poisson+24I is weak, nothing aggregates and the
    # whole matrix becomes the coarsest level (the reference hands it to sparse QR).  Here the coarse
    # operator is a dense inverse, so this half runs at n = 2700 (same structure, 100x less dense work).
    n2 = 2700
    X2 = amg.SparseMatrixCSC.from_scipy(amg.poisson(n2).to_scipy() + 24.0 * sp.identity(n2, format="csc"))
    ml = amg.smoothed_aggregation(X2, strength=amg.SymmetricStrength(0.05))
    assert len(ml) == 1
    ref2 = spl.spsolve(X2.to_scipy(), b[:n2])
    x = oracle.OracleHierarchy(ml).solve(b[:n2], reltol=1e-10)
    assert np.linalg.norm(x - ref2) <= 1e-10 * np.linalg.norm(ref2) * 10


def test_regression_46(amg, fx):
    # test/test_regression.jl:25-39 (bug.jld2)
    a = fx.matrix("bug")
    b = np.zeros(4)
    b[0], b[1] = 1, -1
    for f in (amg.smoothed_aggregation, amg.ruge_stuben):
        H = oracle.OracleHierarchy(f(a))
        x = H.solve(b)
        assert np.sum((a.matvec(x) - b) ** 2) < 1e-10
        x = H.pcg(b, maxiter=1000)
        assert np.sum((a.matvec(x) - b) ** 2) < 1e-10


def test_regression_95_nonsymmetric(amg, fx):
    # test/test_regression.jl:71-83
    N = 10000
    A = fx.sprand_plus_diag(N, 0.001, 5.0, seed=5)
    b = np.ones(N)
    for f in (amg.ruge_stuben, amg.smoothed_aggregation):
        x = oracle.OracleHierarchy(f(A, symmetry=amg.NoSymmetry())).solve(b)
        assert np.allclose(A.matvec(x), b, rtol=1e-8)


def test_elasticity_nns(amg, fx):
    # test/nns_test.jl:214-223
    A, b, B = fx.matrix("elastic"), fx.array("elastic_b"), fx.array("elastic_B")
    x, res = oracle.OracleHierarchy(amg.smoothed_aggregation(A, B=B)).solve(b, log=True, reltol=1e-10)
    assert _isapprox(A.matvec(x), b)
    x2, res2 = oracle.OracleHierarchy(amg.smoothed_aggregation(A, coarse_solver=amg.Pinv)).solve(b, log=True, reltol=1e-10)
    assert not _isapprox(A.matvec(x2), b) and res2[0] > res2[-1]


def test_nns_B_forms_agree(amg):
    # test/nns_test.jl:6-24
    A = amg.poisson(100)
    b = np.random.default_rng(4).random(100)
    xs = []
    for B in (None, np.ones(100), np.ones((100, 1))):
        H = oracle.OracleHierarchy(amg.smoothed_aggregation(A, B=B))
        xs.append(H.solve(b, maxiter=1, abstol=1e-6))
    assert np.allclose(xs[0], xs[1]) and np.allclose(xs[0], xs[2])


def test_loop_quirks(amg):
    # zero rhs => no iterations; calculate_residual=False => exactly maxiter cycles
    A = amg.poisson(100)
    H = oracle.OracleHierarchy(amg.ruge_stuben(A))
    x, res = H.solve(np.zeros(100), x0=np.ones(100), log=True)
    assert np.array_equal(x, np.ones(100)) and H.iters == 0 and list(res) == [0.0]
    H.solve(np.ones(100), maxiter=3, calculate_residual=False)
    assert H.iters == 3
    # ldiv! == zero + one cycle
    b = np.arange(100.0)
    assert np.array_equal(H.precond(b), H.cycle(np.zeros(100), b))


def test_user_assembled_geometric_hierarchy(amg):
    """test/gmg.jl + test/runtests.jl:104-108: a hierarchy a USER assembles from the package's own pieces — `Level(A, P, R,
    pre, post)` with an explicit linear-interpolation `P` (plain CSC), `R = adjoint(P)`, Galerkin `R*A*P`, `Pinv` on the
    coarsest level — has `length(ml) == 10` on poisson(10^6); the oracle then cycles on it like on any other hierarchy."""
    import oracle
    from algebraicmultigrid_jl_b200 import _hostlib
    from algebraicmultigrid_jl_b200.multilevel import coarse_b_, coarse_x_, residual_
    from algebraicmultigrid_jl_b200.smoother import setup_smoother
    from algebraicmultigrid_jl_b200.sparse import adjoint

    def extend(levels, A, pre, post):
        size_f = A.m
        size_c = (size_f - 1) // 2 + 1 if size_f % 2 == 0 else (size_f - 1) // 2      # gmg.jl:25
        k = np.arange(1, size_c + 1)
        rows = [2 * k - 1]                                                              # I = 2k (1-based)
        cols = [k - 1]
        vals = [np.ones(size_c)]
        k = np.arange(1, size_c)
        rows += [2 * k, 2 * k]                                                          # I = 2k+1 -> columns k and k+1
        cols += [k - 1, k]
        vals += [np.full(size_c - 1, 0.5), np.full(size_c - 1, 0.5)]
        import scipy.sparse as sp

        P = amg.SparseMatrixCSC.from_scipy(sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                                                         shape=(size_f, size_c)))
        R = adjoint(P)
        levels.append(amg.Level(A, P, R, setup_smoother(pre, A, amg.HermitianSymmetry()),
                                setup_smoother(post, A, amg.HermitianSymmetry())))
        Rm = P.transpose()
        return _hostlib.spgemm(_hostlib.spgemm(Rm, A), P)

    def multigrid(A, max_levels=10, max_coarse=10):
        pre, post = amg.GaussSeidel(), amg.GaussSeidel()
        levels = []
        w = amg.MultiLevelWorkspace(1, np.dtype(np.float64))
        while len(levels) + 1 < max_levels and A.m > max_coarse:
            residual_(w, A.m)
            A = extend(levels, A, pre, post)
            coarse_x_(w, A.m)
            coarse_b_(w, A.m)
        return amg.MultiLevel(levels, A, amg.Pinv(A), pre, post, w)

    ml = multigrid(amg.poisson(10 ** 6))
    assert len(ml) == 10                                                                # runtests.jl:108
    assert [lv.A.m for lv in ml.levels][:3] == [1000000, 500000, 250000]
    # the same construction on a size the coarsest level resolves: the cycle converges to the solution
    A = amg.poisson(1000)
    ml = multigrid(A)
    b = oracle.mul(A, np.ones(A.n))
    x, hist = oracle.OracleHierarchy(ml).solve(b, log=True, reltol=1e-10)
    assert hist[-1] <= 1e-10 * hist[0] and len(hist) < 40
    assert np.abs(x - 1).max() < 1e-6
