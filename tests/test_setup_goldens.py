"""Host setup library vs the reference's own setup goldens (what gets uploaded must be what the
reference would have built).  Sources: /root/reference/test/runtests.jl:19-110, sa_tests.jl,
test_regression.jl:7-12,41-57, nns_test.jl."""
import numpy as np
import pytest


def test_classical_strength_poisson5(amg):
    # runtests.jl:22-30
    S, T = amg.Classical(0.2)(amg.poisson(5))
    ref = np.array([[1.0, 0.5, 0, 0, 0], [0.5, 1, 0.5, 0, 0], [0, 0.5, 1, 0.5, 0], [0, 0, 0.5, 1, 0.5], [0, 0, 0, 0.5, 1]])
    assert np.array_equal(S.todense(), ref)


def test_classical_strength_graph(amg, fx):
    # runtests.jl:31-33
    S, T = amg.Classical(0.25)(fx.matrix("test"))
    ref = fx.matrix("ref_S_test")
    assert (S.todense() - ref.todense()).max() < 1e-10
    assert np.abs(S.todense() - ref.todense()).max() < 1e-10


def test_rs_splitting(amg, fx):
    # runtests.jl:38-49
    assert list(amg.RS()(amg.poisson(7))) == [0, 1, 0, 1, 0, 1, 0]
    S, T = amg.Classical(0.25)(fx.matrix("thing"))
    assert list(amg.RS()(S)) == [0, 0, 1, 0, 1, 1, 0, 1, 0, 0, 1, 0, 1, 1, 0, 0, 1, 0, 0, 0, 1, 0, 1, 0, 1, 0, 0, 1, 0, 0,
                                 1, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 0]
    assert list(amg.RS()(fx.matrix("ref_S_test"))) == list(fx.array("ref_split"))


def test_direct_interpolation(amg, fx):
    # runtests.jl:55-67
    A = amg.poisson(5)
    P, R = amg.direct_interpolation(A, A.copy(), [1, 0, 1, 0, 1])
    ref = np.array([[1.0, 0, 0], [0.5, 0.5, 0], [0, 1, 0], [0, 0.5, 0.5], [0, 0, 1]])
    assert np.array_equal(P.materialize().todense(), ref)
    ml = amg.ruge_stuben(fx.matrix("thing"))
    assert ml.levels[1].A.m == 19


def test_multilevel_shapes(amg, fx):
    # runtests.jl:76-102 ; README.md:32-45
    ml = amg.ruge_stuben(amg.poisson(1000))
    assert len(ml) == 8
    assert [l.A.m for l in ml.levels] == [1000, 500, 250, 125, 62, 31, 15]
    assert [l.A.nnz for l in ml.levels] == [2998, 1498, 748, 373, 184, 91, 43]
    assert ml.final_A.m == 7 and ml.final_A.nnz == 19
    assert abs(amg.operator_complexity(ml) - 1.9859906604402935) < 1e-12
    assert round(amg.grid_complexity(ml), 3) == 1.99

    ml = amg.ruge_stuben(fx.matrix("randlap"))
    assert len(ml) == 3
    assert [l.A.m for l in ml.levels] == [100, 17]
    assert [l.A.nnz for l in ml.levels] == [2066, 289]
    assert ml.final_A.m == 2 and ml.final_A.nnz == 4
    assert round(amg.operator_complexity(ml), 3) == 1.142
    assert round(amg.grid_complexity(ml), 3) == 1.190


def test_show(amg):
    s = repr(amg.ruge_stuben(amg.poisson(1000)))
    assert "No. of Levels: 8" in s and "    1         1000         2998 [50.35%]" in s


def test_no_level_hierarchies(amg):
    # test_regression.jl:41-57 (issue #31)
    for sz in (10, 5, 2):
        for f in (amg.ruge_stuben, amg.smoothed_aggregation):
            ml = f(amg.poisson(sz))
            assert not ml.levels and ml.final_A.shape == (sz, sz)
            assert amg.operator_complexity(ml) == 1 and amg.grid_complexity(ml) == 1


def test_rs_rejects_B(amg):
    with pytest.raises(RuntimeError):
        amg.ruge_stuben(amg.poisson(100), B=np.ones(100))


def test_sa_onetoall(amg, fx):
    # test_regression.jl:7-12, sa_tests.jl:391-396 (issue #24)
    ml = amg.smoothed_aggregation(fx.matrix("onetoall"))
    assert ml.levels[1].A.shape == (11, 11)
    assert ml.final_A.shape == (2, 2)


def _symmetric_soc(amg, A, theta):
    # sa_tests.jl:3-23, in numpy
    a = A.todense()
    D = np.abs(np.diag(a))
    S = np.zeros_like(a)
    n = a.shape[0]
    for i in range(n):
        for j in range(n):
            if i != j and a[i, j] != 0 and abs(a[i, j] ** 2) >= theta * theta * D[i] * D[j]:
                S[i, j] = a[i, j]
    S = np.abs(S + np.diag(D))
    for j in range(n):
        m = max(0.0, S[:, j].max())
        S[:, j] /= m
    return S


def test_symmetric_strength(amg):
    # sa_tests.jl:26-38 on the poisson cases
    for s in (2, 3, 5, 7, 10, 11, 19):
        A = amg.poisson(s)
        for theta in (0.0, 0.1, 0.5, 1.0, 10.0):
            S, _ = amg.SymmetricStrength(theta)(A)
            assert np.sum((S.todense() - _symmetric_soc(amg, A, theta)) ** 2) < 1e-6


def _stand_agg(C):
    # sa_tests.jl:64-135 restated with dense arrays (python, tiny cases only)
    n = C.shape[0]
    R = set(range(n))
    j = 0
    Cpts = []
    aggregates = -np.ones(n, dtype=int)
    nbrs = [set(np.nonzero(C[:, i])[0]) - {i} for i in range(n)]  # column i pattern
    for i in range(n):
        Ni = nbrs[i] | {i}
        if len(Ni - R) == 0 and len(Ni) > 1:
            Cpts.append(i)
            for x in Ni:
                aggregates[x] = j
            R -= Ni
            j += 1
    old_R = set(R)
    for i in range(n):
        if i not in R:
            continue
        for x in sorted(nbrs[i]):
            if x not in old_R:
                aggregates[i] = aggregates[x]
                R.discard(i)
                break
    for i in range(n):
        if i not in R:
            continue
        Ni = nbrs[i] | {i}
        if len(Ni) == 1:
            continue
        Cpts.append(i)
        for x in Ni:
            if x in R:
                aggregates[x] = j
        j += 1
    return aggregates, j


def test_standard_aggregation_corner_cases(amg):
    # sa_tests.jl:140-186
    import scipy.sparse as sp

    S_chain = amg.SparseMatrixCSC.from_scipy(sp.csc_matrix((np.ones(6), ([0, 1, 1, 2, 2, 3], [1, 0, 2, 1, 3, 2])), shape=(4, 4)))
    agg = amg.StandardAggregation()(S_chain)
    assert agg.m == 2 and np.all(agg.todense().sum(axis=0) == 1)
    S_iso = amg.SparseMatrixCSC.from_scipy(sp.identity(5, format="csc"))
    assert amg.StandardAggregation()(S_iso).nnz == 0
    S_empty = amg.SparseMatrixCSC(0, 0, [0], [], [])
    assert amg.StandardAggregation()(S_empty).shape == (0, 0)
    A_diag = amg.SparseMatrixCSC.from_scipy(2.0 * sp.identity(20, format="csc"))
    ml = amg.smoothed_aggregation(A_diag)
    assert len(ml) == 1 and ml.final_A.shape == (20, 20)
    A_iso = amg.SparseMatrixCSC.from_scipy(sp.diags([[-0.5] * 4, [1.0, 1.0, 100.0, 1.0, 1.0], [-0.5] * 4], [-1, 0, 1], format="csc"))
    S5, _ = amg.SymmetricStrength(0.25)(A_iso)
    agg5 = amg.StandardAggregation()(S5)
    assert agg5.m == 2 and agg5.colptr[3] - agg5.colptr[2] == 0


def test_standard_aggregation_poisson(amg):
    for s in (2, 3, 5, 7, 10, 11, 19):
        A = amg.poisson(s)
        for theta in (0.0, 0.02, 0.1, 1.0):
            C, _ = amg.SymmetricStrength(theta)(amg.SparseMatrixCSC.from_scipy(A.to_scipy() + A.to_scipy().T))
            agg = amg.StandardAggregation()(C)
            ref, nagg = _stand_agg(C.todense())
            dense = np.zeros((nagg, s))
            for i, a in enumerate(ref):
                if a >= 0:
                    dense[a, i] = 1
            assert agg.m == nagg
            assert np.sum((agg.todense() - dense) ** 2) < 1e-6


def test_fit_candidates_vector(amg):
    # sa_tests.jl:204-268
    cases = [
        (amg.SparseMatrixCSC.from_julia(2, 5, range(1, 7), [1, 1, 1, 2, 2], np.ones(5)), np.ones(5)),
        (amg.SparseMatrixCSC.from_julia(2, 5, range(1, 7), [2, 2, 1, 1, 1], np.ones(5)), np.ones(5)),
        (amg.SparseMatrixCSC.from_julia(3, 9, range(1, 11), [1, 1, 1, 2, 2, 2, 3, 3, 3], np.ones(9)), np.ones(9)),
        (amg.SparseMatrixCSC.from_julia(3, 9, range(1, 11), [3, 2, 1, 1, 2, 3, 2, 1, 3], np.ones(9)), np.arange(1.0, 10.0)),
        (amg.SparseMatrixCSC.from_julia(2, 5, [1, 2, 3, 3, 4, 5], [1, 1, 2, 2], np.ones(4)), np.array([1.0, 1, 5, 2, 3])),
        (amg.SparseMatrixCSC.from_julia(3, 9, [1, 2, 3, 3, 4, 5, 6, 6, 7, 8], [1, 1, 2, 2, 2, 3, 3], np.ones(7)), np.arange(1.0, 10.0)),
    ]
    for AggOp, B in cases:
        B = B.copy()
        B[np.diff(AggOp.colptr) == 0] = 0
        Q, Rc = amg.fit_candidates(AggOp, B)
        Qd = Q.todense()
        assert np.allclose(B, Qd @ Rc)
        assert np.allclose(Qd @ (Qd.T @ B), B)


def test_fit_candidates_matrix(amg):
    # nns_test.jl:28-107
    import scipy.sparse as sp

    def agg(rows, cols, m, n):
        return amg.SparseMatrixCSC.from_scipy(sp.csc_matrix((np.ones(len(rows)), (np.array(rows) - 1, np.array(cols) - 1)), shape=(m, n)))

    cases = [
        (agg([1, 2, 3, 4, 5], [1, 1, 1, 2, 2], 5, 2), np.ones((5, 1))),
        (agg([1, 2, 3, 4, 5], [2, 2, 1, 1, 1], 5, 2), np.ones((5, 1))),
        (agg(range(1, 10), np.repeat([1, 2, 3], 3), 9, 3), np.ones((9, 1))),
        (agg(range(1, 10), [3, 2, 1, 1, 2, 3, 2, 1, 3], 9, 3), np.arange(9.0).reshape(9, 1)),
        (agg([1, 2, 3, 4], [1, 1, 2, 2], 4, 2), np.c_[np.ones(4), np.arange(4.0)]),
        (agg(range(1, 10), np.repeat([1, 2, 3], 3), 9, 3), np.c_[np.ones(9), np.arange(9.0)]),
        (agg(range(1, 10), [1, 1, 2, 2, 3, 3, 4, 4, 4], 9, 4), np.c_[np.ones(9), np.arange(9.0)]),
        (agg([1, 2, 3, 4], [1, 1, 2, 2], 4, 2), np.c_[np.ones(4), 1e-20 * np.arange(4.0)]),
        (agg([1, 2, 3, 4], [1, 1, 2, 2], 4, 2), 1e-20 * np.c_[np.ones(4), np.arange(4.0)]),
        (agg([1, 2, 4, 5], [1, 1, 2, 2], 5, 2), np.c_[np.ones(5), np.arange(1.0, 6.0)]),
        (agg([1, 2, 4, 5], [1, 1, 2, 2], 5, 2), np.c_[np.ones(5), np.arange(1.0, 6.0), np.arange(5.0, 0.0, -1)]),
        (agg([2, 3, 4, 5, 6], [1, 1, 2, 2, 2], 7, 2), np.c_[np.ones(7), np.arange(1.0, 8.0)]),
    ]
    for AggT, fine in cases:
        fine = fine.copy()
        d = AggT.todense()
        for i in range(d.shape[0]):
            if not d[i].any():
                fine[i, :] = 0
        Q, R = amg.fit_candidates(AggT.transpose(), fine)
        Qd = Q.todense()
        assert np.allclose(fine, Qd @ R)
        assert np.allclose(fine, Qd @ (Qd.T @ fine))


def test_jacobi_prolongator(amg, fx):
    # sa_tests.jl:382-388 vs test/ref_R.jl
    A = amg.poisson(100)
    x = amg.JacobiProlongation(4 / 3)(A, amg.poisson(100), 1, 1)
    assert np.sum((x.todense() - fx.matrix("ref_R").todense()) ** 2) < 1e-6


def test_elastic_fit_candidates(amg, fx):
    # nns_test.jl:225-233
    A = fx.matrix("elastic")
    B = fx.array("elastic_B")
    AggOp = amg.StandardAggregation()(A)
    Q, R = amg.fit_candidates(AggOp, B)
    mask = np.diff(AggOp.colptr) == 0
    Bm = B.copy()
    Qd = Q.todense()
    assert np.allclose(Bm[~mask], (Qd @ R)[~mask])
    ml = amg.smoothed_aggregation(A, B=B)
    assert ml.levels[0].P.n == ml.levels[1].A.m if len(ml.levels) > 1 else True


def test_gallery(amg):
    # gallery.jl: sizes / nnz closed forms (SURVEY §8 header)
    A = amg.poisson((64, 64))
    assert A.n == 4096 and A.nnz == 5 * 4096 - 4 * 64
    A = amg.poisson((16, 16, 16))
    assert A.n == 4096 and A.nnz == 7 * 4096 - 6 * 256
    d = A.todense()
    assert np.array_equal(d, d.T) and np.all(np.diag(d) == 6)
    assert d[0, 1] == -1 and d[0, 16] == -1 and d[0, 256] == -1 and d[15, 16] == 0
    assert A.is_bitsymmetric()


def test_elasticity_3d_generator(amg):
    """The synthetic 3-D Q1 elasticity problem (near-null-space laid out like create_nns_frame,
    test/nns_test.jl:138-164): symmetric positive definite, the six rigid-body modes lie in the kernel of every row that
    does not touch the clamped face, and smoothed_aggregation(A; B=B) coarsens with six candidates per aggregate."""
    A, b, B = amg.elasticity_3d(6, 5, 4)
    S = A.to_scipy().toarray()
    assert A.n == 3 * 6 * 6 * 5 and B.shape == (A.n, 6) and b.shape == (A.n,)
    assert np.abs(S - S.T).max() <= 1e-15 and np.linalg.eigvalsh(S).min() > 0
    free_nodes = np.nonzero(np.arange(7 * 6 * 5) % 7 != 0)[0]
    away = np.repeat((free_nodes % 7) >= 2, 3)
    assert np.abs(S @ B)[away].max() <= 1e-13 and np.abs(S @ B)[~away].max() > 1e-3
    assert np.linalg.matrix_rank(B) == 6
    ml = amg.smoothed_aggregation(A, B=B)
    assert len(ml.levels) >= 1 and ml.levels[0].P.shape[1] % 6 == 0


def test_host_galerkin_product_and_splitting_shortcuts(amg, fx):
    """The reworked host setup against independent computations: the chunked Galerkin product equals scipy's product
    entry for entry (sorted rows, structural zeros kept, also with far more chunks than columns or threads), the
    out-of-place `remove_diag!` equals the in-place statement of the reference (splitting.jl:8-18), and the C/F splitting
    computed with S' taken from T's pattern equals the one computed with an explicit transpose."""
    import scipy.sparse as sp

    from algebraicmultigrid_jl_b200 import _hostlib

    rng = np.random.default_rng(5)
    for (m, k, n, da, db) in ((40, 30, 50, 0.2, 0.15), (7, 5, 3, 0.9, 0.9), (300, 300, 300, 0.02, 0.03), (5, 4, 0, 0.5, 0.5)):
        Asp = sp.random(m, k, da, format="csc", random_state=rng)
        Bsp = sp.random(k, n, db, format="csc", random_state=rng)
        A, B = amg.SparseMatrixCSC.from_scipy(Asp), amg.SparseMatrixCSC.from_scipy(Bsp)
        Cm = _hostlib.spgemm(A, B)
        ref = (Asp @ Bsp).tocsc()
        ref.sort_indices()
        assert Cm.shape == (m, n)
        for j in range(n):
            rows = Cm.rowval[Cm.colptr[j]:Cm.colptr[j + 1]]
            assert np.all(np.diff(rows) > 0)                                  # sorted, no duplicates
        assert np.abs(Cm.to_scipy() - ref).max() <= 1e-14 if n else True
        assert Cm.nnz >= ref.nnz                                              # structural zeros are kept, never dropped
    # exact cancellation is KEPT as a stored zero (stdlib spmatmul semantics the nnz goldens rely on)
    A = amg.SparseMatrixCSC.from_dense(np.array([[1.0, -1.0], [0.0, 2.0]]))
    B = amg.SparseMatrixCSC.from_dense(np.array([[1.0], [1.0]]))
    Cm = _hostlib.spgemm(A, B)
    assert Cm.nnz == 2 and list(Cm.nzval) == [0.0, 2.0]

    # remove_diag!: zero the diagonal, then dropzeros!
    S = amg.SparseMatrixCSC.from_dense(np.array([[4.0, 1.0, 0.0], [2.0, 5.0, 3.0], [0.0, 7.0, 6.0]]))
    S.nzval[1] = 0.0            # a stored zero off the diagonal goes as well
    _hostlib.remove_diag(S)
    assert list(S.colptr) == [0, 0, 2, 3] and list(S.rowval) == [0, 2, 1] and list(S.nzval) == [1.0, 7.0, 3.0]

    # S' from T's pattern == explicit transpose, on an irregular strength matrix (RS coarse level) and a nonsymmetric one
    ml = amg.ruge_stuben(amg.poisson((14, 14, 14)))
    for At in (ml.levels[1].A, ml.levels[2].A, fx.sprand_plus_diag(300, 0.04, 4.0, seed=3)):
        s1, t1 = amg.Classical(0.25)(At)
        fast = amg.RS()(s1)
        s2, _ = amg.Classical(0.25)(At)
        s2._transpose_of = None
        slow = amg.RS()(s2)
        assert np.array_equal(fast, slow)
        cp, rv = _hostlib.offdiag_pattern(t1)
        tt = s2.transpose()                      # s2 had its diagonal removed by RS
        assert np.array_equal(cp, tt.colptr) and np.array_equal(rv, tt.rowval)


def test_approximate_spectral_radius_and_diagonal_weighting(amg):
    """test/sa_tests.jl:270-312 (`test_approximate_spectral_radius`): the estimate equals max |eig| on small diagonal,
    random and symmetrised matrices; and the `DiagonalWeighting` Jacobi prolongation smoother built on it
    (aggregation.jl:19-24) scales D^-1 A by omega / rho."""
    rng = np.random.default_rng(0)
    cases = [np.array([[2.0, 0.0], [0.0, 1.0]]), np.array([[-2.0, 0.0], [0.0, 1.0]]),
             np.array([[100.0, 0.0, 0.0], [0.0, 101.0, 0.0], [0.0, 0.0, 99.0]])]
    cases += [rng.random((i, i)) for i in range(2, 6)]
    for M in cases + [M + M.T for M in cases]:
        expected = np.abs(np.linalg.eigvals(M)).max()
        got = amg.approximate_spectral_radius(M, rng=np.random.default_rng(1))
        assert np.isclose(got, expected, rtol=1e-8), (M.shape, got, expected)
    # a sparse operator through the package's own matvec; a larger one within the stopping tolerance (1 %)
    A = amg.poisson((12, 12))
    rho = amg.approximate_spectral_radius(A, rng=np.random.default_rng(2))
    exact = np.abs(np.linalg.eigvalsh(A.todense())).max()
    assert abs(rho - exact) <= 0.02 * exact

    # JacobiProlongation with DiagonalWeighting: P = T - (omega / rho(D^-1 A)) D^-1 A T
    S, _ = amg.SymmetricStrength()(A)
    T = amg.SparseMatrixCSC.from_dense(np.kron(np.eye(72), np.ones((2, 1))))
    P = amg.JacobiProlongation(4.0 / 3.0)(A, T, S, None, 1, amg.DiagonalWeighting())
    Ad = A.todense()
    DinvA = Ad / np.diag(Ad)[:, None]
    rho_exact = np.abs(np.linalg.eigvals(DinvA)).max()
    Td = T.todense()
    # P = T - c D^-1 A T with c = omega / rho: recover c from the result and compare with omega / rho(D^-1 A) (1 % estimate)
    Y = DinvA @ Td
    c = float(np.sum((Td - P.todense()) * Y) / np.sum(Y * Y))
    assert np.abs(P.todense() - (Td - c * Y)).max() <= 1e-12
    assert abs(c - (4.0 / 3.0) / rho_exact) <= 0.03 * c
    ml = amg.smoothed_aggregation(A, smooth=lambda A_, T_, S_, B_: amg.JacobiProlongation(4.0 / 3.0)(A_, T_, S_, B_, 1, amg.DiagonalWeighting()))
    assert len(ml) >= 2 and ml.levels[0].P.shape[0] == A.n
