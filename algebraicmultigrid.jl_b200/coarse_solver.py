"""Coarse solvers — host-side mirror of ``/root/reference/src/coarse_solver.jl``.

Factorising the coarsest matrix is setup work and stays on the host; what the cycle needs is the
APPLY ``cs(x, b)`` (``multilevel.jl:180,228``), which the device performs as a dense GEMV with the
operator returned by ``dense_operator()``.
"""
from __future__ import annotations

import numpy as np


DENSE_LIMIT = 16384   # rows: up to here the device applies the coarse solve as ONE dense operator (b200amg_set_coarse)


def _check_dense_limit(A, who):
    """``Pinv`` needs the dense n x n pseudo-inverse (``coarse_solver.jl:11``): a coarsest level beyond DENSE_LIMIT rows is
    refused with a clear message instead of exhausting memory (use ``QRSolver`` / ``LinearSolveWrapper`` there)."""
    if A.m > DENSE_LIMIT:
        raise ValueError(f"{who}: the coarsest matrix has {A.m} rows; the dense coarse operator is limited to {DENSE_LIMIT} "
                         "(raise max_levels / lower max_coarse so that coarsening continues, or use QRSolver / "
                         "LinearSolveWrapper, which keep a sparse factorisation)")


class CoarseSolver:
    """``cs(x, b)`` (``multilevel.jl:180,228``).  Two ways to reach the device: ``dense_operator()`` returns the matrix whose
    product is the solve (applied by a CUDA kernel), or ``None`` — then the object itself is called on the HOST from inside the
    cycle (``b200amg_set_coarse_callback``: the reference's coarse solver is any callable, ``coarse_solver.jl:24-58``)."""

    def dense_operator(self):
        raise NotImplementedError

    def __call__(self, x, b):
        """In-place host apply, ``x .= cs(b)`` (one column or an n x m block)."""
        op = self.dense_operator()
        x[...] = op @ b
        return x


class _SparseLU:
    """Sparse LU of a coarsest matrix too large for a dense operator (the reference keeps a sparse factorisation for ANY size,
    ``coarse_solver.jl:35-42,66-81``; SPQR does not exist here, a nonsingular matrix gives the same solution)."""

    def __init__(self, A):
        import scipy.sparse.linalg as spl

        # fill-reducing ordering: minimum degree on A + A' for structurally symmetric matrices (what Galerkin coarse operators
        # are) — on a 3-D Laplacian of 64 000 rows half the fill and a third of the factorisation time of SuperLU's default
        # COLAMD (12 s against 33 s)
        sym = A.m == A.n and A.is_bitsymmetric()
        self.lu = spl.splu(A.to_scipy().tocsc(), permc_spec="MMD_AT_PLUS_A" if sym else "COLAMD")

    def solve_into(self, x, b):
        x[...] = self.lu.solve(np.ascontiguousarray(b, dtype=np.float64))
        return x


class Pinv(CoarseSolver):
    """Moore-Penrose pseudo-inverse (``coarse_solver.jl:9-16``)."""

    def __init__(self, A):
        _check_dense_limit(A, "Pinv")
        self.pinvA = np.linalg.pinv(A.todense()) if A.n else np.zeros((0, 0))

    def dense_operator(self):
        return self.pinvA

    def __repr__(self):
        return "Pinv"


class QRSolver(CoarseSolver):
    """``qr(A)`` then ``F \\ b`` per column (``coarse_solver.jl:66-81``) — the reference's default
    (``:84``).  For a nonsingular coarse matrix ``F \\ b == inv(A) b``; the host forms that operator
    from a dense Householder QR.  For a numerically singular one (SPQR would return a basic
    solution) the minimum-norm operator ``pinv(A)`` is used instead — documented deviation."""

    def __init__(self, A):
        self.sparse = None
        if A.m > DENSE_LIMIT:   # sparse factorisation, applied on the host from inside the cycle
            self.op = None
            self.sparse = _SparseLU(A)
            return
        a = A.todense()
        n = a.shape[0]
        if n == 0:
            self.op = np.zeros((0, 0))
            return
        q, r = np.linalg.qr(a)
        d = np.abs(np.diag(r))
        if d.min() <= max(a.shape) * np.finfo(float).eps * d.max():
            self.op = np.linalg.pinv(a)
        else:
            import scipy.linalg as sl

            self.op = sl.solve_triangular(r, q.T)

    def dense_operator(self):
        return self.op

    def __call__(self, x, b):
        return self.sparse.solve_into(x, b) if self.sparse is not None else super().__call__(x, b)

    def __repr__(self):
        return "QRSolver"


class LinearSolveWrapperInternal(CoarseSolver):
    def __init__(self, A, alg):
        import scipy.sparse.linalg as spl

        self.alg = alg
        self.sparse = None
        if A.m > DENSE_LIMIT:
            self.op = None
            self.sparse = _SparseLU(A)
            return
        lu = spl.splu(A.to_scipy().tocsc())
        self.op = lu.solve(np.eye(A.n)) if A.n else np.zeros((0, 0))

    def dense_operator(self):
        return self.op

    def __call__(self, x, b):
        return self.sparse.solve_into(x, b) if self.sparse is not None else super().__call__(x, b)

    def __repr__(self):
        return str(self.alg)


class LinearSolveWrapper:
    """``LinearSolveWrapper(alg)`` (``coarse_solver.jl:50-58``).  LinearSolve.jl does not exist here;
    ``alg`` is a label (e.g. ``"UMFPACKFactorization"``) and the host uses a sparse LU."""

    def __init__(self, alg="UMFPACKFactorization"):
        self.alg = alg

    def __call__(self, A):
        return LinearSolveWrapperInternal(A, self.alg)


def UMFPACKFactorization():
    return "UMFPACKFactorization"


def _default_coarse_solver(A):
    return QRSolver
