"""Cycle engine front-end — host-side mirror of ``/root/reference/src/multilevel.jl``.

``Level`` / ``MultiLevel`` keep the reference's field names.  ``_solve`` / ``_solve_`` (the
reference's ``_solve!``) keep its signature, defaults and loop quirks (``multilevel.jl:152-198``),
but the loop body — cycle, residual, norm — runs on the device through ``b200amg_solve``
(``include/b200amg.h``).  The hierarchy is uploaded once, on first use, and the handle is cached on
the ``MultiLevel`` (the analogue of the reference's preallocated ``MultiLevelWorkspace``).
"""
from __future__ import annotations

import numpy as np

from .sparse import Adjoint, SparseMatrixCSC


class Level:
    """``Level{TA,TP,TR}`` (``multilevel.jl:1-8``)."""

    __slots__ = ("A", "P", "R", "presmoother", "postsmoother")

    def __init__(self, A, P, R, presmoother, postsmoother):
        self.A, self.P, self.R = A, P, R
        self.presmoother, self.postsmoother = presmoother, postsmoother

    def __repr__(self):
        return f"Level with R {self.R.shape} | A {self.A.shape} | P {self.P.shape}"


class MultiLevelWorkspace:
    """``MultiLevelWorkspace{TX,bs}`` (``multilevel.jl:23-59``).  The vectors themselves live on the
    device inside the handle; the host object only records element type, block size and lengths."""

    def __init__(self, bs=1, eltype=np.float64):
        self.bs = bs
        self.eltype = np.dtype(eltype)
        self.res_vecs, self.coarse_xs, self.coarse_bs = [], [], []


def residual_(w, n):
    w.res_vecs.append(n)


def coarse_x_(w, n):
    w.coarse_xs.append(n)


def coarse_b_(w, n):
    w.coarse_bs.append(n)


def blocksize(w):
    return w.bs


class MultiLevel:
    """``MultiLevel`` (``multilevel.jl:14-21``)."""

    def __init__(self, levels, final_A, coarse_solver, presmoother, postsmoother, workspace):
        self.levels = levels
        self.final_A = final_A
        self.coarse_solver = coarse_solver
        self.presmoother = presmoother      # deprecated in the reference too
        self.postsmoother = postsmoother
        self.workspace = workspace
        self._dev = None
        self._partition = None

    def __len__(self):
        return len(self.levels) + 1                                   # multilevel.jl:61

    def device(self):
        """Upload once, reuse: the device-resident hierarchy behind this object."""
        if self._dev is None:
            from . import _devlib

            self._dev = _devlib.DeviceHierarchy(self, partition=self._partition)
        return self._dev

    def partition(self, rank, world_size, nccl_unique_id, levels=1):
        """Make this process one rank of a hierarchy whose ``levels`` finest levels are split by rows over
        ``world_size`` GPUs (call before first use); the remaining levels live on rank 0."""
        if self._dev is not None:
            raise RuntimeError("partition() must be called before the hierarchy is uploaded")
        self._partition = (rank, world_size, nccl_unique_id, int(levels))

    def release(self):
        if self._dev is not None:
            self._dev.close()
            self._dev = None

    def __repr__(self):                                               # multilevel.jl:63-96
        total_nnz = self.final_A.nnz + sum(l.A.nnz for l in self.levels)
        lines = []
        for i, level in enumerate(self.levels, 1):
            lines.append("   %2d   %10d   %10d [%5.2f%%]" % (i, level.A.m, level.A.nnz, 100 * level.A.nnz / total_nnz))
        lines.append("   %2d   %10d   %10d [%5.2f%%]" % (len(self.levels) + 1, self.final_A.m, self.final_A.nnz,
                                                         100 * self.final_A.nnz / max(total_nnz, 1)))
        return ("Multilevel Solver\n-----------------\n"
                f"Operator Complexity: {round(operator_complexity(self), 3)}\n"
                f"Grid Complexity: {round(grid_complexity(self), 3)}\n"
                f"No. of Levels: {len(self)}\n"
                f"Coarse Solver: {self.coarse_solver!r}\n"
                "Level     Unknowns     NonZeros\n-----     --------     --------\n" + "\n".join(lines) + "\n")


def operator_complexity(ml):                                          # multilevel.jl:98-105
    if ml.levels:
        return (sum(l.A.nnz for l in ml.levels) + ml.final_A.nnz) / ml.levels[0].A.nnz
    return 1.0


def grid_complexity(ml):                                              # multilevel.jl:107-114
    if ml.levels:
        return (sum(l.A.m for l in ml.levels) + ml.final_A.m) / ml.levels[0].A.m
    return 1.0


class Cycle:
    code = -1

    def __repr__(self):
        return type(self).__name__ + "()"


class V(Cycle):
    code = 0


class W(Cycle):
    code = 1


class F(Cycle):
    code = 2


def _result_dtype(ml, b):
    return np.promote_types(ml.workspace.eltype, np.asarray(b).dtype)


def _solve(ml, b, cycle=None, **kwargs):
    """``_solve(ml, b[, cycle]; kwargs...)`` (``multilevel.jl:152-157``): zero initial guess of the
    promoted element type, then ``_solve_``."""
    x = np.zeros(np.shape(b), dtype=_result_dtype(ml, b))
    return _solve_(x, ml, b, cycle if cycle is not None else V(), **kwargs)


def _solve_(x, ml, b, cycle=None, *, maxiter=100, abstol=None, reltol=None, verbose=False, log=False,
            calculate_residual=True, **kwargs):
    """``_solve!(x, ml, b, cycle; maxiter, abstol, reltol, verbose, log, calculate_residual)``
    (``multilevel.jl:158-198``).  ``x`` is the initial guess and is updated in place."""
    cycle = V() if cycle is None else cycle
    b = np.asarray(b)
    real_eps = np.finfo(b.dtype if b.dtype.kind == "f" else np.float64).eps
    if abstol is None:
        abstol = 0.0
    if reltol is None:
        reltol = float(np.sqrt(real_eps))
    n = ml.final_A.m if len(ml) == 1 else ml.levels[0].A.m
    if b.ndim == 2:
        return _solve_block_(x, ml, b, cycle, n, int(maxiter), float(abstol), float(reltol), verbose, log, calculate_residual)
    if b.shape[0] != n or np.shape(x)[0] != n:
        raise ValueError(f"DimensionMismatch: A has {n} rows, b has {b.shape[0]}, x has {np.shape(x)[0]}")
    xd = np.ascontiguousarray(x, dtype=np.float64)
    residuals, iters = ml.device().solve(xd, np.ascontiguousarray(b, dtype=np.float64), cycle.code, int(maxiter),
                                         float(abstol), float(reltol), bool(calculate_residual))
    if verbose and calculate_residual:
        # the reference prints the residual of the PREVIOUS iteration (multilevel.jl:185-187)
        for itr in range(1, iters + 1):
            print("Norm of residual at iteration %6d is %.4e" % (itr, residuals[itr - 1]))
    if xd is not x:
        x[...] = xd
    if log:
        return x, np.asarray(residuals, dtype=np.promote_types(ml.workspace.eltype, b.dtype))
    return x


def _solve_block_(x, ml, b, cycle, n, maxiter, abstol, reltol, verbose, log, calculate_residual):
    """Matrix right-hand sides (the reference's block size > 1 workspaces, ``multilevel.jl:28-59``): the reference
    relaxes, restricts and prolongs column by column (``smoother.jl:77,118``; stdlib ``mul!`` loops over columns) and
    tests ONE norm over all columns (Frobenius, ``multilevel.jl:170,190``).  ``b200amg_solve_block`` does exactly that
    with every column resident on the device for the whole call: nothing but one sum of squares per column crosses PCIe
    per iteration."""
    if b.shape[0] != n or np.shape(x) != b.shape:
        raise ValueError(f"DimensionMismatch: A has {n} rows, b is {b.shape}, x is {np.shape(x)}")
    dev = ml.device()
    xd = np.asfortranarray(x, dtype=np.float64)
    bd = np.asfortranarray(b, dtype=np.float64)
    if xd is x or np.shares_memory(xd, x):
        xd = xd.copy(order="F")
    residuals, iters = dev.solve_block(xd, bd, cycle.code, maxiter, abstol, reltol, bool(calculate_residual))
    if verbose and calculate_residual:
        for itr in range(1, iters + 1):
            print("Norm of residual at iteration %6d is %.4e" % (itr, residuals[itr - 1]))
    x[...] = xd
    if log:
        return x, np.asarray(residuals)
    return x


# --- CommonSolve front-end (multilevel.jl:241-264) --------------------------------------------
class AMGSolver:
    def __init__(self, ml, b):
        self.ml, self.b = ml, b


class AMGAlg:
    pass


class RugeStubenAMG(AMGAlg):
    pass


class SmoothedAggregationAMG(AMGAlg):
    pass


def init(alg, A, b, *args, **kwargs):
    from .aggregation import smoothed_aggregation
    from .classical import ruge_stuben

    if isinstance(alg, RugeStubenAMG):
        return AMGSolver(ruge_stuben(A, **kwargs), b)
    return AMGSolver(smoothed_aggregation(A, **kwargs), b)


def solve_(solt, *args, **kwargs):
    return _solve(solt.ml, solt.b, *args, **kwargs)


def solve(A, b, alg, *args, **kwargs):
    """kwargs go to BOTH setup and solve, each swallowing what it does not know (multilevel.jl:252-264)."""
    return solve_(init(alg, A, b, *args, **kwargs), *args, **kwargs)
