"""Preconditioner wrapper — host-side mirror of ``/root/reference/src/preconditioner.jl``."""
from __future__ import annotations

import numpy as np

from .multilevel import V


class Preconditioner:
    """``Preconditioner{ML,C}`` (``preconditioner.jl:1-6``)."""

    def __init__(self, ml, cycle, init="zero"):
        self.ml, self.cycle, self.init = ml, cycle, init


def aspreconditioner(ml, cycle=None):
    return Preconditioner(ml, V() if cycle is None else cycle)


def ldiv_(x, p, b=None):
    """``ldiv!(x, p, b)``: ``x .= 0`` (or ``x .= b``) then exactly one cycle, no residual
    (``preconditioner.jl:12-19``).  ``ldiv!(p, b)`` overwrites ``b`` (``:11``)."""
    if b is None:                      # ldiv!(p, b) with (x, p) = (p, b)
        p, b = x, p
        out = np.empty_like(b)
        ldiv_(out, p, b)
        b[...] = out
        return b
    xd = np.ascontiguousarray(x, dtype=np.float64)
    p.ml.device().precond(xd, np.ascontiguousarray(b, dtype=np.float64), p.cycle.code, p.init == "zero")
    if xd is not x:
        x[...] = xd
    return x


def mul_(out, p, x):
    """``mul!(b, p, x) = A_1 * x`` (``preconditioner.jl:20``)."""
    od = np.ascontiguousarray(out, dtype=np.float64)
    p.ml.device().apply(0, 0, od, np.ascontiguousarray(x, dtype=np.float64))
    if od is not out:
        out[...] = od
    return out


def backslash(p, b):
    """``p \\ b`` (``preconditioner.jl:22-24``)."""
    return ldiv_(np.empty_like(np.asarray(b, dtype=np.float64)), p, b)


def cg(A, b, Pl=None, *, abstol=0.0, reltol=None, maxiter=None, log=False):
    """Device-resident preconditioned CG with the call shape of ``IterativeSolvers.cg(A, b; Pl=p, ...)``
    as the reference's tests use it (``test/cycle_tests.jl:25``, ``test/runtests.jl:186,204``).  ``A`` must be
    the fine-level matrix of ``Pl.ml``; the whole iteration runs in ``b200amg_pcg``."""
    if Pl is None:
        raise NotImplementedError("cg without an AMG preconditioner is outside this package's path")
    b = np.ascontiguousarray(b, dtype=np.float64)
    if reltol is None:
        reltol = float(np.sqrt(np.finfo(np.float64).eps))
    if maxiter is None:
        maxiter = b.shape[0]
    x = np.zeros_like(b)
    res, iters = Pl.ml.device().pcg(x, b, Pl.cycle.code, int(maxiter), float(abstol), float(reltol))
    if log:
        return x, {"iters": iters, "resnorm": res}
    return x
