"""``precs`` builders — host-side mirror of ``/root/reference/src/precs.jl``.

LinearSolve.jl does not exist here; what its ``precs`` API needs from this package is a callable ``(A, p) -> (Pl, Pr)``
that builds the left preconditioner from the matrix of the problem.  The builders below are those callables: the left
preconditioner is ``aspreconditioner(hierarchy)`` (its ``ldiv!`` is one cycle on the device), the right one the identity.
``preconditioner.cg(A, b, Pl=...)`` (the device-resident PCG) takes the result the way ``KrylovJL_CG(precs = ...)`` does in
the reference's test (``test/runtests.jl:227-240``).
"""
from __future__ import annotations

from .aggregation import smoothed_aggregation
from .classical import ruge_stuben
from .preconditioner import aspreconditioner


class _Identity:
    """``LinearAlgebra.I`` as a right preconditioner."""

    def __repr__(self):
        return "I"


I = _Identity()


class SmoothedAggregationPreconBuilder:
    """``SmoothedAggregationPreconBuilder(; blocksize = 1, kwargs...)`` (``precs.jl:1-17``)."""

    def __init__(self, blocksize=1, **kwargs):
        self.blocksize, self.kwargs = int(blocksize), kwargs

    def __call__(self, A, p=None):
        return aspreconditioner(smoothed_aggregation(A, self.blocksize, **self.kwargs)), I


class RugeStubenPreconBuilder:
    """``RugeStubenPreconBuilder(; blocksize = 1, kwargs...)`` (``precs.jl:20-36``)."""

    def __init__(self, blocksize=1, **kwargs):
        self.blocksize, self.kwargs = int(blocksize), kwargs

    def __call__(self, A, p=None):
        return aspreconditioner(ruge_stuben(A, self.blocksize, **self.kwargs)), I
