"""Strength of connection (``/root/reference/src/strength.jl``) — setup phase, host."""
from . import _hostlib
from .sparse import SparseMatrixCSC


class Classical:
    """``Classical(θ=0.25)``; calling it returns ``(S, T)`` with ``S = copy(T')`` (``strength.jl:7-37``)."""

    def __init__(self, theta=0.25):
        self.theta = theta

    def __call__(self, at: SparseMatrixCSC):
        t = _hostlib.classical_strength(at, self.theta)
        s = t.transpose()
        s._transpose_of = t   # lets ``RS`` take S' from T's pattern instead of transposing S back
        return s, t


class SymmetricStrength:
    """``SymmetricStrength(θ=0)`` (``strength.jl:72-122``); returns ``(S, S)``."""

    def __init__(self, theta=0.0):
        self.theta = theta

    def __call__(self, a: SparseMatrixCSC, bsr_flag=False):
        if bsr_flag and self.theta == 0:
            s = SparseMatrixCSC.identity_pattern(a, 1.0)   # strength.jl:81-84
            return s, s
        s = _hostlib.symmetric_strength(a, self.theta)
        return s, s
