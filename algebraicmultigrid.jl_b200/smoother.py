"""Smoother protocol — host-side mirror of ``/root/reference/src/smoother.jl``.

The configuration objects (``GaussSeidel``, ``Jacobi``, ``SOR``, the sweeps) and the two protocol
functions ``setup_smoother(config, A, symmetry)`` / ``smooth_(x, s, b)`` (the reference's
``smooth!``) keep the reference's names, argument meaning and error behaviour
(``smoother.jl:1-49``).  The relaxation itself runs in the CUDA engine behind
``include/b200amg.h`` (``b200amg_smoother_*`` for standalone caches, ``b200amg_smooth`` for a
level of a hierarchy); there is no CPU implementation on this path.
"""
from __future__ import annotations

import numpy as np

from .sparse import SparseMatrixCSC
from .utils import HermitianSymmetry, NoSymmetry


class SingularException(ArithmeticError):
    """``LinearAlgebra.SingularException(col)`` (thrown by ``DiagonalIndices``, ``smoother.jl:239-241``)."""

    def __init__(self, col):
        super().__init__(f"SingularException({col})")
        self.info = col


class Sweep:
    name = "?"

    def __repr__(self):
        return type(self).__name__ + "()"


class SymmetricSweep(Sweep):
    name = "symmetric"


class ForwardSweep(Sweep):
    name = "forward"


class BackwardSweep(Sweep):
    name = "backward"


class Smoother:
    """Abstract smoother configuration.  Calling ``config(A, x, b[, symmetry])`` is the in-place
    convenience form of ``smoother.jl:34-38``: ``setup_smoother`` then ``smooth!``."""

    kind = "none"

    def __call__(self, A, x, b, symmetry=None):
        s = setup_smoother(self, A, HermitianSymmetry() if symmetry is None else symmetry)
        smooth_(x, s, b)
        return None


class GaussSeidel(Smoother):
    """``GaussSeidel(; iter=1)`` (symmetric sweep) / ``GaussSeidel(sweep; iter=1)`` / ``GaussSeidel(sweep, iter)``
    (``smoother.jl:18-23``)."""

    kind = "gs"

    def __init__(self, sweep=None, iter=1):
        if isinstance(sweep, int) and not isinstance(sweep, Sweep):   # GaussSeidel(iter) is not a reference form
            raise TypeError("GaussSeidel(sweep::Sweep, iter::Int)")
        self.sweep = SymmetricSweep() if sweep is None else sweep
        self.iter = int(iter)
        self.omega = 1.0

    @property
    def sweep_name(self):
        return self.sweep.name

    def __repr__(self):
        return f"GaussSeidel({self.sweep!r}, {self.iter})"


class Jacobi(Smoother):
    """``Jacobi(ω; iter=1)`` (``smoother.jl:92-99``; not exported by the reference either)."""

    kind = "jacobi"
    sweep_name = "symmetric"

    def __init__(self, omega=0.5, iter=1):
        self.omega = float(omega)
        self.iter = int(iter)

    def __repr__(self):
        return f"Jacobi({self.omega}, iter={self.iter})"


class SOR(Smoother):
    """``SOR(ω; iter=1)`` (symmetric) / ``SOR(ω, sweep)`` / ``SOR(ω, sweep, iter)`` (``smoother.jl:173-180``)."""

    kind = "sor"

    def __init__(self, omega, sweep=None, iter=1):
        self.omega = float(omega)
        self.sweep = SymmetricSweep() if sweep is None else sweep
        self.iter = int(iter)

    @property
    def sweep_name(self):
        return self.sweep.name

    def __repr__(self):
        return f"SOR({self.omega}, {self.sweep!r}, {self.iter})"


class SmootherCache:
    """What ``setup_smoother`` returns: holds A (as the reference's caches do) and, once it has been
    applied standalone, a device-side smoother object.  Inside a ``MultiLevel`` the hierarchy's own
    device handle owns the level's smoother state instead."""

    def __init__(self, config, A, symmetry):
        self.config = config
        self.A = A
        self.symmetry = symmetry
        self._dev = None

    @property
    def symmetry_name(self):
        return "none" if isinstance(self.symmetry, NoSymmetry) else "hermitian"

    @property
    def iter(self):
        return self.config.iter

    def _device(self):
        if self._dev is None:
            from . import _devlib

            self._dev = _devlib.DeviceSmoother(self.A, self.config, self.symmetry_name)
        return self._dev


def _check_diagonal(A: SparseMatrixCSC):
    """``DiagonalIndices(A)`` (``smoother.jl:233-248``): every column needs a stored non-zero diagonal."""
    cp, rv, nz = A.colptr, A.rowval, A.nzval
    cols = np.repeat(np.arange(A.n, dtype=np.int64), np.diff(cp))
    hit = (rv == cols) & (nz != 0)
    has = np.zeros(A.n, dtype=bool)
    has[cols[hit]] = True
    if not has.all():
        raise SingularException(int(np.argmin(has)) + 1)


def setup_smoother(config, A, symmetry):
    """``setup_smoother(config::Smoother, A, symmetry)`` (``smoother.jl:40-49`` + the per-type methods)."""
    if not isinstance(config, (GaussSeidel, Jacobi, SOR)) or not isinstance(symmetry, (HermitianSymmetry, NoSymmetry)):
        raise RuntimeError(
            "setup_smoother(config, matrix, symmetry) not dispatched for smoother type "
            f"{type(config).__name__} and symmetry type {type(symmetry).__name__}")
    if isinstance(symmetry, NoSymmetry) and isinstance(config, (GaussSeidel, SOR)):
        _check_diagonal(A)
    return SmootherCache(config, A, symmetry)


def smooth_(x, s: SmootherCache, b):
    """``smooth!(x, smoother, b)``: relaxation sweeps updating ``x`` in place on the device."""
    if np.ndim(x) != np.ndim(b) or np.shape(x)[1:] != np.shape(b)[1:]:
        raise AssertionError("x and b must have the same number of columns")   # smoother.jl:76
    s._device().apply(x, b)
    return None
