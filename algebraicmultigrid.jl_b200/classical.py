"""Ruge-Stuben setup — host-side mirror of ``/root/reference/src/classical.jl`` (setup phase, host)."""
from __future__ import annotations

import numpy as np

from . import _hostlib
from .coarse_solver import _default_coarse_solver
from .multilevel import Level, MultiLevel, MultiLevelWorkspace, coarse_b_, coarse_x_, residual_
from .smoother import GaussSeidel, setup_smoother
from .sparse import Adjoint
from .splitting import RS
from .strength import Classical
from .utils import HermitianSymmetry, Symmetric, get_symmetry_and_data


def ruge_stuben(A, bs=1, *, strength=None, symmetry=None, CF=None, presmoother=None, postsmoother=None,
                max_levels=10, max_coarse=10, coarse_solver=None, **kwargs):
    """``ruge_stuben(A; ...)`` (``classical.jl:1-34``)."""
    if isinstance(A, Symmetric):
        A, sym = get_symmetry_and_data(A)
        symmetry = sym
    strength = Classical(0.25) if strength is None else strength
    symmetry = HermitianSymmetry() if symmetry is None else symmetry
    CF = RS() if CF is None else CF
    presmoother = GaussSeidel() if presmoother is None else presmoother
    postsmoother = GaussSeidel() if postsmoother is None else postsmoother
    coarse_solver = _default_coarse_solver(A) if coarse_solver is None else coarse_solver
    if kwargs.get("B") is not None:
        raise RuntimeError("near null space `B` is only supported for smoothed aggregation AMG, not Ruge-Stüben AMG.")

    levels = []
    eltype = getattr(A, "eltype", np.dtype(np.float64))
    w = MultiLevelWorkspace(bs, eltype)
    residual_(w, A.m)
    while len(levels) + 1 < max_levels and A.m > max_coarse:
        A, stop = extend_hierarchy_rs_(levels, strength, CF, A, presmoother, postsmoother, symmetry)
        if stop:
            break
        coarse_x_(w, A.m)
        coarse_b_(w, A.m)
        residual_(w, A.m)
    _hostlib.spgemm_release()
    cs = coarse_solver(A)
    return MultiLevel(levels, A, cs, presmoother, postsmoother, w)


def _as_eltype(M, A):
    """A hierarchy built from a Float32 matrix holds Float32 operators on every level (the reference computes them in Float32;
    here they are computed in fp64 and rounded once): keeps ``eltype(ml)`` and lets the device store them with 4-byte values."""
    if getattr(A, "eltype", np.dtype(np.float64)) != np.float32:
        return M
    if isinstance(M, Adjoint):
        return Adjoint(M.parent.astype(np.float32))
    return M.astype(np.float32)


def extend_hierarchy_rs_(levels, strength, CF, A, presmoother, postsmoother, symmetry):
    """``extend_hierarchy_rs!`` (``classical.jl:36-55``)."""
    At = A if isinstance(symmetry, HermitianSymmetry) else A.transpose()
    S, T = strength(At)
    splitting = CF(S)
    P, R = direct_interpolation(At, T, splitting)
    if P.shape[1] == 0:
        return A, True
    if getattr(A, "eltype", np.dtype(np.float64)) == np.float32:
        R = _as_eltype(R, A)
        P = Adjoint(R) if isinstance(P, Adjoint) else _as_eltype(P, A)
    RAP = _hostlib.spgemm(_hostlib.spgemm(R, A), R.transpose())      # (R*A)*P, structural zeros kept
    pre = setup_smoother(presmoother, A, symmetry)
    post = setup_smoother(postsmoother, A, symmetry)
    levels.append(Level(A, P, R, pre, post))
    return _as_eltype(RAP, A), False


def direct_interpolation(At, T, splitting):
    """``direct_interpolation`` (``classical.jl:57-68``): returns ``(P, R)`` with ``P = R'`` lazy."""
    R = _hostlib.direct_interpolation(At, T, np.asarray(splitting))
    return Adjoint(R), R
