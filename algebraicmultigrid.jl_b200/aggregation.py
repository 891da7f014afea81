"""Smoothed-aggregation setup — host-side mirror of ``/root/reference/src/aggregation.jl`` (setup phase, host)."""
from __future__ import annotations

import numpy as np

from . import _hostlib
from .aggregate import StandardAggregation
from .coarse_solver import _default_coarse_solver
from .multilevel import Level, MultiLevel, MultiLevelWorkspace, coarse_b_, coarse_x_, residual_
from .smoother import GaussSeidel, setup_smoother
from .sparse import Adjoint, SparseMatrixCSC
from .strength import SymmetricStrength
from .utils import HermitianSymmetry, Symmetric, get_symmetry_and_data


class DiagonalWeighting:
    pass


class LocalWeighting:
    pass


def _diagonal_weight(A, omega):
    """``weight(::DiagonalWeighting, S, ω)`` (``aggregation.jl:19-24``): ``(ω / ρ(D⁻¹S)) D⁻¹S`` with D = diag(S) and ρ the
    approximate spectral radius (random start vector: the prolongator is not bit-reproducible, as in the reference)."""
    from .utils import approximate_spectral_radius

    n = A.n
    cols = np.repeat(np.arange(n), np.diff(A.colptr))
    diag = np.zeros(n)
    on_diag = A.rowval == cols
    diag[A.rowval[on_diag]] = A.nzval[on_diag]
    with np.errstate(divide="ignore"):
        d_inv = 1.0 / diag                                         # 1 ./ diag(S): Inf where the diagonal is missing, like Julia
    scaled = SparseMatrixCSC(A.m, A.n, A.colptr.copy(), A.rowval.copy(), A.nzval * d_inv[A.rowval])   # scale_rows
    rho = approximate_spectral_radius(scaled)
    scaled.nzval *= omega / rho
    return scaled


class JacobiProlongation:
    """``JacobiProlongation(ω)`` (``aggregation.jl:1-17``): ``P = T - (ω D^-1 A) T`` with LocalWeighting (the default),
    D = row sums of |A| (``:26-47``), or DiagonalWeighting, D = diag(A) scaled by the spectral radius (``:19-24``)."""

    def __init__(self, omega):
        self.omega = omega

    def __call__(self, A, T, S=None, B=None, degree=1, weighting=None):
        if isinstance(weighting, DiagonalWeighting):
            d_inv_s = _diagonal_weight(A, self.omega)
        elif weighting is None or isinstance(weighting, LocalWeighting):
            d_inv_s = _hostlib.local_weight(A, self.omega)
        else:
            raise TypeError(f"unknown weighting {weighting!r}")
        P = T
        for _ in range(degree):
            P = _hostlib.sub(P, _hostlib.spgemm(d_inv_s, P))
        return P


def smoothed_aggregation(A, bs=1, *, B=None, symmetry=None, strength=None, aggregate=None, smooth=None,
                         presmoother=None, postsmoother=None, improve_candidates=None, max_levels=10,
                         max_coarse=10, diagonal_dominance=False, keep=False, verbose=False,
                         coarse_solver=None, **kwargs):
    """``smoothed_aggregation(A; ...)`` (``aggregation.jl:61-114``)."""
    if isinstance(A, Symmetric):
        A, sym = get_symmetry_and_data(A)
        symmetry = sym
    symmetry = HermitianSymmetry() if symmetry is None else symmetry
    strength = SymmetricStrength() if strength is None else strength
    aggregate = StandardAggregation() if aggregate is None else aggregate
    smooth = JacobiProlongation(4.0 / 3.0) if smooth is None else smooth
    presmoother = GaussSeidel() if presmoother is None else presmoother
    postsmoother = GaussSeidel() if postsmoother is None else postsmoother
    improve_candidates = GaussSeidel(iter=4) if improve_candidates is None else improve_candidates
    coarse_solver = _default_coarse_solver(A) if coarse_solver is None else coarse_solver

    n = A.m
    B = np.ones(n) if B is None else np.array(B, dtype=np.float64, copy=True)
    assert A.m == B.shape[0]
    levels = []
    bsr_flag = False
    eltype = getattr(A, "eltype", np.dtype(np.float64))
    w = MultiLevelWorkspace(bs, eltype)
    residual_(w, A.m)
    while len(levels) + 1 < max_levels and A.m > max_coarse:
        A, B, bsr_flag, stop = extend_hierarchy_sa_(levels, strength, aggregate, smooth, improve_candidates,
                                                    diagonal_dominance, keep, A, B, presmoother, postsmoother,
                                                    symmetry, bsr_flag, verbose)
        if stop:
            break
        coarse_x_(w, A.m)
        coarse_b_(w, A.m)
        residual_(w, A.m)
    _hostlib.spgemm_release()
    cs = coarse_solver(A)
    ml = MultiLevel(levels, A, cs, presmoother, postsmoother, w)
    if verbose:
        print(ml)
    return ml


def extend_hierarchy_sa_(levels, strength, aggregate, smooth, improve_candidates, diagonal_dominance, keep, A, B,
                         presmoother, postsmoother, symmetry, bsr_flag, verbose):
    """``extend_hierarchy_sa!`` (``aggregation.jl:116-157``)."""
    if isinstance(symmetry, HermitianSymmetry):
        S, _T = strength(A, bsr_flag)
    else:
        S, _T = strength(A.transpose(), bsr_flag)
    AggOp = aggregate(S)
    if AggOp.m == 0:
        return A, B, bsr_flag, True
    b = np.zeros(B.shape)
    _improve_candidates(improve_candidates, A, B, b)
    T, B = fit_candidates(AggOp, B)
    P = smooth(A, T, S, B)
    if P.n == 0:
        return A, B, True, True
    f32 = getattr(A, "eltype", np.dtype(np.float64)) == np.float32     # a Float32 hierarchy holds Float32 operators on every level
    if f32:
        P = P.astype(np.float32)
    R = construct_R(symmetry, P)
    RAP = _hostlib.spgemm(_hostlib.spgemm(P.transpose(), A), P)      # (R*A)*P
    if f32:
        RAP = RAP.astype(np.float32)
    pre = setup_smoother(presmoother, A, symmetry)
    post = setup_smoother(postsmoother, A, symmetry)
    levels.append(Level(A, P, R, pre, post))
    return RAP, B, True, False


def _improve_candidates(config, A, B, b):
    """``improve_candidates(A, B, b)`` (``aggregation.jl:135-136``): a setup-time relaxation on A·B = 0,
    always the Hermitian-fast variant (``smoother.jl:34-38``).  Gauss-Seidel (the default) is done by the
    host setup library like the rest of setup; other smoothers go through the device smoother."""
    if config is None:
        return
    if config.kind == "gs":
        _hostlib.gs_sweeps(A, b, B, config.iter, forward=config.sweep_name in ("forward", "symmetric"),
                           backward=config.sweep_name in ("backward", "symmetric"))
    else:
        if B.ndim == 1:
            config(A, B, b)
        else:
            for c in range(B.shape[1]):
                col = np.ascontiguousarray(B[:, c])
                config(A, col, np.ascontiguousarray(b[:, c]))
                B[:, c] = col


def construct_R(symmetry, P):
    return Adjoint(P)                                                # aggregation.jl:158-159


def fit_candidates(AggOp, B, tol=1e-10):
    """``fit_candidates`` (``aggregation.jl:161-230``).  Returns ``(T, B_coarse)``."""
    A = AggOp.transpose()                                            # adjoint(AggOp) == copy(AggOp')
    if B.ndim == 1:
        tx, rc = _hostlib.fit_candidates_vec(A, B, tol)
        return SparseMatrixCSC(A.m, A.n, A.colptr, A.rowval, tx), rc
    n_fine, m = B.shape
    n_agg = A.n
    assert A.m == n_fine
    n_coarse = m * n_agg
    R = np.zeros((n_coarse, m))
    cols = [[] for _ in range(n_coarse)]
    for agg in range(n_agg):
        rows = A.rowval[A.colptr[agg]:A.colptr[agg + 1]]
        M = B[rows, :]
        q, rj = np.linalg.qr(M, mode="reduced")
        r = min(len(rows), m)
        offset = agg * m
        for lj in range(r):
            for li in range(len(rows)):
                val = q[li, lj]
                if abs(val) >= tol:
                    cols[offset + lj].append((int(rows[li]), float(val)))
        R[offset:offset + r, :] = rj[:r, :]
    colptr = [0]
    rowval, nzval = [], []
    for c in cols:
        c.sort()
        for r_, v in c:
            if v != 0.0:
                rowval.append(r_)
                nzval.append(v)
        colptr.append(len(rowval))
    return SparseMatrixCSC(n_fine, n_coarse, colptr, rowval, nzval), R
