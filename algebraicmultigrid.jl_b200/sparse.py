"""Minimal compressed-sparse-column container mirroring what Julia's ``SparseMatrixCSC`` holds.

The reference passes ``SparseMatrixCSC{Float64,Int64}`` objects around
(``/root/reference/src/multilevel.jl:1-8``); the host side here keeps the same three arrays,
0-based and int32 (the device index width) so an upload needs no conversion.  A Julia
caller hands over its own 1-based Int64 arrays through the C-ABI unchanged
(``include/b200amg.h``: ``index_bits`` / ``index_base``).
"""
from __future__ import annotations

import numpy as np

from . import _hostlib

IDX = np.int32


class SparseMatrixCSC:
    """m x n sparse matrix: ``colptr`` (n+1), ``rowval`` (nnz, sorted per column), ``nzval`` (nnz)."""

    __slots__ = ("m", "n", "colptr", "rowval", "nzval", "_bitsym", "eltype", "_transpose_of")

    def __init__(self, m, n, colptr, rowval, nzval, eltype=None):
        self.m = int(m)
        self.n = int(n)
        self.colptr = np.ascontiguousarray(colptr, dtype=IDX)
        self.rowval = np.ascontiguousarray(rowval, dtype=IDX)
        # ``eltype``: the element type the caller's matrix has (Float32 or Float64, test/runtests.jl:244-259).  The values are
        # always HELD as float64 (a Float32 matrix holds binary32-representable numbers; host setup and device kernels compute
        # in fp64, the device stores such operators with 4-byte values: B200AMG_OPT_FP32_STORAGE)
        if eltype is None:
            eltype = np.float32 if getattr(nzval, "dtype", None) == np.float32 else np.float64
        self.eltype = np.dtype(eltype)
        self.nzval = np.ascontiguousarray(nzval, dtype=np.float64)
        if self.colptr.shape[0] != self.n + 1:
            raise ValueError("colptr must have n+1 entries")
        nnz = int(self.colptr[-1]) if self.n >= 0 and self.colptr.size else 0
        if self.rowval.shape[0] < nnz or self.nzval.shape[0] < nnz:
            raise ValueError("rowval/nzval shorter than colptr[end]")
        self.rowval = self.rowval[:nnz]
        self.nzval = self.nzval[:nnz]
        self._bitsym = None
        self._transpose_of = None   # set by ``Classical``: the matrix this one is the (unmodified) transpose of

    # -- constructors -------------------------------------------------------------------
    @classmethod
    def from_julia(cls, m, n, colptr, rowval, nzval):
        """From 1-based arrays as written in the reference's ``test/*.jl`` fixtures."""
        return cls(m, n, np.asarray(colptr, dtype=np.int64) - 1, np.asarray(rowval, dtype=np.int64) - 1, nzval)

    @classmethod
    def from_scipy(cls, mat):
        mat = mat.tocsc()
        mat.sort_indices()
        return cls(mat.shape[0], mat.shape[1], mat.indptr, mat.indices, mat.data.astype(np.float64))

    @classmethod
    def from_dense(cls, a):
        a = np.asarray(a, dtype=np.float64)
        m, n = a.shape
        colptr = [0]
        rowval, nzval = [], []
        for j in range(n):
            nz = np.nonzero(a[:, j])[0]
            rowval.extend(nz.tolist())
            nzval.extend(a[nz, j].tolist())
            colptr.append(len(rowval))
        return cls(m, n, colptr, rowval, nzval)

    @classmethod
    def identity_pattern(cls, other, fill=1.0):
        return cls(other.m, other.n, other.colptr.copy(), other.rowval.copy(), np.full(other.nnz, fill))

    # -- basic properties ---------------------------------------------------------------
    @property
    def shape(self):
        return (self.m, self.n)

    @property
    def nnz(self):
        return int(self.colptr[-1])

    def size(self, d=None):
        return self.shape if d is None else self.shape[d - 1]

    def copy(self):
        return SparseMatrixCSC(self.m, self.n, self.colptr.copy(), self.rowval.copy(), self.nzval.copy(), self.eltype)

    def astype(self, eltype):
        """``T.(A)``: the same pattern with the values rounded to ``eltype`` (``np.float32`` / ``np.float64``)."""
        eltype = np.dtype(eltype)
        vals = self.nzval.astype(np.float32).astype(np.float64) if eltype == np.float32 else self.nzval.copy()
        return SparseMatrixCSC(self.m, self.n, self.colptr.copy(), self.rowval.copy(), vals, eltype)

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csc_matrix((self.nzval, self.rowval, self.colptr), shape=self.shape)

    def todense(self):
        out = np.zeros(self.shape)
        for j in range(self.n):
            sl = slice(self.colptr[j], self.colptr[j + 1])
            out[self.rowval[sl], j] = self.nzval[sl]
        return out

    def transpose(self):
        """``copy(A')`` for a real matrix (``/root/reference/src/utils.jl:21-23``)."""
        return _hostlib.transpose(self)

    def is_bitsymmetric(self):
        if self._bitsym is None:
            self._bitsym = bool(self.m == self.n and _hostlib.is_bitsymmetric(self))
        return self._bitsym

    def matvec(self, x):
        """Host helper for building right-hand sides (``b = A*ones``); not a solve-phase path."""
        return _hostlib.csc_matvec(self, np.ascontiguousarray(x, dtype=np.float64))

    def diag(self):
        d = np.zeros(min(self.m, self.n))
        for j in range(min(self.m, self.n)):
            sl = slice(self.colptr[j], self.colptr[j + 1])
            hit = np.nonzero(self.rowval[sl] == j)[0]
            if hit.size:
                d[j] = self.nzval[sl][hit].sum()
        return d

    def __add__(self, other):
        return SparseMatrixCSC.from_scipy(self.to_scipy() + (other.to_scipy() if isinstance(other, SparseMatrixCSC) else other))

    def __repr__(self):
        return f"SparseMatrixCSC({self.m}x{self.n}, nnz={self.nnz})"


class Adjoint:
    """Lazy adjoint wrapper, the analogue of ``LinearAlgebra.Adjoint`` that the reference uses for
    ``P = R'`` (``src/classical.jl:64-65``) and ``R = P'`` (``src/aggregation.jl:158-159``)."""

    __slots__ = ("parent",)

    def __init__(self, parent: SparseMatrixCSC):
        self.parent = parent

    @property
    def shape(self):
        return (self.parent.n, self.parent.m)

    @property
    def nnz(self):
        return self.parent.nnz

    def size(self, d=None):
        return self.shape if d is None else self.shape[d - 1]

    def materialize(self) -> SparseMatrixCSC:
        return self.parent.transpose()

    def __repr__(self):
        return f"Adjoint({self.parent!r})"


def adjoint(a):
    if isinstance(a, Adjoint):
        return a.parent
    return Adjoint(a)


def size(a, d=None):
    return a.size(d)


def nnz(a):
    return a.nnz
