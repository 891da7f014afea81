"""ctypes binding of ``libb200amg_setup.so`` (host-only hierarchy construction, C++).

See ``csrc/host/amg_setup.cpp`` for the reference file:line each routine follows.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBDIR = os.path.join(_HERE, "_lib")
_LIBPATH = os.path.join(_LIBDIR, "libb200amg_setup.so")

_lib = None

i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    src = os.path.join(_HERE, "csrc", "host", "amg_setup.cpp")
    if force or not os.path.exists(_LIBPATH) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIBPATH)
    ):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "host"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            build()
        _lib = C.CDLL(_LIBPATH)
        _declare(_lib)
    return _lib


def _declare(L):
    i64 = C.c_int64
    L.amgsetup_poisson_nnz.restype = i64
    L.amgsetup_poisson_nnz.argtypes = [C.c_int, i64p]
    L.amgsetup_poisson.restype = C.c_int
    L.amgsetup_poisson.argtypes = [C.c_int, i64p, i32p, i32p, f64p]
    L.amgsetup_transpose.restype = C.c_int
    L.amgsetup_transpose.argtypes = [i64, i64, i32p, i32p, C.c_void_p, i32p, i32p, C.c_void_p]
    L.amgsetup_is_bitsymmetric.restype = C.c_int
    L.amgsetup_is_bitsymmetric.argtypes = [i64, i32p, i32p, f64p]
    L.amgsetup_classical_strength.restype = i64
    L.amgsetup_classical_strength.argtypes = [i64, i32p, i32p, f64p, C.c_double, i32p, i32p, f64p]
    L.amgsetup_symmetric_strength.restype = i64
    L.amgsetup_symmetric_strength.argtypes = [i64, i32p, i32p, f64p, C.c_double, i32p, i32p, f64p]
    L.amgsetup_remove_diag.restype = i64
    L.amgsetup_remove_diag.argtypes = [i64, i32p, i32p, f64p]
    L.amgsetup_remove_diag_copy.restype = i64
    L.amgsetup_remove_diag_copy.argtypes = [i64, i32p, i32p, C.c_void_p, i32p, i32p, C.c_void_p]
    L.amgsetup_rs_cf_splitting.restype = C.c_int
    L.amgsetup_rs_cf_splitting.argtypes = [i64, i32p, i32p, i32p, i32p, i32p]
    L.amgsetup_direct_interpolation.restype = i64
    L.amgsetup_direct_interpolation.argtypes = [i64, i32p, i32p, f64p, i32p, i32p, i32p, i32p, C.c_void_p, C.c_void_p, C.POINTER(i64)]
    L.amgsetup_spgemm_begin.restype = i64
    L.amgsetup_spgemm_begin.argtypes = [i64, i64, i64, i32p, i32p, f64p, i32p, i32p, f64p]
    L.amgsetup_spgemm_release.restype = None
    L.amgsetup_spgemm_release.argtypes = []
    L.amgsetup_spgemm_fetch.restype = C.c_int
    L.amgsetup_spgemm_fetch.argtypes = [i32p, i32p, f64p]
    L.amgsetup_standard_aggregation.restype = i64
    L.amgsetup_standard_aggregation.argtypes = [i64, i32p, i32p, f64p, i64p]
    L.amgsetup_fit_candidates_vec.restype = C.c_int
    L.amgsetup_fit_candidates_vec.argtypes = [i64, i32p, i32p, f64p, C.c_double, f64p, f64p]
    L.amgsetup_local_weight.restype = C.c_int
    L.amgsetup_local_weight.argtypes = [i64, i32p, i32p, f64p, C.c_double, f64p]
    L.amgsetup_sub.restype = i64
    L.amgsetup_sub.argtypes = [i64, i32p, i32p, f64p, i32p, i32p, f64p, i32p, i32p, f64p]
    L.amgsetup_gs_sweeps.restype = C.c_int
    L.amgsetup_gs_sweeps.argtypes = [i64, i32p, i32p, f64p, f64p, f64p, i64, C.c_int, C.c_int, C.c_int]
    L.amgsetup_residual_allcores.restype = C.c_double
    L.amgsetup_residual_allcores.argtypes = [i64, i32p, i32p, f64p, f64p, f64p, f64p, C.c_int]
    L.amgsetup_csc_matvec.restype = C.c_int
    L.amgsetup_csc_matvec.argtypes = [i64, i64, i32p, i32p, f64p, f64p, f64p]


def _csc(m, n, colptr, rowval, nzval):
    from .sparse import SparseMatrixCSC

    return SparseMatrixCSC(m, n, colptr, rowval, nzval)


def poisson(dims):
    dims = np.ascontiguousarray(dims, dtype=np.int64)
    L = lib()
    nnz = L.amgsetup_poisson_nnz(len(dims), dims)
    if nnz >= 2**31:
        raise OverflowError("nnz does not fit the int32 device index width")
    n = int(np.prod(dims))
    colptr = np.empty(n + 1, np.int32)
    rowval = np.empty(nnz, np.int32)
    nzval = np.empty(nnz, np.float64)
    rc = L.amgsetup_poisson(len(dims), dims, colptr, rowval, nzval)
    if rc:
        raise RuntimeError(f"amgsetup_poisson failed ({rc})")
    return _csc(n, n, colptr, rowval, nzval)


def transpose(a):
    nnz = a.nnz
    tp = np.empty(a.m + 1, np.int32)
    tr = np.empty(nnz, np.int32)
    tv = np.empty(nnz, np.float64)
    lib().amgsetup_transpose(a.m, a.n, a.colptr, a.rowval, a.nzval.ctypes.data, tp, tr, tv.ctypes.data)
    return _csc(a.n, a.m, tp, tr, tv)


def is_bitsymmetric(a):
    return lib().amgsetup_is_bitsymmetric(a.n, a.colptr, a.rowval, a.nzval)


def classical_strength(at, theta):
    tp = np.empty(at.n + 1, np.int32)
    tr = np.empty(at.nnz, np.int32)
    tv = np.empty(at.nnz, np.float64)
    nnz = lib().amgsetup_classical_strength(at.n, at.colptr, at.rowval, at.nzval, float(theta), tp, tr, tv)
    return _csc(at.m, at.n, tp, tr[:nnz], tv[:nnz])   # views: no copy (the untouched tail is never resident)


def symmetric_strength(a, theta):
    sp_ = np.empty(a.n + 1, np.int32)
    sr = np.empty(a.nnz, np.int32)
    sv = np.empty(a.nnz, np.float64)
    nnz = lib().amgsetup_symmetric_strength(a.n, a.colptr, a.rowval, a.nzval, float(theta), sp_, sr, sv)
    return _csc(a.m, a.n, sp_, sr[:nnz], sv[:nnz])


def remove_diag(s):
    """``remove_diag!`` (``splitting.jl:8-18``): the caller's S is mutated like in the reference (its arrays are replaced by
    the filtered ones, built on all cores)."""
    cp = np.empty(s.n + 1, np.int32)
    rv = np.empty(s.nnz, np.int32)
    nz = np.empty(s.nnz, np.float64)
    nnz = lib().amgsetup_remove_diag_copy(s.n, s.colptr, s.rowval, s.nzval.ctypes.data, cp, rv, nz.ctypes.data)
    s.colptr = cp
    s.rowval = rv[:nnz]
    s.nzval = nz[:nnz]
    s._bitsym = None
    return s


def offdiag_pattern(t):
    """``(colptr, rowval)`` of ``t`` without its diagonal and without stored zeros: the pattern of
    ``transpose(remove_diag!(transpose(t)))`` without forming either transpose."""
    cp = np.empty(t.n + 1, np.int32)
    rv = np.empty(t.nnz, np.int32)
    nnz = lib().amgsetup_remove_diag_copy(t.n, t.colptr, t.rowval, t.nzval.ctypes.data, cp, rv, None)
    return cp, rv[:nnz]


def rs_cf_splitting(s, t):
    """``t``: the transpose of ``s`` as a matrix or as a ``(colptr, rowval)`` pattern (only patterns are read)."""
    out = np.empty(s.n, np.int32)
    tcp, trv = (t.colptr, t.rowval) if hasattr(t, "colptr") else t
    rc = lib().amgsetup_rs_cf_splitting(s.n, s.colptr, s.rowval, tcp, trv, out)
    if rc:
        raise RuntimeError(f"rs_cf_splitting failed ({rc})")
    return out


def direct_interpolation(at, t, splitting):
    splitting = np.ascontiguousarray(splitting, dtype=np.int32)
    n = at.n
    rp = np.empty(n + 1, np.int32)
    nc = C.c_int64(0)
    L = lib()
    nnz = L.amgsetup_direct_interpolation(n, at.colptr, at.rowval, at.nzval, t.colptr, t.rowval, splitting, rp, None, None, C.byref(nc))
    rj = np.empty(nnz, np.int32)
    rx = np.empty(nnz, np.float64)
    rc = L.amgsetup_direct_interpolation(n, at.colptr, at.rowval, at.nzval, t.colptr, t.rowval, splitting, rp, rj.ctypes.data, rx.ctypes.data, C.byref(nc))
    if rc < 0:
        raise RuntimeError(f"direct_interpolation failed ({rc})")
    return _csc(int(nc.value), n, rp, rj, rx)


_SPGEMM_BACKEND = os.environ.get("B200AMG_GALERKIN", "host")


def set_galerkin_backend(name):
    """Where the Galerkin products ``R*A`` and ``(R*A)*P`` of the setup phase run: ``"host"`` (OpenMP, the default) or
    ``"device"`` (``b200amg_spgemm_*``, csrc/device/spgemm.cuh; bit-identical result, needs a GPU)."""
    global _SPGEMM_BACKEND
    if name not in ("host", "device"):
        raise ValueError("galerkin backend must be 'host' or 'device'")
    _SPGEMM_BACKEND = name


def spgemm(a, b):
    """``a * b`` keeping structural zeros (Julia ``SparseArrays`` semantics)."""
    if _SPGEMM_BACKEND == "device":
        from . import _devlib

        return _devlib.spgemm(a, b)
    if a.n != b.m:
        raise ValueError(f"DimensionMismatch: {a.shape} * {b.shape}")
    L = lib()
    nnz = L.amgsetup_spgemm_begin(a.m, a.n, b.n, a.colptr, a.rowval, a.nzval, b.colptr, b.rowval, b.nzval)
    if nnz < 0:
        raise OverflowError("spgemm result does not fit int32 indices")
    cp = np.empty(b.n + 1, np.int32)
    cj = np.empty(nnz, np.int32)
    cx = np.empty(nnz, np.float64)
    L.amgsetup_spgemm_fetch(cp, cj, cx)
    return _csc(a.m, b.n, cp, cj, cx)


def residual_allcores(a, x, b, reps=5):
    """``(r, seconds)``: ``r = b - a x`` on all host cores for a numerically symmetric ``a`` (columns walked as rows) —
    a courtesy figure for bench.py, not reference behaviour (the reference's solve phase is single-threaded)."""
    r = np.empty(a.n, np.float64)
    sec = lib().amgsetup_residual_allcores(a.n, a.colptr, a.rowval, a.nzval, np.ascontiguousarray(x, dtype=np.float64),
                                            np.ascontiguousarray(b, dtype=np.float64), r, int(reps))
    return r, float(sec)


def spgemm_release():
    """Free what the Galerkin products keep between calls (end of a setup): the per-thread accumulators of the host product
    and, when the device backend was used, its hash-table scratch on the GPU (up to a few GB)."""
    lib().amgsetup_spgemm_release()
    if _SPGEMM_BACKEND == "device":
        from . import _devlib

        _devlib.lib().b200amg_spgemm_release()


def standard_aggregation(s):
    x = np.empty(s.n, np.int64)
    nagg = lib().amgsetup_standard_aggregation(s.n, s.colptr, s.rowval, s.nzval, x)
    return x, int(nagg)


def fit_candidates_vec(a, b, tol):
    tx = np.empty(a.nnz, np.float64)
    rc = np.empty(a.n, np.float64)
    lib().amgsetup_fit_candidates_vec(a.n, a.colptr, a.rowval, np.ascontiguousarray(b, dtype=np.float64), float(tol), tx, rc)
    return tx, rc


def local_weight(a, omega):
    wx = np.empty(a.nnz, np.float64)
    lib().amgsetup_local_weight(a.n, a.colptr, a.rowval, a.nzval, float(omega), wx)
    return _csc(a.m, a.n, a.colptr.copy(), a.rowval.copy(), wx)


def sub(a, b):
    if a.shape != b.shape:
        raise ValueError("DimensionMismatch")
    cp = np.empty(a.n + 1, np.int32)
    cj = np.empty(a.nnz + b.nnz, np.int32)
    cx = np.empty(a.nnz + b.nnz, np.float64)
    nnz = lib().amgsetup_sub(a.n, a.colptr, a.rowval, a.nzval, b.colptr, b.rowval, b.nzval, cp, cj, cx)
    return _csc(a.m, a.n, cp, cj[:nnz].copy(), cx[:nnz].copy())


def gs_sweeps(a, b, x, iters, forward=True, backward=True):
    """Setup-time relaxation (improve_candidates).  ``x``/``b`` column-major n x ncols."""
    ncols = 1 if x.ndim == 1 else x.shape[1]
    xf = np.asfortranarray(x, dtype=np.float64)
    bf = np.asfortranarray(b, dtype=np.float64)
    xv = xf.reshape(-1, order="F").copy()
    bv = bf.reshape(-1, order="F").copy()
    lib().amgsetup_gs_sweeps(a.n, a.colptr, a.rowval, a.nzval, bv, xv, ncols, int(iters), int(forward), int(backward))
    out = xv.reshape(x.shape, order="F")
    x[...] = out
    return x


def csc_matvec(a, x):
    y = np.empty(a.m, np.float64)
    lib().amgsetup_csc_matvec(a.m, a.n, a.colptr, a.rowval, a.nzval, x, y)
    return y
