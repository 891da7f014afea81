"""ctypes wrapper around ``liboracle.so`` (``amg_oracle.c``).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "liboracle.so")
_lib = None

i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

KIND = {"none": 0, "gs": 1, "jacobi": 2, "sor": 3}
SWEEP = {"forward": 1, "backward": 2, "symmetric": 3}
CYCLE = {"V": 0, "W": 1, "F": 2}


def build(force=False):
    src = os.path.join(_HERE, "amg_oracle.c")
    if force or not os.path.exists(_LIBPATH) or os.path.getmtime(src) > os.path.getmtime(_LIBPATH):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            build()
        L = C.CDLL(_LIBPATH)
        i64, dbl, ci = C.c_int64, C.c_double, C.c_int
        L.oracle_mul.restype = None
        L.oracle_mul.argtypes = [i64, i64, i32p, i32p, f64p, ci, f64p, f64p]
        L.oracle_norm.restype = dbl
        L.oracle_norm.argtypes = [i64, f64p]
        L.oracle_smooth.restype = ci
        L.oracle_smooth.argtypes = [i64, i32p, i32p, f64p, ci, ci, ci, dbl, ci, f64p, f64p]
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = []
        csc = [i64, i64, i32p, i32p, f64p, ci]
        L.oracle_add_level.restype = ci
        L.oracle_add_level.argtypes = [C.c_void_p, i64, i32p, i32p, f64p] + csc + csc + [ci, ci, ci, dbl, ci, ci, ci, dbl, ci]
        L.oracle_set_coarse.restype = ci
        L.oracle_set_coarse.argtypes = [C.c_void_p, i64, i32p, i32p, f64p, f64p]
        L.oracle_destroy.restype = None
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_coarse_solve.restype = None
        L.oracle_coarse_solve.argtypes = [C.c_void_p, f64p, f64p]
        L.oracle_cycle.restype = ci
        L.oracle_cycle.argtypes = [C.c_void_p, f64p, f64p, ci]
        L.oracle_solve.restype = ci
        L.oracle_solve.argtypes = [C.c_void_p, f64p, f64p, ci, ci, dbl, dbl, ci, C.c_void_p, ci, C.POINTER(ci), C.POINTER(ci)]
        L.oracle_precond.restype = ci
        L.oracle_precond.argtypes = [C.c_void_p, f64p, f64p, ci, ci]
        L.oracle_pcg.restype = ci
        L.oracle_pcg.argtypes = [C.c_void_p, f64p, f64p, ci, ci, ci, dbl, dbl, C.c_void_p, ci, C.POINTER(ci), C.POINTER(ci)]
        _lib = L
    return _lib


def _f64(v):
    return np.ascontiguousarray(v, dtype=np.float64)


def mul(a, x, adjoint=False):
    """``mul!(y, A, x)`` / ``mul!(y, A', x)`` with the stdlib's loop order."""
    y = np.empty(a.n if adjoint else a.m)
    lib().oracle_mul(a.m, a.n, a.colptr, a.rowval, a.nzval, int(adjoint), _f64(x), y)
    return y


def norm(v):
    v = _f64(v)
    return lib().oracle_norm(v.size, v)


def _cfg(s):
    """(kind, sweep, iter, omega) from a smoother config object of the host package (duck-typed)."""
    if s is None:
        return 0, 0, 0, 0.0
    return KIND[s.kind], SWEEP.get(getattr(s, "sweep_name", "symmetric"), 3), int(s.iter), float(getattr(s, "omega", 1.0))


def smooth(a, config, x, b, symmetry="hermitian"):
    """``config(A, x, b, symmetry)`` — in place on ``x`` (``src/smoother.jl:34-38``)."""
    k, sw, it, om = _cfg(config)
    x = np.ascontiguousarray(x, dtype=np.float64)
    rc = lib().oracle_smooth(a.n, a.colptr, a.rowval, a.nzval, k, sw, it, om, 0 if symmetry == "hermitian" else 1, x, _f64(b))
    if rc < 0:
        raise ZeroDivisionError(f"SingularException({-rc})")
    return x


class OracleHierarchy:
    """The oracle's view of a ``MultiLevel``: borrows the host arrays of ``ml`` (kept alive here)."""

    def __init__(self, ml):
        self._ml = ml
        self._keep = []
        L = lib()
        self._h = L.oracle_create()
        for lv in ml.levels:
            A = lv.A
            Pst, Padj = _storage(lv.P)
            Rst, Radj = _storage(lv.R)
            pk = _cfg(lv.presmoother.config)
            qk = _cfg(lv.postsmoother.config)
            sym = 0 if lv.presmoother.symmetry_name == "hermitian" else 1
            self._keep += [A, Pst, Rst]
            rc = L.oracle_add_level(self._h, A.n, A.colptr, A.rowval, A.nzval,
                                    Pst.m, Pst.n, Pst.colptr, Pst.rowval, Pst.nzval, Padj,
                                    Rst.m, Rst.n, Rst.colptr, Rst.rowval, Rst.nzval, Radj,
                                    *pk, *qk, sym)
            if rc < 0:
                raise ZeroDivisionError(f"SingularException({-rc})")
        fa = ml.final_A
        self._inv = np.asfortranarray(ml.coarse_solver.dense_operator(), dtype=np.float64)
        self._invflat = np.ascontiguousarray(self._inv.reshape(-1, order="F"))
        L.oracle_set_coarse(self._h, fa.n, fa.colptr, fa.rowval, fa.nzval, self._invflat)
        self.n = ml.levels[0].A.n if ml.levels else fa.n

    def __del__(self):
        try:
            if self._h:
                lib().oracle_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def solve(self, b, cycle="V", x0=None, maxiter=100, abstol=0.0, reltol=None, log=False, calculate_residual=True):
        b = _f64(b)
        if reltol is None:
            reltol = float(np.sqrt(np.finfo(np.float64).eps))
        x = np.zeros(self.n) if x0 is None else np.array(x0, dtype=np.float64)
        cap = maxiter + 2
        res = np.zeros(cap)
        nres, iters = C.c_int(0), C.c_int(0)
        lib().oracle_solve(self._h, x, b, CYCLE[cycle], int(maxiter), float(abstol), float(reltol), int(calculate_residual),
                           res.ctypes.data, cap, C.byref(nres), C.byref(iters))
        self.iters = iters.value
        return (x, res[: nres.value].copy()) if log else x

    def cycle(self, x, b, cycle="V"):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lib().oracle_cycle(self._h, x, _f64(b), CYCLE[cycle])
        return x

    def precond(self, b, cycle="V", init_zero=True):
        x = np.zeros(self.n)
        lib().oracle_precond(self._h, x, _f64(b), CYCLE[cycle], int(init_zero))
        return x

    def coarse_solve(self, b):
        x = np.zeros(len(b))
        lib().oracle_coarse_solve(self._h, x, _f64(b))
        return x

    def pcg(self, b, cycle="V", precond=True, maxiter=None, abstol=0.0, reltol=None, log=False):
        b = _f64(b)
        if reltol is None:
            reltol = float(np.sqrt(np.finfo(np.float64).eps))
        if maxiter is None:
            maxiter = self.n
        cap = maxiter + 2
        res = np.zeros(cap)
        x = np.zeros(self.n)
        nres, iters = C.c_int(0), C.c_int(0)
        lib().oracle_pcg(self._h, x, b, CYCLE[cycle], int(precond), int(maxiter), float(abstol), float(reltol),
                         res.ctypes.data, cap, C.byref(nres), C.byref(iters))
        self.iters = iters.value
        return (x, res[: nres.value].copy()) if log else x


def _storage(op):
    """(stored CSC, adjoint flag) of a level operator that may be a lazy Adjoint."""
    if hasattr(op, "parent"):
        return op.parent, 1
    return op, 0
