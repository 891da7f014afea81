"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol that
include/b200amg.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200amg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200amg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(amg):
    from algebraicmultigrid_jl_b200 import _devlib

    _devlib.build()
    L = _devlib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 27
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_devlib.SYMBOLS) == declared
    assert L.b200amg_version() == 100


def test_no_device_is_an_error_not_a_fallback(amg):
    from algebraicmultigrid_jl_b200 import _devlib

    if _devlib.device_count() > 0:
        return
    h = C.c_void_p()
    rc = _devlib.lib().b200amg_create(C.byref(h), 0)
    assert rc == -9 and not h.value
    assert b"no CPU fallback" in _devlib.lib().b200amg_last_error()
    A = amg.poisson(100)
    ml = amg.ruge_stuben(A)
    try:
        amg._solve(ml, np.ones(100))
    except _devlib.B200AmgError as e:
        assert e.code == -9
    else:
        raise AssertionError("the solve phase ran without a GPU")


def test_device_galerkin_backend_has_no_cpu_fallback(amg):
    """`set_galerkin_backend("device")` without a GPU must raise, never quietly use the host product."""
    import pytest

    from algebraicmultigrid_jl_b200 import _devlib

    with pytest.raises(ValueError):
        amg.set_galerkin_backend("cpu")
    L = _devlib.lib()
    assert L.b200amg_spgemm_fetch(None, None, None) == -5 or L.b200amg_spgemm_fetch(None, None, None) < 0   # nothing pending
    assert L.b200amg_spgemm_release() == 0
    if _devlib.device_count() > 0:
        return
    A = amg.poisson(50)
    try:
        amg.set_galerkin_backend("device")
        with pytest.raises(_devlib.B200AmgError) as e:
            amg.ruge_stuben(A)
        assert e.value.code == -9
    finally:
        amg.set_galerkin_backend("host")
    assert len(amg.ruge_stuben(A).levels) >= 1


def test_null_and_state_errors(amg):
    from algebraicmultigrid_jl_b200 import _devlib

    L = _devlib.lib()
    assert L.b200amg_create(None, 0) == -1
    assert L.b200amg_finalize(None) == -1
    assert L.b200amg_destroy(None) == 0
    assert L.b200amg_num_levels(None) == 0


def test_coarse_solver_host_callable_logic(amg, monkeypatch):
    """Host side of b200amg_set_coarse_callback (no GPU needed): above the dense limit QRSolver / LinearSolveWrapper keep a
    sparse factorisation (the reference does for any size, coarse_solver.jl:35-42,66-81) and are applied as callables;
    Pinv refuses; the ctypes trampoline hands numpy views of the staging vectors to the callable and turns an exception
    into a status instead of letting it cross the C boundary."""
    import pytest

    from algebraicmultigrid_jl_b200 import _devlib, coarse_solver as cs_mod

    A = amg.poisson((12, 12))
    dense = np.linalg.inv(A.todense())
    b = np.random.default_rng(0).standard_normal(A.n)
    monkeypatch.setattr(cs_mod, "DENSE_LIMIT", 100)
    for make in (amg.QRSolver, amg.LinearSolveWrapper(amg.UMFPACKFactorization())):
        cs = make(A)
        assert cs.dense_operator() is None
        x = np.empty(A.n)
        cs(x, b)
        assert np.abs(x - dense @ b).max() <= 1e-12 * np.abs(x).max()
        X = np.empty((A.n, 3))
        B3 = np.stack([b, 2 * b, -b], axis=1)
        cs(X, B3)
        assert np.abs(X - dense @ B3).max() <= 1e-11 * np.abs(X).max()
    with pytest.raises(ValueError, match="dense coarse operator"):
        amg.Pinv(A)
    monkeypatch.setattr(cs_mod, "DENSE_LIMIT", 16384)
    small = amg.QRSolver(A)
    x = np.empty(A.n)
    small(x, b)                                            # below the limit the callable form applies the dense operator
    assert np.abs(x - small.dense_operator() @ b).max() == 0.0

    # the trampoline: C pointers in, numpy views out, status back
    fn = _devlib._coarse_trampoline(lambda xx, bb: xx.__setitem__(Ellipsis, 2.0 * bb))
    cb = _devlib.COARSE_FN(fn)
    xb, bb = (C.c_double * 5)(), (C.c_double * 5)(*range(5))
    assert cb(None, 5, 1, xb, bb) == 0 and list(xb) == [0.0, 2.0, 4.0, 6.0, 8.0]

    def boom(xx, bb):
        raise RuntimeError("no factorisation")

    fn = _devlib._coarse_trampoline(boom)
    assert _devlib.COARSE_FN(fn)(None, 5, 1, xb, bb) == 1 and isinstance(fn.last_exception, RuntimeError)
