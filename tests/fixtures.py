"""Loader for tests/golden/fixtures.npz (the reference's test matrices, converted by
tests/golden/make_fixtures.py) plus small helpers shared by the test modules."""
import os

import numpy as np

import algebraicmultigrid_jl_b200 as amg

_NPZ = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures.npz"))


def matrix(name):
    m, n = _NPZ[name + "_shape"]
    return amg.SparseMatrixCSC.from_julia(int(m), int(n), _NPZ[name + "_colptr"], _NPZ[name + "_rowval"], _NPZ[name + "_nzval"])


def array(name):
    return _NPZ[name].copy()


def spdiagm_tridiag(n):
    """spdiagm(0 => 2*ones(N), -1 => -ones(N-1), 1 => -ones(N-1)) (test/sa_tests.jl:318)."""
    return amg.poisson(n)


def sprand_plus_diag(n, density, shift, seed):
    """Analogue of `sprand(N,N,density) + shift*I` (test/test_smoothers.jl:10, test_regression.jl:75):
    uniform(0,1) entries at random positions.  Julia's RNG stream cannot be reproduced; the tests
    that use this only need a diagonally dominant non-symmetric matrix."""
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)
    a = sp.random(n, n, density=density, random_state=rng, data_rvs=rng.random, format="csc") + shift * sp.identity(n, format="csc")
    return amg.SparseMatrixCSC.from_scipy(a)
