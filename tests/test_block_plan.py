"""CPU tests of the blocked Gauss-Seidel plan (csrc/device/block_plan.h) through the C-ABI entry b200amg_block_plan_check:
the plan's invariants (tiles / stages / steps partition the level, every dependency is either inside a tile in step order or
covered by a cross-tile requirement on an EARLIER ticket, the tile graph is acyclic) and the host emulation of the kernel's
sweep against the oracle's sequential sweep.  No GPU needed."""
import numpy as np
import pytest

import oracle


def _check(amg, A, sm, **kw):
    from algebraicmultigrid_jl_b200 import _devlib

    r = np.random.default_rng(A.n)
    x, b = r.standard_normal(A.n), r.standard_normal(A.n)
    sweep = {"forward": 1, "backward": 2, "symmetric": 3}[sm.sweep_name]
    st, msg, perm, xo = _devlib.block_plan_check(A, x, b, omega=getattr(sm, "omega", 1.0), sor=sm.kind == "sor", sweep=sweep, **kw)
    assert st["ok"] == 1, msg
    assert sorted(perm.tolist()) == list(range(A.n))
    ref = oracle.smooth(A, sm, x.copy(), b)
    assert np.abs(xo - ref).max() <= 1e-13 * np.abs(ref).max()
    # the pass sweep's layout (pass_plan.h: slabs, near / far codes, window, look-ahead rules) on the same plan
    st2, msg2, _, xo2 = _devlib.block_plan_check(A, x, b, omega=getattr(sm, "omega", 1.0), sor=sm.kind == "sor", sweep=sweep,
                                                 emulate_pass=True, **kw)
    assert st2["ok"] == 1, msg2
    assert np.abs(xo2 - ref).max() <= 1e-13 * np.abs(ref).max()
    return st


@pytest.mark.parametrize("dims", [(300,), (40, 33), (20, 17, 12), (40, 40, 40)])
def test_plan_and_emulated_sweep_on_stencils(amg, dims):
    A = amg.poisson(dims)
    for sm in (amg.GaussSeidel(), amg.GaussSeidel(amg.ForwardSweep()), amg.SOR(1.2, amg.BackwardSweep())):
        _check(amg, A, sm)


def test_plan_on_rs_hierarchy_levels_and_fe_matrix(amg, fx):
    ml = amg.ruge_stuben(amg.poisson((40, 40, 40)))
    multi = 0
    for lv in ml.levels:
        st = _check(amg, lv.A, amg.GaussSeidel())
        multi += st["tiles"] > 1
    assert multi >= 2          # the big levels really are cut into several tiles
    _check(amg, fx.matrix("thing"), amg.GaussSeidel())


@pytest.mark.parametrize("kw", [dict(tile_rows=700), dict(block_a=300, block_b=3), dict(stage_nnz=256, stage_rows=64, window=512, depth=4)])
def test_forced_tile_shapes(amg, kw):
    """contiguous tiles, hand-picked (a, b) blocks of the monotone coordinates, tiny stages / window: every shape must give
    a sound plan and the sequential sweep's result."""
    ml = amg.ruge_stuben(amg.poisson((32, 32, 32)))
    for A in (ml.levels[0].A, ml.levels[1].A):
        st = _check(amg, A, amg.GaussSeidel(), **kw)
        assert st["tiles"] > 1


def test_nonsymmetric_pattern_is_rejected(amg, fx):
    from algebraicmultigrid_jl_b200 import _devlib

    A = fx.sprand_plus_diag(300, 0.03, 5.0, seed=3)
    with pytest.raises(_devlib.B200AmgError):
        _devlib.block_plan_check(A)
