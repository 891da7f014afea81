"""algebraicmultigrid.jl_b200 — B200-native AMG solve-phase engine behind AlgebraicMultigrid.jl's API surface.

Host side (this package): the reference's names — ``ruge_stuben``, ``smoothed_aggregation``,
``MultiLevel``, ``_solve`` / ``_solve_`` (``_solve!``), ``aspreconditioner`` / ``ldiv_`` (``ldiv!``),
``setup_smoother`` / ``smooth_`` (``smooth!``), ``GaussSeidel`` / ``Jacobi`` / ``SOR``, ``V`` / ``W`` / ``F``,
``Pinv`` / ``QRSolver``, ``poisson``.  Setup runs on the host (C++ library ``libb200amg_setup.so``);
the solve phase runs on the GPU through the C-ABI in ``include/b200amg.h`` (``libb200amg.so``).
There is no CPU fallback for the solve phase: without the CUDA library / a GPU the solve entry
points raise.
"""
from . import _hostlib
from ._hostlib import set_galerkin_backend
from .aggregate import StandardAggregation
from .aggregation import DiagonalWeighting, JacobiProlongation, LocalWeighting, fit_candidates, smoothed_aggregation
from .classical import direct_interpolation, ruge_stuben
from .coarse_solver import LinearSolveWrapper, Pinv, QRSolver, UMFPACKFactorization
from .gallery import elasticity_2d, elasticity_3d, poisson
from .multilevel import (F, Level, MultiLevel, MultiLevelWorkspace, RugeStubenAMG, SmoothedAggregationAMG, V, W,
                         _solve, _solve_, grid_complexity, init, operator_complexity, solve, solve_)
from .preconditioner import Preconditioner, aspreconditioner, backslash, cg, ldiv_, mul_
from .precs import RugeStubenPreconBuilder, SmoothedAggregationPreconBuilder
from .smoother import (SOR, BackwardSweep, ForwardSweep, GaussSeidel, Jacobi, SingularException, SymmetricSweep,
                       setup_smoother, smooth_)
from .sparse import Adjoint, SparseMatrixCSC, adjoint, nnz, size
from .splitting import RS
from .strength import Classical, SymmetricStrength
from .utils import Hermitian, HermitianSymmetry, NoSymmetry, Symmetric, approximate_spectral_radius

__all__ = [n for n in dir() if not n.startswith("__")]
