"""Standard aggregation (``/root/reference/src/aggregate.jl:12-134``) — setup phase, host."""
import numpy as np

from . import _hostlib
from .sparse import SparseMatrixCSC


class StandardAggregation:
    def __call__(self, s: SparseMatrixCSC) -> SparseMatrixCSC:
        """Returns ``AggOp`` (n_agg x n_fine); isolated nodes are empty columns (``aggregate.jl:118-131``)."""
        x, nagg = _hostlib.standard_aggregation(s)
        mask = x != -1
        colptr = np.concatenate(([0], np.cumsum(mask))).astype(np.int32)
        return SparseMatrixCSC(nagg, s.n, colptr, x[mask].astype(np.int32), np.ones(int(mask.sum())))
