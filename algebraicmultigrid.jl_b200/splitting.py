"""Ruge-Stuben C/F splitting (``/root/reference/src/splitting.jl``) — setup phase, host."""
from . import _hostlib

F_NODE, C_NODE, U_NODE = 0, 1, 2


class RS:
    def __call__(self, s):
        """Mutates ``s`` (removes its diagonal) exactly like the reference (``splitting.jl:20-23``)."""
        t = getattr(s, "_transpose_of", None)
        _hostlib.remove_diag(s)
        if t is not None:
            # S = copy(T') straight from ``Classical``: the pattern of S' after remove_diag! is T's pattern minus the
            # diagonal and the stored zeros — one filter pass instead of a transpose (checked by the entry count)
            tp = _hostlib.offdiag_pattern(t)
            s._transpose_of = None   # S no longer is T'
            if tp[1].shape[0] == s.nnz:
                return _hostlib.rs_cf_splitting(s, tp)
        return _hostlib.rs_cf_splitting(s, s.transpose())
