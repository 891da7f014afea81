"""Ruge-Stuben C/F splitting (``/root/reference/src/splitting.jl``) — setup phase, host."""
from . import _hostlib

F_NODE, C_NODE, U_NODE = 0, 1, 2


class RS:
    def __call__(self, s):
        """Mutates ``s`` (removes its diagonal) exactly like the reference (``splitting.jl:20-23``)."""
        _hostlib.remove_diag(s)
        return _hostlib.rs_cf_splitting(s, s.transpose())
