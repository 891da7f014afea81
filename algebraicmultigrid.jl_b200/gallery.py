"""Benchmark / test matrices (``/root/reference/src/gallery.jl:1-63``)."""
from . import _hostlib


def poisson(sz):
    """``poisson(n)``: tridiag(-1, 2, -1); ``poisson((n1,..,nN))``: (2N+1)-point stencil with centre
    2N and neighbours -1, Dirichlet truncation, first index fastest (``gallery.jl:42-63``)."""
    if isinstance(sz, (int,)) or hasattr(sz, "__index__"):
        dims = (int(sz),)
    else:
        dims = tuple(int(s) for s in sz)
    return _hostlib.poisson(dims)
