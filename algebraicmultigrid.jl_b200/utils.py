"""Symmetry tags (``/root/reference/src/utils.jl:1-19``): they select the smoother variant."""


class NoSymmetry:
    def __repr__(self):
        return "NoSymmetry()"


class HermitianSymmetry:
    def __repr__(self):
        return "HermitianSymmetry()"


class Symmetric:
    """Wrapper mirroring ``LinearAlgebra.Symmetric`` so ``get_symmetry_and_data`` has something to unwrap."""

    def __init__(self, data):
        self.data = data


Hermitian = Symmetric


def get_symmetry_and_data(a):
    """``src/utils.jl:7-19``: unwrap Symmetric/Hermitian -> HermitianSymmetry, anything else -> NoSymmetry."""
    if isinstance(a, Symmetric):
        return a.data, HermitianSymmetry()
    return a, NoSymmetry()
