"""Symmetry tags (``/root/reference/src/utils.jl:1-19``: they select the smoother variant) and the spectral-radius estimate of
the setup phase (``:25-145``)."""
import numpy as np



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


# ---- approximate_spectral_radius (``src/utils.jl:25-145``) — setup phase, host -----------------------------------------
def _matvec(A, v):
    return A.matvec(v) if hasattr(A, "matvec") else np.asarray(A) @ v


def approximate_eigenvalues(A, tol, maxiter, symmetric, v0):
    """One Arnoldi run of ``maxiter`` steps from ``v0`` (``src/utils.jl:77-118``): returns the eigenvectors and eigenvalues of
    the square Hessenberg matrix, the (maxiter + 1) x maxiter Hessenberg matrix, the Krylov basis and the breakdown flag."""
    n = A.shape[0]
    v0 = v0 / np.linalg.norm(v0)
    H = np.zeros((maxiter + 1, maxiter), dtype=v0.dtype)
    V = [v0]
    breakdown = np.finfo(np.float64).eps * 1e6
    flag = False
    for j in range(maxiter):
        w = _matvec(A, V[-1]).astype(v0.dtype, copy=True)
        for i, v in enumerate(V):                       # modified Gram-Schmidt
            H[i, j] = np.dot(np.conj(v), w)
            w -= H[i, j] * v
        H[j + 1, j] = np.linalg.norm(w)
        if abs(H[j + 1, j]) < breakdown:
            flag = True
            if H[j + 1, j] != 0:
                V.append(w / H[j + 1, j])
            break
        V.append(w / H[j + 1, j])
    m = maxiter
    eigs, vects = np.linalg.eig(H[:m, :m])
    return vects, eigs, H, V, flag


def approximate_spectral_radius(A, tol=0.01, maxiter=15, restart=5, rng=None):
    """``approximate_spectral_radius(A, tol, maxiter, restart)`` (``src/utils.jl:25-55``): restarted Arnoldi estimate of
    max |eig(A)| from a random start vector (``rng``: a ``numpy.random.Generator``; the reference uses the global RNG)."""
    rng = np.random.default_rng() if rng is None else rng
    n = A.shape[0]
    v0 = rng.random(A.shape[1])
    maxiter = min(n, maxiter)
    ev = np.zeros(maxiter)
    max_index = 0
    for _ in range(restart + 1):
        evect, ev, H, V, flag = approximate_eigenvalues(A, tol, maxiter, False, v0)
        nvecs = ev.shape[0]
        X = np.stack(V[:nvecs], axis=1) if len(V) >= nvecs else np.stack(V + [np.zeros_like(V[0])] * (nvecs - len(V)), axis=1)
        max_index = int(np.argmax(np.abs(ev)))
        error = H[nvecs, nvecs - 1] * evect[-1, max_index]
        v_new = X @ evect[:, max_index]
        # a dominant eigenvalue of a real matrix with a real eigenvector keeps the iteration real (what the reference's
        # in-place ``mul!(v0, X, evect[:, max_index])`` requires); a genuinely complex pair continues in complex arithmetic
        v0 = v_new.real if np.abs(v_new.imag).max(initial=0.0) <= 1e-14 * max(np.abs(v_new).max(initial=0.0), 1e-300) else v_new
        if abs(ev[max_index]) == 0 or abs(error) / abs(ev[max_index]) < tol or flag:
            break
    return float(abs(ev[max_index]))
