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


def elasticity_2d(nx, ny, E=1.0, nu=0.3):
    """Synthetic plane-strain linear elasticity on an ``nx x ny`` grid of unit Q1 (bilinear) elements, clamped on
    the edge x = 0 — the structured-grid analogue of the reference's only elasticity input, the 208-dof fixture
    ``test/lin_elastic_2d.jld2`` (``test/nns_test.jl:213-223``; the reference ships no generator).
    Returns ``(A, b, B)``: the SPD stiffness matrix on the free dofs (u, v interleaved per node, node index = i + (nx+1) j
    minus the clamped nodes), a right-hand side (unit downward traction on the edge x = nx) and the three rigid-body
    modes ``[1 0 -y; 0 1 x]`` restricted to the free dofs — the near-null-space ``B`` of ``smoothed_aggregation(A; B=B)``,
    laid out as ``create_nns_frame`` does for the beam (``test/nns_test.jl:138-164``)."""
    import numpy as np
    import scipy.sparse as sp

    from .sparse import SparseMatrixCSC

    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    mu = E / (2 * (1 + nu))
    D = np.array([[lam + 2 * mu, lam, 0.0], [lam, lam + 2 * mu, 0.0], [0.0, 0.0, mu]])
    # 8 x 8 element matrix of the unit square, 2 x 2 Gauss points
    gp = np.array([-1.0, 1.0]) / np.sqrt(3.0)
    xi_n = np.array([-1.0, 1.0, 1.0, -1.0])
    eta_n = np.array([-1.0, -1.0, 1.0, 1.0])
    Ke = np.zeros((8, 8))
    for xi in gp:
        for eta in gp:
            dN_dxi = 0.25 * xi_n * (1 + eta_n * eta)
            dN_deta = 0.25 * eta_n * (1 + xi_n * xi)
            dNx, dNy = 2.0 * dN_dxi, 2.0 * dN_deta          # unit element: d/dx = 2 d/dxi
            Bm = np.zeros((3, 8))
            Bm[0, 0::2] = dNx
            Bm[1, 1::2] = dNy
            Bm[2, 0::2] = dNy
            Bm[2, 1::2] = dNx
            Ke += 0.25 * (Bm.T @ D @ Bm)                      # det J = 1/4, weights 1
    Ke = 0.5 * (Ke + Ke.T)
    nnx, nny = nx + 1, ny + 1
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    n0 = (ii + nnx * jj).ravel()
    nodes = np.stack([n0, n0 + 1, n0 + 1 + nnx, n0 + nnx], axis=1)          # counter-clockwise
    dofs = np.empty((nodes.shape[0], 8), dtype=np.int64)
    dofs[:, 0::2] = 2 * nodes
    dofs[:, 1::2] = 2 * nodes + 1
    rows = np.repeat(dofs, 8, axis=1).ravel()
    cols = np.tile(dofs, (1, 8)).ravel()
    vals = np.tile(Ke.ravel(), nodes.shape[0])
    ndof = 2 * nnx * nny
    K = sp.coo_matrix((vals, (rows, cols)), shape=(ndof, ndof)).tocsc()
    node_i = np.arange(nnx * nny) % nnx
    free_nodes = np.nonzero(node_i != 0)[0]
    free = np.empty(2 * free_nodes.size, dtype=np.int64)
    free[0::2] = 2 * free_nodes
    free[1::2] = 2 * free_nodes + 1
    A = K[free][:, free].tocsc()
    A.sort_indices()
    x = (free_nodes % nnx).astype(np.float64)
    y = (free_nodes // nnx).astype(np.float64)
    B = np.zeros((free.size, 3))
    B[0::2, 0] = 1.0
    B[1::2, 1] = 1.0
    B[0::2, 2] = -y
    B[1::2, 2] = x
    f = np.zeros(ndof)
    edge = np.nonzero(node_i == nx)[0]
    w = np.ones(edge.size)
    w[[0, -1]] = 0.5
    f[2 * edge + 1] = -w / ny
    return SparseMatrixCSC.from_scipy(A), f[free], B


def elasticity_3d(nx, ny, nz, E=1.0, nu=0.3):
    """Synthetic 3-D linear elasticity on an ``nx x ny x nz`` grid of unit Q1 (trilinear, 8-node) elements, clamped on the
    face x = 0 — the volume analogue of :func:`elasticity_2d` (the reference ships no generator).  Returns
    ``(A, b, B)``: the SPD stiffness matrix on the free dofs (u, v, w interleaved per node, node index = i + (nx+1) j +
    (nx+1)(ny+1) k minus the clamped nodes; up to 81 entries per row), a right-hand side (unit downward traction on the
    face x = nx) and the SIX rigid-body modes — three translations and the rotations ``[0 -z y; z 0 -x; -y x 0]`` —
    restricted to the free dofs: the near-null-space ``B`` of ``smoothed_aggregation(A; B=B)``, one row per dof and one
    column per mode as ``create_nns_frame`` lays it out (``test/nns_test.jl:138-164``)."""
    import numpy as np
    import scipy.sparse as sp

    from .sparse import SparseMatrixCSC

    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    mu = E / (2 * (1 + nu))
    D = np.zeros((6, 6))
    D[:3, :3] = lam
    D[np.arange(3), np.arange(3)] = lam + 2 * mu
    D[np.arange(3, 6), np.arange(3, 6)] = mu
    # 24 x 24 element matrix of the unit cube, 2 x 2 x 2 Gauss points; local node order: x fastest, then y, then z
    sx = np.array([-1.0, 1.0, -1.0, 1.0, -1.0, 1.0, -1.0, 1.0])
    sy = np.array([-1.0, -1.0, 1.0, 1.0, -1.0, -1.0, 1.0, 1.0])
    sz = np.array([-1.0, -1.0, -1.0, -1.0, 1.0, 1.0, 1.0, 1.0])
    gp = np.array([-1.0, 1.0]) / np.sqrt(3.0)
    Ke = np.zeros((24, 24))
    for xi in gp:
        for eta in gp:
            for zeta in gp:
                dNx = 2.0 * 0.125 * sx * (1 + sy * eta) * (1 + sz * zeta)      # unit element: d/dx = 2 d/dxi
                dNy = 2.0 * 0.125 * sy * (1 + sx * xi) * (1 + sz * zeta)
                dNz = 2.0 * 0.125 * sz * (1 + sx * xi) * (1 + sy * eta)
                Bm = np.zeros((6, 24))
                Bm[0, 0::3] = dNx
                Bm[1, 1::3] = dNy
                Bm[2, 2::3] = dNz
                Bm[3, 0::3], Bm[3, 1::3] = dNy, dNx                          # gamma_xy
                Bm[4, 1::3], Bm[4, 2::3] = dNz, dNy                          # gamma_yz
                Bm[5, 0::3], Bm[5, 2::3] = dNz, dNx                          # gamma_zx
                Ke += 0.125 * (Bm.T @ D @ Bm)                                 # det J = 1/8, weights 1
    Ke = 0.5 * (Ke + Ke.T)
    nnx, nny, nnz_ = nx + 1, ny + 1, nz + 1
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    n0 = (ii + nnx * jj + nnx * nny * kk).ravel()
    off = np.array([0, 1, nnx, nnx + 1, nnx * nny, nnx * nny + 1, nnx * nny + nnx, nnx * nny + nnx + 1])
    nodes = n0[:, None] + off[None, :]
    dofs = np.empty((nodes.shape[0], 24), dtype=np.int64)
    for c in range(3):
        dofs[:, c::3] = 3 * nodes + c
    ndof = 3 * nnx * nny * nnz_
    K = sp.csc_matrix((ndof, ndof))
    slab = 100_000                             # assemble in slabs of elements: 576 triplets each
    for s in range(0, nodes.shape[0], slab):
        d = dofs[s:s + slab]
        rows = np.repeat(d, 24, axis=1).ravel()
        cols = np.tile(d, (1, 24)).ravel()
        vals = np.tile(Ke.ravel(), d.shape[0])
        K = K + sp.coo_matrix((vals, (rows, cols)), shape=(ndof, ndof)).tocsc()
    node_i = np.arange(nnx * nny * nnz_) % nnx
    free_nodes = np.nonzero(node_i != 0)[0]
    free = np.empty(3 * free_nodes.size, dtype=np.int64)
    for c in range(3):
        free[c::3] = 3 * free_nodes + c
    A = K[free][:, free].tocsc()
    A.sort_indices()
    x = (free_nodes % nnx).astype(np.float64)
    y = ((free_nodes // nnx) % nny).astype(np.float64)
    z = (free_nodes // (nnx * nny)).astype(np.float64)
    B = np.zeros((free.size, 6))
    B[0::3, 0] = 1.0
    B[1::3, 1] = 1.0
    B[2::3, 2] = 1.0
    B[1::3, 3], B[2::3, 3] = -z, y            # rotation about x: (0, -z, y)
    B[0::3, 4], B[2::3, 4] = z, -x            # rotation about y: (z, 0, -x)
    B[0::3, 5], B[1::3, 5] = -y, x            # rotation about z: (-y, x, 0)
    f = np.zeros(ndof)
    face = np.nonzero(node_i == nx)[0]
    fy, fz = (face // nnx) % nny, face // (nnx * nny)
    w = np.where((fy == 0) | (fy == ny), 0.5, 1.0) * np.where((fz == 0) | (fz == nz), 0.5, 1.0)
    f[3 * face + 2] = -w / (ny * nz)
    return SparseMatrixCSC.from_scipy(A), f[free], B
