"""The row-partitioned fine level on CPU: world_size = 2 and 3 `gloo` process groups run the same
distributed algorithm the CUDA engine runs over NCCL — halo exchange from the C++ partition plan
(b200amg_partition_plan, host-only), local Jacobi / residual / restriction on [owned | halo]
vectors, coarse_b gathered on rank 0, coarse_x windows sent back, prolongation — with numpy doing
the local arithmetic, and compare with the serial oracle on the same hierarchy."""
import os
import socket

import numpy as np
import pytest

import oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _local_rows(csr, r0, r1, lo, hi, halo_cols):
    """rows [r0, r1) of a scipy CSR with columns remapped to [owned | halo]"""
    import scipy.sparse as sp

    blk = csr[r0:r1].tocoo()
    owned = (blk.col >= lo) & (blk.col < hi)
    col = np.where(owned, blk.col - lo, (hi - lo) + np.searchsorted(halo_cols, blk.col))
    return sp.csr_matrix((blk.data, (blk.row, col)), shape=(r1 - r0, (hi - lo) + len(halo_cols)))


def _worker(rank, world, port, dims, method, out):
    import torch.distributed as dist

    import algebraicmultigrid_jl_b200 as amg
    from algebraicmultigrid_jl_b200 import _devlib

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch

    A = amg.poisson(dims)
    jac = amg.Jacobi(2.0 / 3.0)
    ml = (amg.ruge_stuben if method == "rs" else amg.smoothed_aggregation)(A, presmoother=jac, postsmoother=jac)
    lv = ml.levels[0]
    pl = _devlib.partition_plan(lv, rank, world)
    lo, hi = int(pl["row_split"][rank]), int(pl["row_split"][rank + 1])
    clo, chi = int(pl["coarse_split"][rank]), int(pl["coarse_split"][rank + 1])
    nloc, halo = hi - lo, pl["halo_cols"]
    As = lv.A.to_scipy().tocsr()
    Pm = (lv.P.materialize() if hasattr(lv.P, "materialize") else lv.P).to_scipy().tocsr()
    Rm = (lv.R.materialize() if hasattr(lv.R, "materialize") else lv.R).to_scipy().tocsr()
    Aloc = _local_rows(As, lo, hi, lo, hi, halo)
    Rloc = _local_rows(Rm, clo, chi, lo, hi, halo)
    cxlo, cxhi = int(pl["cx_lo"][rank]), int(pl["cx_hi"][rank])
    Ploc = Pm[lo:hi, cxlo:cxhi]
    assert Pm[lo:hi].nnz == Ploc.nnz                      # the window covers every referenced coarse entry
    d = As.diagonal()[lo:hi]

    def exchange(v):
        """v: [owned | halo] numpy vector, halo filled in place"""
        reqs, bufs = [], []
        for q in range(world):
            if q == rank:
                continue
            s0, s1 = pl["send_off"][q], pl["send_off"][q + 1]
            r0, r1 = pl["recv_off"][q], pl["recv_off"][q + 1]
            if s1 > s0:
                reqs.append(dist.isend(torch.from_numpy(v[pl["send_idx"][s0:s1]].copy()), q))
            if r1 > r0:
                t = torch.empty(int(r1 - r0), dtype=torch.float64)
                bufs.append((t, r0, r1))
                reqs.append(dist.irecv(t, q))
        for r in reqs:
            r.wait()
        for t, r0, r1 in bufs:
            v[nloc + r0: nloc + r1] = t.numpy()

    def jacobi(x, b, w=2.0 / 3.0):
        exchange(x)
        ax_off = Aloc @ x - d * x[:nloc]
        x[:nloc] = (1 - w) * x[:nloc] + w * ((b - ax_off) / d)

    rng = np.random.default_rng(0)
    bfull = rng.random(A.n)
    b = bfull[lo:hi].copy()
    x = np.zeros(nloc + len(halo))
    nc = Rm.shape[0]
    # ---- one V-cycle, level 1 distributed, levels >= 2 on rank 0 through the oracle ----
    jacobi(x, b)
    exchange(x)
    res = np.zeros_like(x)
    res[:nloc] = b - Aloc @ x
    exchange(res)
    cb_mine = Rloc @ res
    cb = np.zeros(nc)
    if rank == 0:
        cb[clo:chi] = cb_mine
        for q in range(1, world):
            c0, c1 = int(pl["coarse_split"][q]), int(pl["coarse_split"][q + 1])
            t = torch.empty(c1 - c0, dtype=torch.float64)
            dist.recv(t, q)
            cb[c0:c1] = t.numpy()
    else:
        dist.send(torch.from_numpy(cb_mine), 0)
    if rank == 0:
        sub = amg.MultiLevel(ml.levels[1:], ml.final_A, ml.coarse_solver, None, None, ml.workspace)
        cx = oracle.OracleHierarchy(sub).solve(cb, maxiter=1, calculate_residual=False)
        for q in range(1, world):
            dist.send(torch.from_numpy(cx[int(pl["cx_lo"][q]): int(pl["cx_hi"][q])].copy()), q)
        cxw = cx[cxlo:cxhi]
    else:
        t = torch.empty(cxhi - cxlo, dtype=torch.float64)
        dist.recv(t, 0)
        cxw = t.numpy()
    x[:nloc] += Ploc @ cxw
    jacobi(x, b)
    # ---- assemble and compare with the serial oracle cycle ----
    parts = [None] * world
    dist.all_gather_object(parts, x[:nloc])
    if rank == 0:
        xfull = np.concatenate(parts)
        ref = oracle.OracleHierarchy(ml).solve(bfull, maxiter=1, calculate_residual=False)
        out.put((float(np.abs(xfull - ref).max() / np.abs(ref).max()), int(pl["row_split"][1]), len(halo)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dims,method", [(2, (12, 12, 12), "rs"), (3, (10, 9, 8), "rs"), (2, (24, 24), "sa")])
def test_partitioned_cycle_matches_serial(world, dims, method):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dims, method, q)) for r in range(world)]
    for p in procs:
        p.start()
    err, split, nhalo = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-12, err
    assert 0 < split < int(np.prod(dims)) and nhalo > 0


# ---- more than one level partitioned (B200AMG_OPT_PART_LEVELS > 1) --------------------------------------------
def _worker_multi(rank, world, port, dims, method, plevels, out):
    """`plevels` finest levels split by rows: a level below a partitioned one inherits its parent's coarse-row
    ownership (b200amg_partition_plan_child), so the parent's restriction lands in the child's owned b with no
    communication and the prolongation reads the child's x after one halo exchange; the first unpartitioned level
    and everything below it run on rank 0 (through the oracle)."""
    import torch
    import torch.distributed as dist

    import algebraicmultigrid_jl_b200 as amg
    from algebraicmultigrid_jl_b200 import _devlib

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = amg.poisson(dims)
    jac = amg.Jacobi(2.0 / 3.0)
    ml = (amg.ruge_stuben if method == "rs" else amg.smoothed_aggregation)(A, presmoother=jac, postsmoother=jac)
    plevels = min(plevels, len(ml.levels))
    L = []   # per partitioned level: plan + local operators
    for k in range(plevels):
        lv = ml.levels[k]
        pl = _devlib.partition_plan(lv, rank, world) if k == 0 else _devlib.partition_plan(lv, rank, world, ml.levels[k - 1], L[k - 1]["pl"])
        if k > 0:
            assert np.array_equal(pl["row_split"], L[k - 1]["pl"]["coarse_split"])   # inherited ownership
        lo, hi = int(pl["row_split"][rank]), int(pl["row_split"][rank + 1])
        clo, chi = int(pl["coarse_split"][rank]), int(pl["coarse_split"][rank + 1])
        As = lv.A.to_scipy().tocsr()
        Pm = (lv.P.materialize() if hasattr(lv.P, "materialize") else lv.P).to_scipy().tocsr()
        Rm = (lv.R.materialize() if hasattr(lv.R, "materialize") else lv.R).to_scipy().tocsr()
        L.append(dict(pl=pl, lo=lo, hi=hi, clo=clo, chi=chi, nloc=hi - lo, halo=pl["halo_cols"], Pm=Pm, nc=Rm.shape[0],
                      Aloc=_local_rows(As, lo, hi, lo, hi, pl["halo_cols"]), Rloc=_local_rows(Rm, clo, chi, lo, hi, pl["halo_cols"]),
                      d=As.diagonal()[lo:hi]))
    for k in range(plevels):   # the prolongation of level k in the numbering of what lies below it
        E = L[k]
        if k + 1 < plevels:    # columns = the child's [owned | halo]
            Cn = L[k + 1]
            E["Ploc"] = _local_rows(E["Pm"], E["lo"], E["hi"], Cn["lo"], Cn["hi"], Cn["halo"])
        else:                  # columns = my coarse_x window
            cxlo, cxhi = int(E["pl"]["cx_lo"][rank]), int(E["pl"]["cx_hi"][rank])
            E["Ploc"] = E["Pm"][E["lo"]:E["hi"], cxlo:cxhi]
            assert E["Pm"][E["lo"]:E["hi"]].nnz == E["Ploc"].nnz

    def exchange(E, v):
        pl, nloc = E["pl"], E["nloc"]
        reqs, bufs = [], []
        for q in range(world):
            if q == rank:
                continue
            s0, s1 = pl["send_off"][q], pl["send_off"][q + 1]
            r0, r1 = pl["recv_off"][q], pl["recv_off"][q + 1]
            if s1 > s0:
                reqs.append(dist.isend(torch.from_numpy(v[pl["send_idx"][s0:s1]].copy()), q))
            if r1 > r0:
                t = torch.empty(int(r1 - r0), dtype=torch.float64)
                bufs.append((t, r0, r1))
                reqs.append(dist.irecv(t, q))
        for r in reqs:
            r.wait()
        for t, r0, r1 in bufs:
            v[nloc + r0: nloc + r1] = t.numpy()

    def jacobi(E, x, b, w=2.0 / 3.0):
        exchange(E, x)
        nloc = E["nloc"]
        ax_off = E["Aloc"] @ x - E["d"] * x[:nloc]
        x[:nloc] = (1 - w) * x[:nloc] + w * ((b - ax_off) / E["d"])

    def cycle(k, x, b):
        E = L[k]
        nloc, pl = E["nloc"], E["pl"]
        jacobi(E, x, b)
        exchange(E, x)
        res = np.zeros_like(x)
        res[:nloc] = b - E["Aloc"] @ x
        exchange(E, res)
        cb_mine = E["Rloc"] @ res                       # my coarse rows: no communication
        if k + 1 < plevels:
            Cn = L[k + 1]
            xc = np.zeros(Cn["nloc"] + len(Cn["halo"]))
            cycle(k + 1, xc, cb_mine)
            exchange(Cn, xc)                             # the one extra halo exchange of the prolongation
            x[:nloc] += E["Ploc"] @ xc
        else:
            cb = np.zeros(E["nc"])
            if rank == 0:
                cb[E["clo"]:E["chi"]] = cb_mine
                for q in range(1, world):
                    c0, c1 = int(pl["coarse_split"][q]), int(pl["coarse_split"][q + 1])
                    t = torch.empty(c1 - c0, dtype=torch.float64)
                    dist.recv(t, q)
                    cb[c0:c1] = t.numpy()
                sub = amg.MultiLevel(ml.levels[k + 1:], ml.final_A, ml.coarse_solver, None, None, ml.workspace)
                cx = oracle.OracleHierarchy(sub).solve(cb, maxiter=1, calculate_residual=False)
                for q in range(1, world):
                    dist.send(torch.from_numpy(cx[int(pl["cx_lo"][q]): int(pl["cx_hi"][q])].copy()), q)
                cxw = cx[int(pl["cx_lo"][0]): int(pl["cx_hi"][0])]
            else:
                dist.send(torch.from_numpy(cb_mine.copy()), 0)
                t = torch.empty(int(pl["cx_hi"][rank]) - int(pl["cx_lo"][rank]), dtype=torch.float64)
                dist.recv(t, 0)
                cxw = t.numpy()
            x[:nloc] += E["Ploc"] @ cxw
        jacobi(E, x, b)

    bfull = np.random.default_rng(0).random(A.n)
    x = np.zeros(L[0]["nloc"] + len(L[0]["halo"]))
    cycle(0, x, bfull[L[0]["lo"]:L[0]["hi"]].copy())
    parts = [None] * world
    dist.all_gather_object(parts, x[:L[0]["nloc"]])
    if rank == 0:
        ref = oracle.OracleHierarchy(ml).solve(bfull, maxiter=1, calculate_residual=False)
        out.put((float(np.abs(np.concatenate(parts) - ref).max() / np.abs(ref).max()), plevels, [len(E["halo"]) for E in L]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dims,method,plevels", [(2, (12, 12, 12), "rs", 2), (3, (12, 12, 12), "rs", 3), (2, (24, 24), "sa", 2),
                                                       (2, (300,), "rs", 4)])
def test_multi_level_partitioned_cycle_matches_serial(world, dims, method, plevels):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_multi, args=(r, world, port, dims if len(dims) > 1 else dims[0], method, plevels, q))
             for r in range(world)]
    for p in procs:
        p.start()
    err, used, halos = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-12, err
    assert used >= 2 and all(h > 0 for h in halos)


def test_partition_plan_properties(amg):
    from algebraicmultigrid_jl_b200 import _devlib

    A = amg.poisson((16, 16, 16))
    ml = amg.ruge_stuben(A)
    lv = ml.levels[0]
    world = 4
    plans = [_devlib.partition_plan(lv, r, world) for r in range(world)]
    n = A.n
    for r, pl in enumerate(plans):
        assert pl["row_split"][0] == 0 and pl["row_split"][-1] == n and np.all(np.diff(pl["row_split"]) > 0)
        assert pl["coarse_split"][0] == 0 and pl["coarse_split"][-1] == lv.R.shape[0]
        lo, hi = pl["row_split"][r], pl["row_split"][r + 1]
        h = pl["halo_cols"]
        assert np.all(np.diff(h) > 0) and not np.any((h >= lo) & (h < hi))
        # what q sends to r is exactly r's halo segment owned by q
        for q in range(world):
            if q == r:
                continue
            seg = h[pl["recv_off"][q]: pl["recv_off"][q + 1]]
            sent = plans[q]["send_idx"][plans[q]["send_off"][r]: plans[q]["send_off"][r + 1]] + plans[q]["row_split"][q]
            assert np.array_equal(seg, sent)
    # a 7-point stencil on (nnz-balanced, not plane-aligned) z-slabs: an interior rank receives about one
    # 16x16 plane from each side — never the whole vector
    assert 2 * 256 <= len(plans[1]["halo_cols"]) <= 3 * 256
