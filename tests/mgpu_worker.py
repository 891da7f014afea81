"""One rank of the multi-GPU parity check (launched by tests/test_multigpu.py through torchrun):
the row-partitioned fine level over NCCL against the CPU oracle of the unpartitioned hierarchy."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import algebraicmultigrid_jl_b200 as amg  # noqa: E402
import oracle  # noqa: E402
from algebraicmultigrid_jl_b200 import _devlib  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    jac = amg.Jacobi(2.0 / 3.0)
    failures = []
    # (method, grid, partitioned levels, halo exchange over peer memory [csrc/device/peer_halo.cuh] or over NCCL send / recv)
    cases = [("rs", (40, 40, 40), 1, 1), ("rs", (40, 40, 40), 2, 1), ("rs", (40, 40, 40), 3, 1), ("rs", (40, 40, 40), 3, 0), ("sa", (96, 96), 1, 1),
             ("sa", (96, 96), 2, 1), ("sa", (96, 96), 2, 0), ("rs", (1000,), 1, 1), ("rs", (1000,), 4, 1)]
    if os.environ.get("MGPU_CASES"):          # e.g. MGPU_CASES=1,2 : a subset, for quick experiments
        cases = [cases[int(v)] for v in os.environ["MGPU_CASES"].split(",")]
    for method, dims, plevels, peer in cases:
        os.environ["B200AMG_PEER_HALO"] = str(peer)      # read when the hierarchy is finalized on the device
        A = amg.poisson(dims if len(dims) > 1 else dims[0])
        build = amg.ruge_stuben if method == "rs" else amg.smoothed_aggregation
        ml = build(A, presmoother=jac, postsmoother=jac)
        box = [_devlib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ml.partition(rank, world, box[0], levels=plevels)
        b = np.random.default_rng(3).random(A.n)
        H = oracle.OracleHierarchy(build(A, presmoother=jac, postsmoother=jac)) if rank == 0 else None
        for cyc, cname in ((amg.V(), "V"), (amg.W(), "W"), (amg.F(), "F")):
            x1 = amg._solve(ml, b, cyc, maxiter=1, calculate_residual=False)
            x, hist = amg._solve(ml, b, cyc, log=True, maxiter=40)
            p = amg.backslash(amg.aspreconditioner(ml, cyc), b)
            # every rank must hold the same assembled x
            t = torch.from_numpy(x.copy()).cuda()
            dist.broadcast(t, src=0)
            same = bool(np.array_equal(t.cpu().numpy(), x))
            if rank == 0:
                r1 = H.solve(b, cycle=cname, maxiter=1, calculate_residual=False)
                xr, histr = H.solve(b, cycle=cname, log=True, maxiter=40)
                e1 = np.abs(x1 - r1).max() / np.abs(r1).max()
                e = np.linalg.norm(x - xr) / np.linalg.norm(xr)
                ep = np.abs(p - H.precond(b, cycle=cname)).max() / np.abs(r1).max()
                ok = e1 < 1e-11 and e < 1e-9 and ep < 1e-11 and len(hist) == len(histr) and np.allclose(hist, histr, rtol=1e-6) and same
                print(f"[mgpu] {method} {dims} part_levels={plevels} {'peer' if ml.device().comm_stats()['peer_halo'] else 'nccl'} {cname}: cycle {e1:.1e} solve {e:.1e} precond {ep:.1e} iters {len(hist) - 1}/{len(histr) - 1} "
                      f"same_on_all_ranks={same} {'OK' if ok else 'FAIL'}", flush=True)
                if not ok:
                    failures.append((method, dims, plevels, cname))
            elif not same:
                failures.append(("rank", rank))
        # device-resident CG preconditioned by the PARTITIONED V-cycle (b200amg_pcg on a row-partitioned handle) against the oracle's
        xc, info = amg.cg(A, b, Pl=amg.aspreconditioner(ml), reltol=1e-10, log=True)
        tc = torch.from_numpy(xc.copy()).cuda()
        dist.broadcast(tc, src=0)
        same = bool(np.array_equal(tc.cpu().numpy(), xc))
        if rank == 0:
            xcr = H.pcg(b, reltol=1e-10)
            ec = np.linalg.norm(xc - xcr) / np.linalg.norm(xcr)
            okc = info["iters"] == H.iters and ec < 1e-8 and same
            print(f"[mgpu] {method} {dims} part_levels={plevels} pcg: iters {info['iters']}/{H.iters} x {ec:.1e} same_on_all_ranks={same} {'OK' if okc else 'FAIL'}",
                  flush=True)
            if not okc:
                failures.append((method, dims, plevels, "pcg"))
        elif not same:
            failures.append(("rank", rank, "pcg"))
        if rank == 0:
            print("[mgpu] partition:", ml.device().partition_info(), flush=True)
        ml.release()
    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    dist.barrier()
    dist.destroy_process_group()
    if flag.item():
        print("[mgpu] FAILED", failures, flush=True)
        sys.exit(1)
    if rank == 0:
        print("[mgpu] ALL OK", flush=True)


if __name__ == "__main__":
    main()
