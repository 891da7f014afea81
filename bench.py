#!/usr/bin/env python
"""bench.py — V-cycle iterations/s and fine-level SpMV GB/s of the B200 AMG solve phase.

Workload (BASELINE.json configs[2], the one `metric` is quoted on): 3D poisson((256,256,256)),
n = 16 777 216, nnz = 117 047 296, fp64, ruge_stuben hierarchy with the default symmetric
Gauss-Seidel smoothers, V-cycle.  A "step" is one iteration of `_solve!`
(/root/reference/src/multilevel.jl:178-193): one V-cycle + the convergence residual r = b - A x + its norm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 256]
                    [--method rs|sa] [--smoother gs|jacobi]

N > 1 (torchrun, one rank per GPU): the finest --part-levels levels (default 3) are row-partitioned (config C4,
Jacobi smoother).
`--impl reference` times the CPU restatement of the reference (oracle/, kind "port": the reference is
pure Julia and no Julia runtime exists in the image) on the host cores; it is single-threaded
because the reference's solve phase is (src/smoother.jl:73-90, README.md:120).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# torchrun exports OMP_NUM_THREADS=1 to every rank; the host-side setup and upload (OpenMP) would then run on one core.
# Give each rank its share of the host cores instead (before any OpenMP runtime is loaded).
_world = int(os.environ.get("WORLD_SIZE", "1"))
if _world > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // _world))

import numpy as np  # noqa: E402

SMI_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={SMI_QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(args, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the roofline kernel, from the committed ncu capture
    (profiles/r01_ncu_traffic.json); only for the workload that capture was taken on."""
    if not (world == 1 and args.size == 256 and args.dim == 3):
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as f:
            t = json.load(f)["csr_stream_kernel<1,1>@poisson256"]
        return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        return None


def build_problem(args):
    import algebraicmultigrid_jl_b200 as amg

    n1 = args.size
    dims = (n1, n1, n1) if args.dim == 3 else (n1, n1)
    t0 = time.time()
    A = amg.poisson(dims)
    if args.smoother == "jacobi":
        sm = amg.Jacobi(2.0 / 3.0)
        kw = dict(presmoother=sm, postsmoother=sm)
    else:
        kw = {}
    ml = amg.ruge_stuben(A, **kw) if args.method == "rs" else amg.smoothed_aggregation(A, **kw)
    b = A.matvec(np.ones(A.n))
    return amg, A, ml, b, time.time() - t0


def workload_name(args):
    dims = "x".join([str(args.size)] * args.dim)
    meth = "ruge_stuben" if args.method == "rs" else "smoothed_aggregation"
    sm = "SymmetricGaussSeidel(iter=1)" if args.smoother == "gs" else "Jacobi(2/3,iter=1)"
    return f"poisson(({dims})) fp64, {meth} V-cycle, pre/post {sm}, b=A*ones, x0=0"


def bench_config(args, A, nlevels):
    """The SAME dict in both arms (`--impl ours` / `--impl reference`): what the workload is, nothing about how it is run."""
    return {"workload": workload_name(args), "n": A.n, "nnz": A.nnz, "levels": nlevels}


def bytes_spmv(n, nnz, vb=8):       # values (vb bytes each: 8 = fp64, 4 = the lossless binary32 copy), int32 column
    return (vb + 4) * nnz + 4 * (n + 1) + 16 * n      # indices + row pointers, x counted once


def bytes_residual(n, nnz, vb=8):   # also one Jacobi sweep, one Gauss-Seidel direction
    return (vb + 4) * nnz + 4 * (n + 1) + 24 * n


def cycle_bytes(infos, stor, smoother):
    """Algorithmic bytes of ONE `_solve!` iteration (V-cycle + convergence residual) over the whole hierarchy, with the value
    width every level is actually stored in: per level (pre + post sweeps) x sweep + residual + restriction + prolongation
    (SURVEY §8d); Gauss-Seidel sweeps read the fp64 values (their kernels have no binary32 path)."""
    tot = 0
    sweeps = 4 if smoother == "gs" else 2          # symmetric GS = 2 directions pre + 2 post; Jacobi 1 + 1
    for i, (lv, st) in enumerate(zip(infos, stor)):
        if lv["nnz_p"] == 0 and i == len(infos) - 1:
            continue                                # the coarsest matrix: dense solve
        n, nnz, nc = lv["n"], lv["nnz_a"], infos[i + 1]["n"] if i + 1 < len(infos) else 0
        vb_s = 8 if smoother == "gs" else (st["A"] or 8)
        tot += sweeps * bytes_residual(n, nnz, vb_s) + bytes_residual(n, nnz, st["A"] or 8)
        tot += ((st["R"] or 8) + 4) * lv["nnz_p"] + 4 * (nc + 1) + 8 * n + 8 * nc
        tot += ((st["P"] or 8) + 4) * lv["nnz_p"] + 4 * (n + 1) + 8 * nc + 16 * n
    if infos:
        tot += bytes_residual(infos[0]["n"], infos[0]["nnz_a"], stor[0]["A"] or 8)      # the convergence residual
    return tot


def run_reference(args):
    """CPU arm: the oracle port of the reference's `_solve!`, one V-cycle iteration per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle

    if args.gpus > 1 and args.smoother == "gs":
        args.smoother = "jacobi"          # same workload as our arm at N > 1 (config C4)
    amg, A, ml, b, t_setup = build_problem(args)
    H = oracle.OracleHierarchy(ml)
    budget = float(os.environ.get("B200AMG_REF_BUDGET_S", "240"))
    x = np.zeros(A.n)
    # one untimed iteration tells what the requested warm-up + steps cost; both are then honoured as far as the budget allows
    t_start = time.time()
    x, h0 = H.solve(b, x0=x, maxiter=1, reltol=0.0, log=True)
    per = max(time.time() - t_start, 1e-9)
    hist = list(h0)
    warm = max(0, min(args.warmup - 1, int(0.25 * budget / per)))
    if warm:
        x, hw = H.solve(b, x0=x, maxiter=warm, reltol=0.0, log=True)
        hist += list(hw[1:])
    warm += 1
    steps = max(1, min(args.steps, int((budget - warm * per) / per)))
    t0 = time.time()
    x, ht = H.solve(b, x0=x, maxiter=steps, reltol=0.0, log=True)
    dt = time.time() - t0
    hist += list(ht[1:])
    val = steps / dt
    line = {
        "impl": "reference", "metric": "V-cycle iterations/s", "value": val, "unit": "V-cycles/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, A, len(ml)),
        "cpu_baseline": {"value": val, "unit": "V-cycles/s", "cores": 1, "kind": "port",
                         "sample": f"{steps} `_solve!` iterations (V-cycle + residual + norm) of the full workload after {warm} warm-up "
                                   "iterations, single thread (the reference solve phase is single-threaded), gcc -O2 -ffp-contract=off"},
        "e2e": {"value": val, "unit": "V-cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores_available": os.cpu_count(), "setup_s": t_setup,
        # what the GPU arm can be compared with: the same iteration counts from x0 = 0 (bench.py's `parity` block does that)
        "residual_history_first_last": [float(hist[0]), float(hist[-1])], "iterations_from_x0": len(hist) - 1,
        "max_abs_err_vs_ones": float(np.abs(x - 1.0).max()),
    }
    print(json.dumps(line), flush=True)


def other_configs(amg, torch, local):
    """Quick, driver-visible numbers for the BASELINE configs that are not the headline: C2 (2-D poisson((1024,1024)),
    smoothed aggregation + Jacobi) and C5 (test/lin_elastic_2d.jld2, SA with near-null-space as CG preconditioner).  Both are
    small against the 126 MB L2, so each is timed twice: back-to-back iterations (operators L2-resident) and single
    iterations with an L2 flush (a 512 MB write) in front of each."""
    out = {}
    dev_t = torch.device("cuda", local)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device=dev_t)     # 512 MB > 4 x L2

    def time_solve(dev, x, b, iters, cold):
        stream = torch.cuda.ExternalStream(dev.stream(), device=dev_t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not cold:
            torch.cuda.synchronize()
            e0.record(stream)
            dev.solve(x, b, 0, iters, 0.0, 0.0, True)
            e1.record(stream)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            torch.cuda.synchronize()
            e0.record(stream)
            dev.solve(x, b, 0, 1, 0.0, 0.0, True)
            e1.record(stream)
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters

    try:                                   # ---- C2
        t0 = time.time()
        A = amg.poisson((1024, 1024))
        jac = amg.Jacobi(2.0 / 3.0)
        ml = amg.smoothed_aggregation(A, presmoother=jac, postsmoother=jac)
        t_setup = time.time() - t0
        dev = ml.device()
        b = torch.from_numpy(A.matvec(np.ones(A.n))).to(dev_t)
        x = torch.zeros(A.n, dtype=torch.float64, device=dev_t)
        dev.solve(x, b, 0, 5, 0.0, 0.0, True)
        x.zero_()
        hot = time_solve(dev, x, b, 50, cold=False)
        cold = time_solve(dev, x, b, 10, cold=True)
        spmv_hot = dev.time_kernel(0, 0, reps=50)
        spmv_cold = dev.time_kernel(0, 0, reps=10, flush_l2=True)
        out["C2"] = {"workload": "poisson((1024,1024)) fp64, smoothed_aggregation V-cycle, pre/post Jacobi(2/3,iter=1)", "n": A.n, "nnz": A.nnz,
                     "levels": dev.nlevels, "setup_s": t_setup,
                     "iters_per_s_l2_resident": 1e3 / hot, "ms_per_iter_l2_resident": hot,
                     "iters_per_s_l2_flushed": 1e3 / cold, "ms_per_iter_l2_flushed": cold,
                     "fine_spmv_gbs_l2_resident": bytes_spmv(A.n, A.nnz, dev.storage_info(0)["A"] or 8) / (spmv_hot * 1e-3) / 1e9,
                     "fine_spmv_gbs_l2_flushed": bytes_spmv(A.n, A.nnz, dev.storage_info(0)["A"] or 8) / (spmv_cold * 1e-3) / 1e9,
                     "fine_value_bytes": dev.storage_info(0)["A"],
                     "l2": "the whole hierarchy (0.09 GB) fits the 126 MB L2: the resident figures are L2 rates; flushed = a 512 MB "
                           "write before every single-iteration call (launch-latency-bound: ~60 kernels of 5-30 us)"}
        ml.release()
    except Exception as exc:
        out["C2"] = {"error": str(exc)[:300]}
    try:                                   # ---- C5
        npz = np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))
        m, n = npz["elastic_shape"]
        A = amg.SparseMatrixCSC.from_julia(int(m), int(n), npz["elastic_colptr"], npz["elastic_rowval"], npz["elastic_nzval"])
        bvec, B = npz["elastic_b"].copy(), npz["elastic_B"].copy()
        ml = amg.smoothed_aggregation(A, B=B)
        dev = ml.device()
        b = torch.from_numpy(bvec).to(dev_t)
        x = torch.zeros(A.n, dtype=torch.float64, device=dev_t)
        hist, iters = dev.pcg(x, b, 0, 200, 0.0, 1e-10)      # warm-up (captures the cycle graph)
        reps = 20
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            x.zero_()
            hist, iters = dev.pcg(x, b, 0, 200, 0.0, 1e-10)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        xh = x.cpu().numpy()
        out["C5"] = {"workload": "test/lin_elastic_2d.jld2 (208 x 208, nnz 2632), smoothed_aggregation with near-null-space B (208 x 3), "
                                 "device-resident CG preconditioned by one V-cycle, reltol 1e-10", "n": A.n, "nnz": A.nnz,
                     "levels": dev.nlevels, "cg_iterations": int(iters), "ms_per_solve": 1e3 * dt,
                     "cg_iterations_per_s": iters / dt, "relative_residual": float(np.linalg.norm(A.matvec(xh) - bvec) / np.linalg.norm(bvec)),
                     "note": "latency-bound by construction (208 unknowns): the figure is the launch / graph-replay latency of one "
                             "preconditioned CG iteration, not a bandwidth number"}
        ml.release()
    except Exception as exc:
        out["C5"] = {"error": str(exc)[:300]}
    try:                                   # ---- the headline workload with MULTICOLOUR Gauss-Seidel: NOT parity, the bandwidth-bound ceiling
        os.environ["B200AMG_GS_MULTICOLOR"] = "1"
        try:
            A = amg.poisson((256, 256, 256))
            t0 = time.time()
            ml = amg.ruge_stuben(A)
            t_setup = time.time() - t0
            dev = ml.device()
        finally:
            os.environ.pop("B200AMG_GS_MULTICOLOR", None)
        b = torch.from_numpy(A.matvec(np.ones(A.n))).to(dev_t)
        x = torch.zeros(A.n, dtype=torch.float64, device=dev_t)
        dev.solve(x, b, 0, 3, 0.0, 0.0, True)
        x.zero_()
        its = 20
        ms_it = time_solve(dev, x, b, its, cold=False)
        x.zero_()
        hist, _ = dev.solve(x, b, 0, its, 0.0, 0.0, True)
        out["C3_multicolour_gauss_seidel"] = {
            "workload": "poisson((256x256x256)) fp64, ruge_stuben V-cycle, pre/post symmetric Gauss-Seidel relaxed COLOUR BY COLOUR (greedy colouring; "
                        "B200AMG_GS_MULTICOLOR=1) instead of in index order",
            "parity": "NONE - a different row order than gs! (src/smoother.jl:73-90): same fixed point, different iterates; shown as the "
                      "bandwidth-bound ceiling next to the exact-order headline (SURVEY 7.2-A)",
            "iters_per_s": 1e3 / ms_it, "ms_per_iter": ms_it, "wavefronts_per_level": [dev.level_info(i)["wavefronts"] for i in range(dev.nlevels)],
            "residual_history_first_last": [float(hist[0]), float(hist[-1])], "max_abs_err_vs_ones": float((x - 1.0).abs().max().item())}
        ml.release()
    except Exception as exc:
        out["C3_multicolour_gauss_seidel"] = {"error": str(exc)[:300]}
    try:                                   # ---- synthetic 3-D elasticity (north_star: "synthetic Poisson / elasticity matrices")
        t0 = time.time()
        A, bvec, B = amg.elasticity_3d(48, 48, 48)
        t_gen = time.time() - t0
        res = {}
        for name, kw in (("gauss_seidel", {}), ("jacobi_0.5", {"presmoother": amg.Jacobi(0.5), "postsmoother": amg.Jacobi(0.5)})):
            t0 = time.time()
            ml = amg.smoothed_aggregation(A, B=B, **kw)
            t_setup = time.time() - t0
            dev = ml.device()
            b = torch.from_numpy(bvec).to(dev_t)
            x = torch.zeros(A.n, dtype=torch.float64, device=dev_t)
            hist, iters = dev.pcg(x, b, 0, 300, 0.0, 1e-8)       # warm-up (captures the cycle graph)
            x.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            hist, iters = dev.pcg(x, b, 0, 300, 0.0, 1e-8)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            xh = x.cpu().numpy()
            res[name] = {"levels": dev.nlevels, "setup_s": t_setup, "cg_iterations": int(iters), "solve_ms": 1e3 * dt,
                         "cg_iterations_per_s": iters / dt,
                         "relative_residual": float(np.linalg.norm(A.matvec(xh) - bvec) / np.linalg.norm(bvec)),
                         "fine_spmv_gbs_l2_flushed": bytes_spmv(A.n, A.nnz, dev.storage_info(0)["A"] or 8) / (dev.time_kernel(0, 0, reps=10, flush_l2=True) * 1e-3) / 1e9}
            ml.release()
        out["elasticity_3d"] = {"workload": "elasticity_3d(48,48,48): Q1 hexahedra, 3 dofs per node, clamped face, smoothed_aggregation with the six "
                                            "rigid-body modes as near-null-space, device-resident CG preconditioned by one V-cycle, reltol 1e-8",
                                "n": A.n, "nnz": A.nnz, "generate_s": t_gen, "smoothers": res}
    except Exception as exc:
        out["elasticity_3d"] = {"error": str(exc)[:300]}
    del flush
    return out


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torchrun (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the solve phase has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if args.smoother == "gs":
            args.smoother = "jacobi"      # C4: Gauss-Seidel does not shard

    amg, A, ml, b, t_setup = build_problem(args)
    from algebraicmultigrid_jl_b200 import _devlib

    if world > 1:
        uid = _devlib.nccl_unique_id() if rank == 0 else None
        box = [uid]
        dist.broadcast_object_list(box, src=0)
        ml.partition(rank, world, box[0], levels=args.part_levels)
    t0 = time.time()
    dev = ml.device()
    t_upload = time.time() - t0
    n, nnz = A.n, A.nnz
    K, W = args.steps, max(args.warmup, 0)

    stream = torch.cuda.ExternalStream(dev.stream(), device=torch.device("cuda", local))
    b_d = torch.from_numpy(b).cuda()
    x_d = torch.zeros(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------------------------------
    if W:
        dev.solve(x_d, b_d, 0, W, 0.0, 0.0, True)
    # ---- timed: exactly K `_solve!` iterations, inputs resident in HBM -------------------------------
    x_d.zero_()
    dev.set_option(1, 1)                      # bracket every fine-level convergence residual with events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = dev.launch_count()
    comm0 = dev.comm_stats() if world > 1 else None
    e0.record(stream)
    hist, iters = dev.solve(x_d, b_d, 0, K, 0.0, 0.0, True)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = dev.launch_count() - launches0
    comm1 = dev.comm_stats() if world > 1 else None
    ms = e0.elapsed_time(e1)
    res_ms = dev.residual_timings()
    dev.set_option(1, 0)
    assert iters == K, (iters, K)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = K / (ms * 1e-3)

    # ---- convergence sanity: the timed iterations did real work ---------------------------------------
    x_h = x_d.cpu().numpy()
    err = float(np.abs(x_h - 1.0).max())

    # ---- e2e: the user call `_solve!(x, ml, b; maxiter=K)` with HOST buffers ---------------------------
    xb = torch.zeros(n, dtype=torch.float64).pin_memory()
    bb = torch.from_numpy(b).pin_memory()
    x_np, b_np = xb.numpy(), bb.numpy()
    amg._solve_(x_np, ml, b_np, amg.V(), maxiter=1, reltol=0.0)      # untimed: first-use allocations of the host-vector path
    x_np[:] = 0.0
    barrier()
    t0 = time.perf_counter()
    amg._solve_(x_np, ml, b_np, amg.V(), maxiter=K, reltol=0.0)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": K / e2e_s, "unit": "V-cycles/s", "h2d_bytes_per_step": 16 * n / K, "d2h_bytes_per_step": 8 * n / K,
           "note": f"one `_solve!` call of {K} iterations on pinned host x, b: 2 H2D + 1 D2H vector copies per call, "
                   "amortised over its iterations; wall clock"}
    # the preconditioner use case: every cycle crosses PCIe (ldiv! with host vectors)
    t0 = time.perf_counter()
    reps = max(3, min(K, 10))
    p = amg.aspreconditioner(ml)
    for _ in range(reps):
        amg.ldiv_(x_np, p, b_np)
    e2e["ldiv_host_vectors_cycles_per_s"] = reps / (time.perf_counter() - t0)

    # ---- per-kernel timings; on a partitioned handle the smoother / halo timings are COLLECTIVE: all ranks take part ----
    spmv_ms = dev.time_kernel(0, 0, reps=20)
    jac_or_gs_ms = dev.time_kernel(0, 2, reps=5)
    halo_ms = dev.time_kernel(0, 7, reps=20) if world > 1 else None
    # ---- N > 1: the same (Jacobi) workload unpartitioned on ONE GPU, so strong scaling can be read off this line ----
    n1_same = None
    if world > 1:
        if rank == 0:
            ml1 = amg.MultiLevel(ml.levels, ml.final_A, ml.coarse_solver, None, None, ml.workspace)
            d1 = ml1.device()
            x1 = torch.zeros(n, dtype=torch.float64, device="cuda")
            s1 = torch.cuda.ExternalStream(d1.stream(), device=torch.device("cuda", local))
            d1.solve(x1, b_d, 0, max(W, 1), 0.0, 0.0, True)
            x1.zero_()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            f0.record(s1)
            d1.solve(x1, b_d, 0, K, 0.0, 0.0, True)
            f1.record(s1)
            torch.cuda.synchronize()
            n1_same = {"value": K / (f0.elapsed_time(f1) * 1e-3), "unit": "V-cycles/s",
                       "what": "the same Jacobi hierarchy, fine level NOT partitioned, on rank 0's GPU alone",
                       "parity_max_abs_diff_vs_partitioned": float((x1 - x_d).abs().max().item())}
            ml1.release()
        dist.barrier()
    # ---- N > 1: synthetic 3-D elasticity, SA with the six rigid-body modes, Jacobi(0.5), CG preconditioned by the partitioned
    # V-cycle, everything device-resident (north_star: "throughput on synthetic Poisson / elasticity matrices at 1/2/4/8 GPUs") ----
    elast = None
    if world > 1 and not args.no_other_configs:
        try:
            Ae, be, Be = amg.elasticity_3d(48, 48, 48)
            jac = amg.Jacobi(0.5)
            t0 = time.time()
            mle = amg.smoothed_aggregation(Ae, B=Be, presmoother=jac, postsmoother=jac)
            t_se = time.time() - t0
            box = [_devlib.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            mle.partition(rank, world, box[0], levels=min(2, len(mle.levels)))
            deve = mle.device()
            be_d = torch.from_numpy(be).cuda()
            xe_d = torch.zeros(Ae.n, dtype=torch.float64, device="cuda")
            he, ite = deve.pcg(xe_d, be_d, 0, 300, 0.0, 1e-8)       # warm-up (captures the cycle graph)
            xe_d.zero_()
            barrier()
            t0 = time.perf_counter()
            he, ite = deve.pcg(xe_d, be_d, 0, 300, 0.0, 1e-8)
            torch.cuda.synchronize()
            dte = time.perf_counter() - t0
            tt = torch.tensor([dte], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dte = float(tt.item())
            xe = xe_d.cpu().numpy()
            elast = {"workload": "elasticity_3d(48,48,48): Q1 hexahedra, 3 dofs per node, smoothed_aggregation with six rigid-body modes, pre/post "
                                 "Jacobi(0.5), CG preconditioned by one V-cycle (b200amg_pcg on the row-partitioned handle), reltol 1e-8",
                     "n": Ae.n, "nnz": Ae.nnz, "levels": deve.nlevels, "partitioned_levels": deve.comm_stats()["partitioned_levels"],
                     "setup_s": t_se, "cg_iterations": int(ite), "solve_ms": 1e3 * dte, "cg_iterations_per_s": ite / dte,
                     "relative_residual": float(np.linalg.norm(Ae.matvec(xe) - be) / np.linalg.norm(be))}
            mle.release()
        except Exception as exc:
            elast = {"error": str(exc)[:300]}
    if rank != 0:
        dist.barrier()
        dist.destroy_process_group()
        return
    # ---- the BASELINE metric's kernel: fine-level residual SpMV r = b - A x, timed live inside every timed iteration ----
    peak, peak_src = measured_peak()
    res_ms_avg = float(np.mean(res_ms)) if len(res_ms) else float("nan")
    stor = [dev.storage_info(i) for i in range(dev.nlevels)]
    vb0 = stor[0]["A"] or 8
    alg = bytes_residual(n // world, nnz // world, vb0) if world > 1 else bytes_residual(n, nnz, vb0)
    achieved = alg / (res_ms_avg * 1e-3) / 1e9
    traffic = ncu_traffic(args, world) if vb0 == 8 else None
    spmv_roofline = {"kernel": "csr residual r=b-A*x, fine level (convergence check of every `_solve!` iteration)",
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0, "algorithmic_bytes": alg,
                     "value_bytes": vb0,
                     "value_bytes_note": ("the fine-level values are exactly representable in binary32 and are stored (also) as 4-byte values; "
                                          "the kernel computes the same fp64 products (bit-identical results), so the algorithmic bytes are "
                                          "8 per entry, not SURVEY §8d's 12" if vb0 == 4 else "fp64 values: 12 bytes per entry (SURVEY §8d)"),
                     "fp64_storage_equivalent_gbs": bytes_residual(n // world, nnz // world) / (res_ms_avg * 1e-3) / 1e9,
                     "avg_launch_ms": res_ms_avg, "launches_timed": int(len(res_ms)), "share_of_step": res_ms_avg / (ms / K),
                     "traffic": traffic,
                     "traffic_source": "static: one `ncu --set full` capture of the fp64-storage kernel on this workload, committed as "
                                       "profiles/r01_ncu_traffic.json (not measured in this run; null for the binary32-storage kernel)"}
    phases = ["presmoother", "residual", "restriction", "coarse_solve", "prolongation", "postsmoother"]
    prof = dev.profile_cycle(0) if world == 1 else np.zeros((dev.nlevels, 6))
    infos = [dev.level_info(i) for i in range(dev.nlevels)]
    # ---- `roofline` = the DOMINANT kernel of the step (largest share of the step time), and the whole step ----
    step_alg = cycle_bytes(infos, stor, args.smoother)
    whole_step = {"algorithmic_bytes": step_alg, "achieved": step_alg / (ms / K * 1e-3) / 1e9 * (1 if world == 1 else 1), "unit": "GB/s",
                  "frac": step_alg / (ms / K * 1e-3) / 1e9 / (peak * world), "peak": peak * world,
                  "what": "all kernels of one `_solve!` iteration over the whole hierarchy against the measured HBM peak of the GPUs used"}
    roofline = dict(spmv_roofline)
    if world == 1 and prof.size:
        sm_ms = prof[:, 0] + prof[:, 5]                    # pre + post smoother per level
        other = np.array([prof[:, 1].max(), prof[:, 2].max(), prof[:, 4].max()])
        li = int(np.argmax(sm_ms))
        if sm_ms[li] / 2 > other.max():                    # a smoother sweep dominates (always, for Gauss-Seidel)
            lv = infos[li]
            launches_per_step = 4 if args.smoother == "gs" else 2
            k_ms = float(sm_ms[li]) / launches_per_step
            k_alg = bytes_residual(lv["n"], lv["nnz_a"], 8 if args.smoother == "gs" else (stor[li]["A"] or 8))
            k_ach = k_alg / (k_ms * 1e-3) / 1e9
            what = ("exact-order Gauss-Seidel sweep (one direction; gs! src/smoother.jl:73-90) of level %d: %d rows, %d dependent wavefronts — "
                    "dependency-latency-bound, not bandwidth-bound" % (li, lv["n"], lv["wavefronts"])) if args.smoother == "gs" else \
                   ("Jacobi sweep of level %d (%d rows)" % (li, lv["n"]))
            roofline = {"kernel": what, "bound": "hbm", "achieved": k_ach, "peak": peak, "unit": "GB/s", "frac": k_ach / peak,
                        "peak_source": peak_src, "algorithmic_bytes": k_alg, "avg_launch_ms": k_ms,
                        "launches_per_step": launches_per_step, "share_of_step": float(sm_ms[li]) / (ms / K),
                        "how_timed": "CUDA events around each phase of one cycle replayed eagerly (b200amg_profile_cycle), after the timed region",
                        "traffic": None}
            if args.smoother == "gs" and args.size == 256 and args.dim == 3 and args.method == "rs" and li == 1:
                try:           # static: one ncu --set full capture of this kernel on this level (not measured in this run)
                    with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as f:
                        t = json.load(f)["gs_tile_kernel<4>@poisson256_level1_one_direction"]
                    roofline["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"]
                    roofline["traffic_source"] = "static: profiles/r02_ncu_traffic.json (one `ncu --set full` capture of gs_tile_kernel<4> on level 1 of this workload)"
                except Exception:
                    pass
    roofline["whole_step"] = whole_step
    roofline["fine_level_residual_spmv"] = spmv_roofline
    pinfo = dev.partition_info() if world > 1 else None
    nl, nnzl = (pinfo["row_hi"] - pinfo["row_lo"], nnz // world) if world > 1 else (n, nnz)
    extra = {
        "fine_spmv_ms": spmv_ms, "fine_spmv_gbs": bytes_spmv(nl, nnzl, vb0) / (spmv_ms * 1e-3) / 1e9,
        "fine_spmv_frac_of_measured_peak": bytes_spmv(nl, nnzl, vb0) / (spmv_ms * 1e-3) / 1e9 / peak,
        "fine_spmv_value_bytes": vb0,
        "fine_spmv_note": "rank 0's row block, local kernel only" if world > 1 else "whole fine level",
        "fine_presmoother_ms": jac_or_gs_ms, "halo_exchange_ms": halo_ms, "partition_rank0": pinfo,
        "fine_presmoother_gbs": (2 * bytes_residual(nl, nnzl) if args.smoother == "gs" else bytes_residual(nl, nnzl, vb0)) / (jac_or_gs_ms * 1e-3) / 1e9,
        "phase_ms_per_level": {ph: [round(float(v), 4) for v in prof[:, i]] for i, ph in enumerate(phases)},
        "levels": infos, "value_bytes_per_level": stor, "setup_s": t_setup, "upload_s": t_upload, "max_abs_err_vs_ones": err,
        "residual_history_first_last": [float(hist[0]), float(hist[-1])],
    }
    # ---- CPU baseline beside it: the oracle port on one host core, bounded sample -------------------------
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle

        H = oracle.OracleHierarchy(ml)
        t0 = time.time()
        xo, ho = H.solve(b, maxiter=1, reltol=0.0, log=True)
        per = time.time() - t0
        s = max(1, min(5, int(20.0 / max(per, 1e-9))))
        t0 = time.time()
        xo, h2 = H.solve(b, x0=xo, maxiter=s, reltol=0.0, log=True)
        dt = time.time() - t0
        ho = np.concatenate([ho, h2[1:]])
        cpu = {"value": s / dt, "unit": "V-cycles/s", "cores": 1, "kind": "port",
               "sample": f"{s} `_solve!` iterations of the same hierarchy after 1 warm-up iteration, single thread "
                         f"(reference solve phase is single-threaded); host has {os.cpu_count()} logical cores"}
        # ---- parity at FULL size: the same 1 + s iterations from x0 = 0 on the device, against the oracle's ----
        xg = torch.zeros(n, dtype=torch.float64, device="cuda")
        hg, itg = dev.solve(xg, b_d, 0, 1 + s, 0.0, 0.0, True)
        xg = xg.cpu().numpy()
        parity = {"iters": int(itg), "oracle_iters": int(len(ho) - 1),
                  "rel_hist_diff": float(np.max(np.abs(hg - ho) / np.abs(ho))) if len(hg) == len(ho) else None,
                  "rel_x_diff": float(np.abs(xg - xo).max() / np.abs(xo).max()),
                  "tolerance": {"rel_hist_diff": 1e-6, "rel_x_diff": 1e-9},
                  "what": f"`_solve!` from x0 = 0, {1 + s} iterations, device vs CPU oracle on the full {workload_name(args)}"}
        parity["ok"] = bool(parity["rel_hist_diff"] is not None and parity["rel_hist_diff"] <= 1e-6 and parity["rel_x_diff"] <= 1e-9)
        # courtesy figure, NOT reference behaviour (its solve phase is single-threaded): the headline kernel r = b - A x on all
        # host cores (OpenMP over rows, host setup library), i.e. what the host's memory system can do on the same bytes
        try:
            from algebraicmultigrid_jl_b200 import _hostlib

            _, sec = _hostlib.residual_allcores(A, xo, b, reps=5)
            cpu["fine_residual_all_cores"] = {"ms": 1e3 * sec, "GBs": bytes_residual(n, nnz) / sec / 1e9, "threads": os.cpu_count(),
                                              "note": "OpenMP row-parallel residual on all host cores; not reference behaviour"}
        except Exception as exc:  # never let a courtesy figure break the bench line
            cpu["fine_residual_all_cores"] = {"error": str(exc)[:200]}
    hier_gb = 12e-9 * sum(i["nnz_a"] for i in infos)
    l2_note = (f"fine-level A is {12e-9 * nnz:.2f} GB, the hierarchy {hier_gb:.2f} GB, L2 is 0.126 GB: "
               + ("every kernel of a cycle streams a different operator several times the size of L2, nothing survives from one "
                  "iteration to the next; no flush" if 12e-9 * nnz > 4 * 0.126 else
                  "this workload is SMALL against L2 (it stays resident): per-kernel rates in this line are L2 rates, not HBM rates"))
    roofline["parity"] = parity if cpu is not None else None
    line = {
        "metric": "V-cycle iterations/s", "value": value, "unit": "V-cycles/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, A, dev.nlevels),
        "l2": l2_note,
        "parallelism": (f"{args.part_levels} finest level(s) row-partitioned x{world}, the rest on rank 0" if world > 1 else "single GPU"),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "parity": parity if cpu is not None else None,
    }
    if world > 1:
        line["n1_same_workload"] = n1_same
        eff = (value / n1_same["value"] / world) if n1_same else None
        line["strong_efficiency_vs_n1_same_workload"] = eff
        # kept inside `roofline` as well: the driver's record keeps that object whole
        roofline["n1_same_workload_value"] = n1_same["value"] if n1_same else None
        roofline["strong_efficiency"] = eff
        roofline["parity_max_abs_diff_vs_partitioned"] = n1_same["parity_max_abs_diff_vs_partitioned"] if n1_same else None
        line["note"] = ("BASELINE config C4: Jacobi smoother (Gauss-Seidel, the N=1 headline config C3, is sequential over the "
                        "index range and does not shard); the finest --part-levels levels are split by rows, the levels below them run "
                        "on rank 0, which bounds the whole-cycle speed-up (Amdahl); strong-scaling baseline = n1_same_workload")
        line["comm"] = {"halo_exchange": "peer memory: every rank maps its neighbours' vectors (CUDA IPC) and stores its boundary entries straight "
                                         "into their halos over NVLink, flag + ack words instead of a rendezvous (csrc/device/peer_halo.cuh)"
                                         if comm1["peer_halo"] else "NCCL grouped send/recv",
                        "peer_halo_exchanges_per_step": (comm1["peer_exchanges"] - comm0["peer_exchanges"]) / K,
                        "nccl_groups_per_step": (comm1["nccl_groups"] - comm0["nccl_groups"]) / K,
                        "nccl_groups_what": "coarse_b rows -> rank 0, coarse_x windows <- rank 0, one all-reduce of the squared residual norm",
                        "partitioned_levels": comm1["partitioned_levels"], "one_halo_exchange_ms": halo_ms}
        roofline["comm"] = line["comm"]
    if world == 1 and not args.no_other_configs:
        line["other_configs"] = other_configs(amg, torch, local)
    if world > 1 and elast is not None:
        line["other_configs"] = {"elasticity_3d": elast}
    line.update(extra)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--method", default="rs", choices=["rs", "sa"])
    ap.add_argument("--smoother", default="gs", choices=["gs", "jacobi"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the quick C2 / C5 numbers")
    ap.add_argument("--part-levels", type=int, default=int(os.environ.get("B200AMG_PART_LEVELS", "3")),
                    help="N > 1: how many of the finest levels are split by rows over the ranks")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
