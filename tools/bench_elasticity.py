#!/usr/bin/env python
"""Synthetic 2-D linear elasticity (config C5 at scale): smoothed aggregation with the rigid-body near-null-space as a
preconditioner for CG, everything on the device.  Prints one JSON line.
    python tools/bench_elasticity.py [--size 1024] [--reltol 1e-8]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import algebraicmultigrid_jl_b200 as amg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--reltol", type=float, default=1e-8)
args = ap.parse_args()
t0 = time.time()
A, b, B = amg.elasticity_2d(args.size, args.size)
t_gen = time.time() - t0
t0 = time.time()
ml = amg.smoothed_aggregation(A, B=B)
t_setup = time.time() - t0
t0 = time.time()
dev = ml.device()
t_upload = time.time() - t0
p = amg.aspreconditioner(ml)
x, info = amg.cg(A, b, Pl=p, reltol=args.reltol, log=True)      # warm-up (captures the cycle graph)
t0 = time.time()
x, info = amg.cg(A, b, Pl=p, reltol=args.reltol, log=True)
t_pcg = time.time() - t0
rel = float(np.linalg.norm(A.matvec(x) - b) / np.linalg.norm(b))
t0 = time.time()
xs, hist = amg._solve(ml, b, log=True, reltol=args.reltol, maxiter=200)
t_solve = time.time() - t0
n, nnz = A.n, A.nnz
spmv_ms = dev.time_kernel(0, 0, reps=20, flush_l2=True)
cyc_ms = dev.time_kernel(0, 5, reps=10)
out = {"problem": f"elasticity_2d({args.size},{args.size}) Q1 plane strain, clamped edge", "n": n, "nnz": nnz, "levels": dev.nlevels,
       "gen_s": t_gen, "setup_s": t_setup, "upload_s": t_upload,
       "pcg": {"iters": info["iters"], "seconds_incl_pcie": t_pcg, "iters_per_s": info["iters"] / t_pcg, "true_rel_residual": rel},
       "standalone_solve": {"iters": len(hist) - 1, "seconds_incl_pcie": t_solve, "converged": bool(hist[-1] <= args.reltol * hist[0])},
       "v_cycle_ms": cyc_ms, "fine_spmv": {"ms": spmv_ms, "GBs": (12 * nnz + 4 * (n + 1) + 16 * n) / spmv_ms / 1e6, "l2_flushed": True}}
# (the CPU restatement is test infrastructure: only tests/, smoke() and bench.py's CPU legs may run it — no CPU leg here)
print(json.dumps(out), flush=True)
