#!/usr/bin/env python
"""Per-kernel timing sweep on one GPU (CUDA events inside b200amg_time_kernel): fine and coarse
level SpMV / residual / restriction / prolongation / pre-smoother, for several STREAM_CHUNK values.
Usage: python tools/tune_kernels.py [--size 256] [--smoother gs|jacobi] [--chunks 1,2,4,0]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import algebraicmultigrid_jl_b200 as amg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--smoother", default="jacobi")
ap.add_argument("--chunks", default="1,2,4,16,0")
ap.add_argument("--levels", type=int, default=3)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
n1 = args.size
t0 = time.time()
A = amg.poisson((n1, n1, n1))
kw = {}
if args.smoother == "jacobi":
    sm = amg.Jacobi(2.0 / 3.0)
    kw = dict(presmoother=sm, postsmoother=sm)
ml = amg.ruge_stuben(A, **kw)
dev = ml.device()
print(f"setup+upload {time.time() - t0:.1f}s", flush=True)
b = A.matvec(np.ones(A.n))
x = np.zeros(A.n)
dev.cycle(x, b, 0)     # fill the level vectors with real data
names = {0: "spmv", 1: "residual", 2: "presmooth", 3: "restrict", 4: "prolong"}
for chunk in [int(c) for c in args.chunks.split(",")]:
    dev.set_option(2, chunk)
    for lv in range(min(args.levels, dev.nlevels - 1)):
        info = dev.level_info(lv)
        n, nnz, nnzp = info["n"], info["nnz_a"], info["nnz_p"]
        nc = dev.level_info(lv + 1)["n"]
        alg = {0: 12 * nnz + 4 * (n + 1) + 16 * n, 1: 12 * nnz + 4 * (n + 1) + 24 * n,
               2: (12 * nnz + 4 * (n + 1) + 24 * n) * (2 if args.smoother == "gs" else 1),
               3: 12 * nnzp + 4 * (nc + 1) + 8 * n + 8 * nc, 4: 12 * nnzp + 4 * (n + 1) + 8 * nc + 16 * n}
        out = {"chunk": chunk, "level": lv, "n": n, "nnz": nnz}
        for what in range(5):
            ms = dev.time_kernel(lv, what, reps=args.reps, flush_l2=(alg[what] < 512e6))
            out[names[what]] = {"ms": round(ms, 4), "GBs": round(alg[what] / ms / 1e6, 1)}
        print(json.dumps(out), flush=True)
if args.smoother == "gs":
    for mode in [int(m) for m in os.environ.get("GS_MODES", "2,3,1").split(",")]:
        dev.set_option(3, mode)
        for lv in range(dev.nlevels - 1):
            info = dev.level_info(lv)
            n, nnz = info["n"], info["nnz_a"]
            ms = dev.time_kernel(lv, 2, reps=5)
            alg = 2 * (12 * nnz + 4 * (n + 1) + 24 * n)
            print(json.dumps({"gs_mode": mode, "level": lv, "n": n, "nnz": nnz, "wavefronts": info["wavefronts"],
                              "sgs_ms": round(ms, 4), "GBs": round(alg / ms / 1e6, 1),
                              "us_per_wavefront": round(1e3 * ms / max(2 * info["wavefronts"], 1), 3)}), flush=True)
if args.smoother == "gs" and os.environ.get("SWEEP_MASKED"):
    dev.set_option(3, 2)
    for masked in (1, 0, 1, 0):
        dev.set_option(16, masked)
        row = {}
        for lv in range(min(3, dev.nlevels - 1)):
            info = dev.level_info(lv)
            ms = dev.time_kernel(lv, 2, reps=5)
            row[lv] = {"sgs_ms": round(ms, 4), "us_per_wavefront": round(1e3 * ms / max(2 * info["wavefronts"], 1), 3)}
        print(json.dumps({"poll_masked": masked, "levels": row}), flush=True)
    dev.set_option(16, 1)
if args.smoother == "gs" and not os.environ.get("SKIP_DSM"):
    # one-cluster sweep with x in distributed shared memory (cluster size fixed at upload: B200AMG_GS_DSM_LOG_NC)
    dev.set_option(3, 2)
    for fence in (0, 4, 1):
        dev.set_option(13, 1)
        dev.set_option(15, 4)
        dev.set_option(14, fence)
        for lv in range(dev.nlevels - 1):
            info = dev.level_info(lv)
            n, nnz = info["n"], info["nnz_a"]
            ms = dev.time_kernel(lv, 2, reps=5)
            print(json.dumps({"gs_dsm": 1, "fence": fence, "log_nc_env": os.environ.get("B200AMG_GS_DSM_LOG_NC"), "level": lv, "n": n,
                              "nnz": nnz, "wavefronts": info["wavefronts"], "sgs_ms": round(ms, 4),
                              "us_per_wavefront": round(1e3 * ms / max(2 * info["wavefronts"], 1), 3)}), flush=True)
    dev.set_option(13, 0)
if args.smoother == "gs" and os.environ.get("SWEEP_SLEEP"):
    dev.set_option(3, 2)
    for poll, gate in ((0, 100), (0, 0), (40, 100), (0, 500), (100, 1000)):
        dev.set_option(5, poll)
        dev.set_option(6, gate)
        row = []
        for lv in range(dev.nlevels - 1):
            info = dev.level_info(lv)
            ms = dev.time_kernel(lv, 2, reps=3)
            row.append(round(1e3 * ms / max(2 * info["wavefronts"], 1), 2))
        print(json.dumps({"poll_sleep": poll, "gate_sleep": gate, "us_per_wavefront_by_level": row}), flush=True)
if args.smoother == "gs" and os.environ.get("SWEEP_CTA"):
    for cta_rows in (65536, 0):
        dev.set_option(3, 2)
        dev.set_option(7, cta_rows)
        row = []
        for lv in range(dev.nlevels - 1):
            info = dev.level_info(lv)
            ms = dev.time_kernel(lv, 2, reps=3)
            row.append(round(1e3 * ms / max(2 * info["wavefronts"], 1), 2))
        print(json.dumps({"cta_rows": cta_rows, "us_per_wavefront_by_level": row}), flush=True)
if args.smoother == "gs" and os.environ.get("SWEEP_CLUSTER"):
    dev.set_option(3, 2)
    for lognc, bs in ((1, 1024), (2, 1024), (3, 1024), (4, 1024), (2, 256), (3, 256), (4, 256)):
        dev.set_option(10, lognc)
        dev.set_option(11, bs)
        row = {}
        for lv in (3, 4):
            info = dev.level_info(lv)
            ms = dev.time_kernel(lv, 2, reps=3)
            row[lv] = round(1e3 * ms / max(2 * info["wavefronts"], 1), 2)
        print(json.dumps({"cluster_ctas": 1 << lognc, "threads": bs, "us_per_wavefront": row}), flush=True)
