#!/usr/bin/env python
"""A/B of the one-cluster Gauss-Seidel sweeps on the coarse part of a hierarchy: gs_dsm_kernel (one consumer group) against
gs_dsm2_kernel (two groups alternating the tiles; B200AMG_OPT_GS_DSM2), per level, symmetric sweep, with parity against the
CPU oracle's sweep.  Usage: python tools/dsm2_ab.py [--size 256] [--from-level 3] [--max-log-nc 2,4]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import algebraicmultigrid_jl_b200 as amg  # noqa: E402
import oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--from-level", type=int, default=3)
ap.add_argument("--max-log-nc", default="2,4")
ap.add_argument("--threads", default="256")
args = ap.parse_args()
A = amg.poisson((args.size,) * 3)
ml = amg.ruge_stuben(A)
sub = amg.MultiLevel(ml.levels[args.from_level:], ml.final_A, ml.coarse_solver, None, None, ml.workspace)
n0 = sub.levels[0].A.n
r = np.random.default_rng(0)
os.environ["B200AMG_GS_BLOCK"] = "0"
dev = sub.device()
dev.cycle(np.zeros(n0), r.random(n0), 0)
refs = {}
for lv, level in enumerate(sub.levels):
    x0, b = r.standard_normal(level.A.n), r.standard_normal(level.A.n)
    refs[lv] = (x0, b, oracle.smooth(level.A, level.presmoother.config, x0.copy(), b))
for mx in [int(v) for v in args.max_log_nc.split(",")]:
    dev.set_option(15, mx)
    for dsm2 in (0, 1, 0, 1):
        dev.set_option(18, dsm2)
        row = {}
        for lv in range(dev.nlevels - 1):
            info = dev.level_info(lv)
            x0, b, ref = refs[lv]
            x = dev.smooth(lv, 0, x0.copy(), b)
            err = float(np.abs(x - ref).max() / np.abs(ref).max())
            ms = dev.time_kernel(lv, 2, reps=5)
            row[lv + args.from_level] = {"n": info["n"], "wavefronts": info["wavefronts"], "sgs_ms": round(ms, 4),
                                         "us_per_wavefront": round(1e3 * ms / max(2 * info["wavefronts"], 1), 3), "rel_err": err}
        print(json.dumps({"max_log_nc": mx, "dsm2": dsm2, "levels": row}), flush=True)
sub.release()
