#!/usr/bin/env python
"""A/B sweep of the distributed-shared-memory Gauss-Seidel kernel's upload-time knobs on the coarse part of a
hierarchy: builds the hierarchy once, then uploads the levels from --from-level down once per configuration
(B200AMG_DSM_THREADS / B200AMG_DSM_LANES / B200AMG_GS_DSM_LOG_NC are read at upload) and times one symmetric sweep per level.
Usage: python tools/dsm_sweep.py [--size 256] [--from-level 3] [--configs "256,0,-1;512,0,-1;512,16,-1"]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import algebraicmultigrid_jl_b200 as amg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--from-level", type=int, default=3)
ap.add_argument("--configs", default="256,0,-1;512,0,-1;512,16,-1;256,16,-1")
ap.add_argument("--baseline", action="store_true")
args = ap.parse_args()
A = amg.poisson((args.size,) * 3)
ml = amg.ruge_stuben(A)
sub = amg.MultiLevel(ml.levels[args.from_level:], ml.final_A, ml.coarse_solver, None, None, ml.workspace)
n0 = sub.levels[0].A.n
b = np.random.default_rng(0).random(n0)


def run(tag, dsm):
    dev = sub.device()
    x = np.zeros(n0)
    dev.cycle(x, b, 0)
    dev.set_option(13, dsm)
    dev.set_option(15, 4)
    row = {}
    for lv in range(dev.nlevels - 1):
        info = dev.level_info(lv)
        ms = dev.time_kernel(lv, 2, reps=5)
        row[lv + args.from_level] = {"n": info["n"], "sgs_ms": round(ms, 4), "us_per_wavefront": round(1e3 * ms / max(2 * info["wavefronts"], 1), 3)}
    print(json.dumps({"config": tag, "levels": row}), flush=True)
    sub.release()


if args.baseline:
    run("baseline (no dsm)", 0)
for cfg in args.configs.split(";"):
    threads, lanes, lognc = [int(v) for v in cfg.split(",")]
    os.environ["B200AMG_DSM_THREADS"] = str(threads)
    if lanes > 0:
        os.environ["B200AMG_DSM_LANES"] = str(lanes)
    else:
        os.environ.pop("B200AMG_DSM_LANES", None)
    if lognc >= 0:
        os.environ["B200AMG_GS_DSM_LOG_NC"] = str(lognc)
    else:
        os.environ.pop("B200AMG_GS_DSM_LOG_NC", None)
    run({"threads": threads, "lanes": lanes or "auto", "log_nc": lognc if lognc >= 0 else "auto"}, 1)
