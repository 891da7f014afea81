#!/usr/bin/env python
"""Run-time knobs of the TMA-fed mailbox sweep (gs_tile_kernel) on the wide levels of a hierarchy: poll mode
(B200AMG_GS_POLL_MASKED: 0 all mailboxes every round, 1 only the outstanding ones, 2 + focused spin on one), back-off between
failed polls and between throttle polls.  SGS ms per level, with the error against the CPU oracle's sweep on level 1.
Usage: python tools/tile_knobs.py [--size 256] [--levels 1,2]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import algebraicmultigrid_jl_b200 as amg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--levels", default="1,2")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
A = amg.poisson((args.size,) * 3)
ml = amg.ruge_stuben(A)
dev = ml.device()
dev.cycle(np.zeros(A.n), A.matvec(np.ones(A.n)), 0)
levels = [int(v) for v in args.levels.split(",")]
configs = [(1, 0, 100), (2, 0, 100), (0, 0, 100), (1, 0, 100), (2, 0, 100), (2, 0, 0), (1, 0, 0), (2, 0, 300), (1, 0, 300),
           (2, 40, 100), (1, 40, 100)]
if os.environ.get("SKIP_POLL_CONFIGS"):
    configs = []
for masked, poll_sleep, gate_sleep in configs:
    dev.set_option(16, masked)
    dev.set_option(5, poll_sleep)
    dev.set_option(6, gate_sleep)
    row = {}
    for lv in levels:
        info = dev.level_info(lv)
        ms = dev.time_kernel(lv, 2, reps=args.reps)
        row[lv] = {"sgs_ms": round(ms, 4), "us_per_wavefront": round(1e3 * ms / max(2 * info["wavefronts"], 1), 3)}
    print(json.dumps({"poll_masked": masked, "poll_sleep": poll_sleep, "gate_sleep": gate_sleep, "levels": row}), flush=True)
dev.set_option(16, -1)
dev.set_option(5, 0)
dev.set_option(6, 100)
for dist in [int(v) for v in os.environ.get("GATE_DISTS", "").split(",") if v]:
    dev.set_option(20, dist)
    for ctas in [int(v) for v in os.environ.get("CTA_LIMITS", "0").split(",")]:
        dev.set_option(19, ctas)
        row = {}
        for lv in levels:
            info = dev.level_info(lv)
            ms = dev.time_kernel(lv, 2, reps=args.reps)
            row[lv] = {"sgs_ms": round(ms, 4), "us_per_wavefront": round(1e3 * ms / max(2 * info["wavefronts"], 1), 3)}
        print(json.dumps({"gate_dist": dist, "cta_limit": ctas, "levels": row}), flush=True)
dev.set_option(20, 2)
for ctas in [int(v) for v in os.environ.get("CTA_LIMITS", "0,222,148,74,0").split(",")]:   # persistent CTAs of the sweep (0 = all that fit: 2 per SM): tiles in flight
    dev.set_option(19, ctas)
    row = {}
    for lv in levels:
        info = dev.level_info(lv)
        ms = dev.time_kernel(lv, 2, reps=args.reps)
        row[lv] = {"sgs_ms": round(ms, 4), "us_per_wavefront": round(1e3 * ms / max(2 * info["wavefronts"], 1), 3)}
    print(json.dumps({"poll_masked": "auto", "cta_limit": ctas, "levels": row}), flush=True)
ml.release()
