#!/usr/bin/env python
"""Where does a wavefront of the distributed-shared-memory Gauss-Seidel sweep (csrc/device/dsm_gs.cuh) go?
Runs one instrumented forward sweep per eligible level and prints the medians of the per-tile stage durations
(SM cycles of thread 0 of the owning CTA) and the time between consecutive wavefront completions (globaltimer).
Usage: B200AMG_GS_DSM=1 python tools/dsm_timeline.py [--size 128] [--levels 2,3,4]"""
import argparse
import os
import sys

os.environ.setdefault("B200AMG_GS_DSM", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import algebraicmultigrid_jl_b200 as amg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--levels", default="")
args = ap.parse_args()
A = amg.poisson((args.size,) * 3)
ml = amg.ruge_stuben(A)
dev = ml.device()
b = A.matvec(np.ones(A.n))
x = np.zeros(A.n)
dev.cycle(x, b, 0)
levels = [int(v) for v in args.levels.split(",")] if args.levels else range(dev.nlevels - 1)
names = ["top -> stage + row registers", "-> previous wavefront complete", "-> gathers landed", "-> sums + lane reduction",
         "-> divide, store, bar.sync", "-> fence + arrivals issued"]
for lv in levels:
    info = dev.level_info(lv)
    if info["n"] / max(info["wavefronts"], 1) >= 1024:
        continue
    try:
        for rep in range(2):
            t = dev.gs_timeline(lv)
    except Exception as e:  # level without a dsm plan
        print(f"level {lv}: {e}")
        continue
    t = t[t[:, 7] > 0]
    span = (t[:, 7].max() - t[:, 7].min()) / 1e3
    print(f"level {lv}: n={info['n']} wavefronts={info['wavefronts']} tiles={len(t)} sweep={span:.1f} us "
          f"({span / max(info['wavefronts'], 1):.3f} us/wavefront)")
    for k, nm in enumerate(names):
        col = t[:, k + 1] - t[:, k]
        print(f"    {nm:34s} median {np.median(col):7.0f} cycles   p90 {np.percentile(col, 90):7.0f}   max {col.max():8.0f}")
    col = t[:, 6] - t[:, 2]
    print(f"    {'wavefront complete -> my arrivals':34s} median {np.median(col):7.0f} cycles   p90 {np.percentile(col, 90):7.0f}")
