#!/usr/bin/env python
"""Where does a dataflow Gauss-Seidel sweep spend its time?  Runs one instrumented forward sweep per
level and prints, per level, the medians of the per-task stage durations and the per-wavefront hop."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import algebraicmultigrid_jl_b200 as amg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=128)
args = ap.parse_args()
A = amg.poisson((args.size,) * 3)
ml = amg.ruge_stuben(A)
dev = ml.device()
b = A.matvec(np.ones(A.n))
x = np.zeros(A.n)
dev.cycle(x, b, 0)
for lv in range(dev.nlevels - 1):
    for rep in range(2):
        t = dev.gs_timeline(lv)
    info = dev.level_info(lv)
    total = (t[:, 7].max() - t[:, 0].min()) / 1e3
    print(f"level {lv}: n={info['n']} wavefronts={info['wavefronts']} tasks={len(t)} sweep={total:.1f} us "
          f"({total / info['wavefronts']:.2f} us/wavefront)")
    # stamps: 0 claimed, 1 earlier-x burst landed, 2 wait done, 3 after acquire+bar, 4 row stored, 5 bar passed, 6 fence done, 7 published
    w = t[:, 2] > 0
    segs = [("claim -> wait done", t[w, 2] - t[w, 0]), ("wait done -> bar.sync passed", t[w, 3] - t[w, 2]),
            ("gather burst (L2 round trip)", t[:, 1] - t[:, 3]), ("compute + store", t[:, 4] - t[:, 1]),
            ("cta barrier", t[:, 5] - t[:, 4]), ("release fence", t[:, 6] - t[:, 5]), ("red issue", t[:, 7] - t[:, 6])]
    for nm, col in segs:
        if len(col):
            print(f"    {nm:30s} median {np.median(col):8.0f} ns   p90 {np.percentile(col, 90):8.0f}   max {col.max():8.0f}")
    # publication-to-detection latency: a task's wait-done stamp minus the LAST publication stamp overall before it
    pub = np.sort(t[:, 7])
    wd = t[t[:, 2] > 0, 2]
    idx = np.searchsorted(pub, wd, side="right") - 1
    lat = wd - pub[np.maximum(idx, 0)]
    print(f"    last publish -> wait done  median {np.median(lat):8.0f} ns   p90 {np.percentile(lat, 90):8.0f}")
