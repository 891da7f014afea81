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
names = ["claim->prefetch issued", "prefetch->wait done", "wait done->fence+sync", "gather+compute+store", "cta barrier",
         "release fence", "red issue"]
for lv in range(dev.nlevels - 1):
    for rep in range(2):
        t = dev.gs_timeline(lv)
    info = dev.level_info(lv)
    d = np.diff(t, axis=1)
    total = (t[:, 7].max() - t[:, 0].min()) / 1e3
    print(f"level {lv}: n={info['n']} wavefronts={info['wavefronts']} tasks={len(t)} sweep={total:.1f} us "
          f"({total / info['wavefronts']:.2f} us/wavefront)")
    for k, nm in enumerate(names):
        col = d[:, k]
        if k in (1, 2) :
            col = col[t[:, 2] > 0] if k == 1 else col[t[:, 2] > 0]
        if len(col):
            print(f"    {nm:26s} median {np.median(col):8.0f} ns   p90 {np.percentile(col, 90):8.0f}   max {col.max():8.0f}")
    # publication-to-detection latency: a task's wait-done stamp minus the LAST publication stamp overall before it
    pub = np.sort(t[:, 7])
    wd = t[t[:, 2] > 0, 2]
    idx = np.searchsorted(pub, wd, side="right") - 1
    lat = wd - pub[np.maximum(idx, 0)]
    print(f"    last publish -> wait done  median {np.median(lat):8.0f} ns   p90 {np.percentile(lat, 90):8.0f}")
