#!/usr/bin/env python
"""Where does a wavefront of the TMA-fed mailbox Gauss-Seidel sweep (gs_tile_kernel, wide levels) go?
One instrumented forward sweep per wide level; stamps are thread 0's view of every tile (globaltimer, ns):
0 tile begins, 1 tile data in shared memory, 2 throttle gate passed, 3 all earlier-ordered neighbours seen,
4 row published, 5 tile closed (bar.sync), 6 poll rounds, 7 wavefront.
Usage: python tools/tile_timeline.py [--size 256] [--levels 0,1,2]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import algebraicmultigrid_jl_b200 as amg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--levels", default="0,1,2")
args = ap.parse_args()
A = amg.poisson((args.size,) * 3)
ml = amg.ruge_stuben(A)
dev = ml.device()
b = A.matvec(np.ones(A.n))
x = np.zeros(A.n)
dev.cycle(x, b, 0)
for lv in [int(v) for v in args.levels.split(",")]:
    info = dev.level_info(lv)
    if lv >= dev.nlevels - 1 or info["n"] / max(info["wavefronts"], 1) < 1024:
        continue
    for rep in range(2):
        t = dev.gs_timeline(lv)
    t = t[t[:, 4] > 0]
    w = t[:, 7]
    nw = int(w.max()) + 1
    done = np.zeros(nw)
    first = np.full(nw, np.inf)
    np.maximum.at(done, w, t[:, 4])
    np.minimum.at(first, w, t[:, 4])
    span = (t[:, 5].max() - t[:, 0].min()) / 1e3
    hop = np.diff(done)
    print(f"level {lv}: n={info['n']} wavefronts={info['wavefronts']} tiles={len(t)} sweep={span:.1f} us ({span / nw:.3f} us/wavefront); "
          f"hop between wavefront completions: median {np.median(hop):.0f} ns p90 {np.percentile(hop, 90):.0f}")
    ok = w >= 1
    prev_done, prev_first = done[np.maximum(w - 1, 0)], first[np.maximum(w - 1, 0)]
    rows = [("tile begins -> data in smem", t[:, 1] - t[:, 0]), ("-> gate passed", t[:, 2] - t[:, 1]), ("-> neighbours seen (poll loop)", t[:, 3] - t[:, 2]),
            ("-> row published (compute)", t[:, 4] - t[:, 3]), ("-> tile closed (bar.sync)", t[:, 5] - t[:, 4]), ("poll rounds", t[:, 6]),
            ("gate passed - previous wavefront's LAST publish", (t[:, 2] - prev_done)[ok]),
            ("neighbours seen - previous wavefront's LAST publish", (t[:, 3] - prev_done)[ok]),
            ("neighbours seen - previous wavefront's FIRST publish", (t[:, 3] - prev_first)[ok]),
            ("wavefront spread (last - first publish)", done - first)]
    for nm, col in rows:
        print(f"    {nm:52s} median {np.median(col):8.0f}   p10 {np.percentile(col, 10):8.0f}   p90 {np.percentile(col, 90):8.0f}   max {col.max():8.0f}")
