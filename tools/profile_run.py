#!/usr/bin/env python
"""Small driver for ncu captures: builds one hierarchy and launches each hot-path kernel a few
times outside CUDA graphs.  python tools/profile_run.py --size 256 --smoother gs --what 0,1,2 --levels 0,1"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import algebraicmultigrid_jl_b200 as amg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--smoother", default="gs")
ap.add_argument("--what", default="0,1,2,3,4")
ap.add_argument("--levels", default="0")
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
A = amg.poisson((args.size,) * 3)
kw = {}
if args.smoother == "jacobi":
    sm = amg.Jacobi(2.0 / 3.0)
    kw = dict(presmoother=sm, postsmoother=sm)
ml = amg.ruge_stuben(A, **kw)
dev = ml.device()
dev.set_option(0, 0)          # no graphs: every kernel is an ordinary launch
b = A.matvec(np.ones(A.n))
x = np.zeros(A.n)
dev.cycle(x, b, 0)
for lv in [int(v) for v in args.levels.split(",")]:
    if lv >= dev.nlevels - 1:
        continue
    for what in [int(v) for v in args.what.split(",")]:
        ms = dev.time_kernel(lv, what, reps=args.reps)
        print(f"level {lv} what {what}: {ms:.4f} ms", flush=True)
