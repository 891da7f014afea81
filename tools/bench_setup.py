#!/usr/bin/env python
"""Setup time of ruge_stuben with the Galerkin products on the host (OpenMP) and on the device (b200amg_spgemm_*).
Usage: python tools/bench_setup.py [--size 128]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import algebraicmultigrid_jl_b200 as amg  # noqa: E402
from algebraicmultigrid_jl_b200 import _devlib, _hostlib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=128)
args = ap.parse_args()
A = amg.poisson((args.size,) * 3)
out = {}
for backend in ("host", "device", "host", "device"):
    amg.set_galerkin_backend(backend)
    t0 = time.time()
    ml = amg.ruge_stuben(A)
    out.setdefault(backend, []).append(round(time.time() - t0, 3))
amg.set_galerkin_backend("host")
lv = ml.levels[0]
R = lv.R.materialize() if hasattr(lv.R, "materialize") else lv.R
P = lv.P.materialize() if hasattr(lv.P, "materialize") else lv.P
for name, fn in (("host", _hostlib.spgemm), ("device", _devlib.spgemm)):
    t0 = time.time()
    RA = fn(R, lv.A)
    t1 = time.time()
    RAP = fn(RA, P)
    t2 = time.time()
    out[name + "_fine_level_products_s"] = [round(t1 - t0, 3), round(t2 - t1, 3)]
print({"size": args.size, "n": A.n, "ruge_stuben_setup_s": out, "cores": os.cpu_count()})
