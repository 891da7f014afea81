#!/usr/bin/env python
"""Blocked Gauss-Seidel sweep (block_gs.cuh) level by level against the CPU oracle, with per-level SGS timings.

    python tools/block_smoke.py --sizes 16,40,96 [--sor]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="16,40,96")
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--sor", action="store_true")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import algebraicmultigrid_jl_b200 as amg
    import oracle

    for size in [int(v) for v in args.sizes.split(",")]:
        A = amg.poisson((size,) * args.dim)
        sm = amg.SOR(1.2) if args.sor else amg.GaussSeidel()
        t0 = time.time()
        ml = amg.ruge_stuben(A, presmoother=sm, postsmoother=sm)
        t1 = time.time()
        dev = ml.device()
        print(f"size {size}^{args.dim}: setup {t1 - t0:.1f} s upload {time.time() - t1:.1f} s", flush=True)
        for li, lv in enumerate(ml.levels):
            M = lv.A
            rng = np.random.default_rng(li)
            x = rng.random(M.n)
            b = rng.random(M.n)
            xd = dev.smooth(li, 0, x.copy(), b)
            err = float("nan")
            if not args.no_oracle:
                xr = oracle.smooth(M, sm, x.copy(), b)
                err = np.abs(xd - xr).max() / np.abs(xr).max()
            ms = dev.time_kernel(li, 2, reps=args.reps)
            info = dev.level_info(li)
            wf = max(info["wavefronts"], 1)
            gbs = 2 * (12 * info["nnz_a"] + 28 * info["n"]) / (ms * 1e-3) / 1e9
            print(f"  level {li}: n={M.n:9d} nnz/row={M.nnz / M.n:5.1f} wavefronts={wf:5d}  rel.err {err:.2e}  SGS {ms:8.4f} ms "
                  f"= {1e3 * ms / (2 * wf):6.3f} us/wavefront  {gbs:7.1f} GB/s", flush=True)
            if not args.no_oracle and not (err < 1e-12):
                raise SystemExit(f"PARITY FAILURE on level {li}")
        ml.release()


if __name__ == "__main__":
    main()
