#!/usr/bin/env python
"""Where the cycles of the pass sweep (pass_gs.cuh) go: compute thread 0 of every tile (group 0).

    B200AMG_GS_PASS=1 B200AMG_GS_BLOCK=2 python tools/pass_timeline.py --size 128 [--levels 0,3]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--levels", default="")
    args = ap.parse_args()
    import algebraicmultigrid_jl_b200 as amg

    A = amg.poisson((args.size,) * 3)
    ml = amg.ruge_stuben(A)
    dev = ml.device()
    b = np.random.default_rng(0).random(A.n)
    dev.solve(np.zeros(A.n), b, 0, 1, 0.0, 0.0, True)      # fills the level vectors
    want = [int(v) for v in args.levels.split(",")] if args.levels else range(len(ml.levels))
    for li in want:
        for bwd in (False, True):
            t = dev.gs_timeline(li, bwd)
            tot, wait, relax, prep, _, own, npass = [t[:, q].astype(float) for q in range(7)]
            ns = t[:, 7]
            print(f"level {li} {'bwd' if bwd else 'fwd'}: tiles {len(t)} sweep {(ns.max() - ns.min()) * 1e-3:9.1f} us (first->last tile end) | per tile: "
                  f"passes {npass.mean():7.1f} cycles {tot.mean():9.0f} | per OWN pass of group 0 (= 2 passes of the tile): barrier wait {wait.sum() / own.sum():6.0f} "
                  f"relax+arrive {relax.sum() / own.sum():6.0f} sum+far+fetch {prep.sum() / own.sum():6.0f} | outside the loop per tile {(tot - wait - relax - prep).mean():8.0f}")


if __name__ == "__main__":
    main()
