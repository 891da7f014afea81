#!/usr/bin/env python
"""CPU model of the BLOCKED exact-order Gauss-Seidel sweep (csrc/device/block_gs.cuh) on real RS hierarchies.

Rows are grouped into tiles (contiguous index ranges); a tile is relaxed by ONE CTA in the order of its LOCAL
level schedule (dependencies inside the tile only), hand-off between steps through shared memory (cost c_step);
dependencies on other tiles go through L2 with latency lam and are waited for per STAGE (a group of consecutive
steps).  The model replays the tile DAG with P persistent CTAs claiming tiles in order and reports the sweep time
per level next to the bandwidth floor — used to choose tile sizes / stage lengths before spending GPU time.

    python tools/block_sweep_sim.py --size 128 [--tile-rows 8192] [--stage-steps 4] [--lam 1.5] [--cstep 0.2]
"""
import argparse
import heapq
import os
import sys
import time

import numpy as np
from numba import njit

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@njit(cache=True)
def global_levels(n, ptr, idx):
    lev = np.zeros(n, np.int32)
    for i in range(n):
        lv = 0
        for k in range(ptr[i], ptr[i + 1]):
            j = idx[k]
            if j < i and lev[j] + 1 > lv:
                lv = lev[j] + 1
        lev[i] = lv
    return lev


@njit(cache=True)
def local_steps(n, ptr, idx, tile_of):
    """local level of every row counting only dependencies inside its own tile"""
    st = np.zeros(n, np.int32)
    for i in range(n):
        lv = 0
        t = tile_of[i]
        for k in range(ptr[i], ptr[i + 1]):
            j = idx[k]
            if j < i and tile_of[j] == t and st[j] + 1 > lv:
                lv = st[j] + 1
        st[i] = lv
    return st


@njit(cache=True)
def requirements(n, ptr, idx, tile_of, step, stage_of_step_ptr, stage_of_step, nst_total, stage_ptr):
    """for every stage (global id) the max (tile, step) it needs from every OTHER tile: returned as dense per-stage
    dict emulation: arrays req_tile, req_step with up to 16 entries per stage"""
    MAXR = 24
    req_tile = -np.ones((nst_total, MAXR), np.int32)
    req_step = -np.ones((nst_total, MAXR), np.int32)
    over = 0
    for i in range(n):
        t = tile_of[i]
        g = stage_ptr[t] + stage_of_step[stage_of_step_ptr[t] + step[i]]
        for k in range(ptr[i], ptr[i + 1]):
            j = idx[k]
            if j < i and tile_of[j] != t:
                tj = tile_of[j]
                sj = step[j]
                found = False
                for q in range(MAXR):
                    if req_tile[g, q] == tj:
                        if sj > req_step[g, q]:
                            req_step[g, q] = sj
                        found = True
                        break
                    if req_tile[g, q] < 0:
                        req_tile[g, q] = tj
                        req_step[g, q] = sj
                        found = True
                        break
                if not found:
                    over += 1
    return req_tile, req_step, over


@njit(cache=True)
def monotone_coords(n, ptr, idx, theta):
    """K(i) = max over lower neighbours (K(j) + [i - j > theta]); J(i) = max (J(j) + (i - j if i - j <= theta else 0)).
    Both are non-decreasing along every dependency, so blocks of (K, J) form an acyclic tile graph."""
    K = np.zeros(n, np.int64)
    J = np.zeros(n, np.int64)
    for i in range(n):
        k = 0
        jj = 0
        for q in range(ptr[i], ptr[i + 1]):
            j = idx[q]
            if j < i:
                d = i - j
                if d > theta:
                    if K[j] + 1 > k:
                        k = K[j] + 1
                    if J[j] > jj:
                        jj = J[j]
                else:
                    if K[j] > k:
                        k = K[j]
                    if J[j] + d > jj:
                        jj = J[j] + d
        K[i] = k
        J[i] = jj
    return K, J


def valley_threshold(n, ptr, idx):
    rows = np.repeat(np.arange(n), np.diff(ptr))
    d = rows - idx
    d = d[d > 0]
    if d.size == 0:
        return None
    lg = np.floor(np.log2(d) * 2).astype(np.int64)
    h = np.bincount(lg)
    sig = h > 0.002 * d.size
    # the HIGHEST run of >= 3 empty half-octave bins (a factor >= 2.8 in distance) between two populated bins
    top = len(h) - 1
    while top >= 0 and not sig[top]:
        top -= 1
    hi = lo = -1
    b = top
    while b > 0:
        if not sig[b]:
            e = b
            while b >= 0 and not sig[b]:
                b -= 1
            if b >= 0 and e - b >= 3:
                hi, lo = e + 1, b
                break
        else:
            b -= 1
    if hi < 0:
        return None
    return float(2.0 ** ((hi + lo + 1) / 4.0))


def pencil_tiles(n, ptr, idx, a_rows, b_planes):
    theta = valley_threshold(n, ptr, idx)
    if theta is None:
        return None, None
    K, J = monotone_coords(n, ptr, idx, int(theta))
    kb = K // b_planes
    ja = J // a_rows
    nJ = int(ja.max()) + 1
    raw = kb * nJ + ja
    uniq, tile_of = np.unique(raw, return_inverse=True)
    return tile_of.astype(np.int32), dict(theta=theta, Kmax=int(K.max()), Jmax=int(J.max()), nJ=nJ)


def simulate(A_ptr, A_idx, n, tile_rows, stage_steps, stage_rows, lam, cstep, crow, P, verbose=False, tile_of=None):
    if tile_of is None:
        tile_of = (np.arange(n) // tile_rows).astype(np.int32)
    ntiles = int(tile_of.max()) + 1
    step = local_steps(n, A_ptr, A_idx, tile_of)
    # steps per tile, rows per (tile, step)
    nsteps = np.zeros(ntiles, np.int64)
    np.maximum.at(nsteps, tile_of, step + 1)
    sptr = np.zeros(ntiles + 1, np.int64)
    sptr[1:] = np.cumsum(nsteps)
    rows_in_step = np.zeros(sptr[-1], np.int64)
    np.add.at(rows_in_step, sptr[tile_of] + step, 1)
    rowlen = np.diff(A_ptr)
    nnz_in_step = np.zeros(sptr[-1], np.int64)
    np.add.at(nnz_in_step, sptr[tile_of] + step, rowlen)
    # stages: consecutive steps, <= stage_steps steps and <= stage_rows rows
    stage_of_step = np.zeros(sptr[-1], np.int32)
    stage_ptr = np.zeros(ntiles + 1, np.int64)
    stage_cost = []
    stage_last_step = []
    for t in range(ntiles):
        g = 0
        acc_s = acc_r = 0
        cost = 0.0
        for s in range(int(nsteps[t])):
            r = int(rows_in_step[sptr[t] + s])
            if acc_s > 0 and (acc_s + 1 > stage_steps or acc_r + r > stage_rows):
                stage_cost.append(cost)
                stage_last_step.append(s - 1)
                g += 1
                acc_s = acc_r = 0
                cost = 0.0
            stage_of_step[sptr[t] + s] = g
            acc_s += 1
            acc_r += r
            cost += cstep + crow * int(nnz_in_step[sptr[t] + s])
        stage_cost.append(cost)
        stage_last_step.append(int(nsteps[t]) - 1)
        stage_ptr[t + 1] = stage_ptr[t] + g + 1
    nst = int(stage_ptr[-1])
    stage_cost = np.array(stage_cost)
    if verbose:
        sizes = np.bincount(tile_of)
        print(f"      tile rows min/mean/max {sizes.min()}/{sizes.mean():.0f}/{sizes.max()}")
    req_tile, req_step, over = requirements(n, A_ptr, A_idx, tile_of, step, sptr, stage_of_step, nst, stage_ptr)
    # replay
    step_end = np.zeros(sptr[-1])          # completion time of every (tile, step)
    free = [0.0] * P
    heapq.heapify(free)
    t_end_tile = np.zeros(ntiles)
    stall = 0.0
    for t in range(ntiles):
        now = heapq.heappop(free)
        for g in range(int(stage_ptr[t]), int(stage_ptr[t + 1])):
            ready = now
            for q in range(req_tile.shape[1]):
                tj = req_tile[g, q]
                if tj < 0:
                    break
                ready = max(ready, step_end[sptr[tj] + req_step[g, q]] + lam)
            stall += ready - now
            # steps of the stage complete one after the other
            s0 = 0 if g == stage_ptr[t] else stage_last_step[g - 1] + 1
            tcur = ready
            for s in range(s0, stage_last_step[g] + 1):
                tcur += cstep + crow * int(nnz_in_step[sptr[t] + s])
                step_end[sptr[t] + s] = tcur
            now = tcur
        t_end_tile[t] = now
        heapq.heappush(free, now)
    total = t_end_tile.max()
    return dict(ntiles=ntiles, nstages=nst, steps_max=int(nsteps.max()), steps_sum=int(nsteps.sum()), total_us=total,
                stall_us=stall, over=over, mean_step_rows=float(n / nsteps.sum()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--tile-rows", type=str, default="4096,8192,16384,32768")
    ap.add_argument("--stage-steps", type=int, default=4)
    ap.add_argument("--stage-rows", type=int, default=512)
    ap.add_argument("--lam", type=float, default=1.5, help="cross-tile hand-off latency, us")
    ap.add_argument("--cstep", type=float, default=0.2, help="us per local step (barrier + dependent arithmetic)")
    ap.add_argument("--crow", type=float, default=0.0, help="us per non-zero of a step (throughput of one CTA)")
    ap.add_argument("--ctas", type=int, default=296)
    ap.add_argument("--levels", type=str, default="")
    ap.add_argument("--pencil", type=str, default="", help="a_rows:b_planes[,a:b...] monotone-coordinate tiles")
    args = ap.parse_args()
    import algebraicmultigrid_jl_b200 as amg

    t0 = time.time()
    A = amg.poisson((args.size,) * 3)
    ml = amg.ruge_stuben(A)
    print(f"setup {time.time() - t0:.1f} s; levels {[l.A.n for l in ml.levels]}")
    want = [int(v) for v in args.levels.split(",")] if args.levels else range(len(ml.levels))
    for li in want:
        L = ml.levels[li].A
        n = L.n
        ptr, idx = L.colptr.astype(np.int64), L.rowval.astype(np.int64)
        lev = global_levels(n, ptr, idx)
        D = int(lev.max()) + 1
        nnz = int(ptr[-1])
        bw_floor_us = (12 * nnz + 28 * n) / 6.5e12 * 1e6
        print(f"level {li}: n={n} nnz/row={nnz / n:.1f} wavefronts={D} bw-floor {bw_floor_us:.0f} us  ideal-latency {D * args.cstep:.0f} us")
        for tr in [int(v) for v in args.tile_rows.split(",")]:
            if tr > n and tr != int(args.tile_rows.split(",")[0]):
                continue
            crow = args.crow
            if args.pencil:
                break
            r = simulate(ptr, idx, n, tr, args.stage_steps, args.stage_rows, args.lam, args.cstep, crow, args.ctas)
            if args.pencil:
                break
            print(f"   tile {tr:6d}: tiles {r['ntiles']:5d} stages {r['nstages']:6d} max-steps/tile {r['steps_max']:5d} mean rows/step {r['mean_step_rows']:6.1f} "
                  f"-> {r['total_us']:8.0f} us  (stall {r['stall_us'] / max(r['ntiles'], 1):6.1f} us/tile, req overflow {r['over']})")
        for spec in [v for v in args.pencil.split(",") if v]:
            a, b = [int(v) for v in spec.split(":")]
            tile_of, info = pencil_tiles(n, ptr, idx, a, b)
            if tile_of is None:
                print(f"   pencil {spec}: no distance valley -> contiguous tiles")
                continue
            r = simulate(ptr, idx, n, 0, args.stage_steps, args.stage_rows, args.lam, args.cstep, args.crow, args.ctas, verbose=True, tile_of=tile_of)
            print(f"   pencil a={a} b={b} theta={info['theta']:.0f} K={info['Kmax'] + 1} Jmax={info['Jmax']}: tiles {r['ntiles']:5d} stages {r['nstages']:6d} "
                  f"max-steps/tile {r['steps_max']:5d} mean rows/step {r['mean_step_rows']:6.1f} -> {r['total_us']:8.0f} us  (stall {r['stall_us'] / max(r['ntiles'], 1):6.1f} us/tile, over {r['over']})")


if __name__ == "__main__":
    main()
