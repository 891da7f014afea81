#!/usr/bin/env python
"""CPU model (no GPU): how long would an exact-order Gauss-Seidel sweep take if dependent rows stayed on one SM?

The sweeps on the wide levels are bound by the hand-off through L2 (profiles/r01_gs_tile_timeline_256.log): today every
dependency edge costs one L2 hand-off.  Model of a SLAB-BLOCKED sweep over the existing wavefront numbering: CTA b owns
the b-th slice of every wavefront and walks the wavefronts in order; a step of a CTA (one wavefront slice) costs
`--intra` us (bar.sync + reading its own earlier results + arithmetic), and a row whose earlier-ordered neighbour lives in
another CTA's slice cannot start before that CTA finished the neighbour's step plus `--cross` us (the measured hand-off).
Reports the critical path of that schedule against `wavefronts x cross` (every edge through L2), per level.

    python tools/blocked_sweep_model.py [--size 64] [--blocks 148,296] [--intra 0.7] [--cross 2.25]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import algebraicmultigrid_jl_b200 as amg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--blocks", default="148,296")
ap.add_argument("--intra", type=float, default=0.7)
ap.add_argument("--cross", type=float, default=2.25)
ap.add_argument("--levels", type=int, default=3)
args = ap.parse_args()


def wavefronts(csr):
    """level schedule of the ascending-index sweep on the (symmetric) pattern: level[i] = 1 + max level of earlier neighbours"""
    n = csr.shape[0]
    ptr, idx = csr.indptr, csr.indices
    level = np.zeros(n, np.int64)
    for i in range(n):
        nb = idx[ptr[i]:ptr[i + 1]]
        nb = nb[nb < i]
        if len(nb):
            level[i] = level[nb].max() + 1
    return level


def model(csr, level, nblocks, intra, cross):
    n = csr.shape[0]
    nlev = int(level.max()) + 1
    order = np.lexsort((np.arange(n), level))          # wavefront order, ascending original index inside a wavefront
    new_of_old = np.empty(n, np.int64)
    new_of_old[order] = np.arange(n)
    lvlptr = np.searchsorted(level[order], np.arange(nlev + 1))
    pos = np.arange(n) - lvlptr[level[order]]           # position inside its wavefront (new numbering)
    length = (lvlptr[1:] - lvlptr[:-1])[level[order]]
    block = (pos * nblocks) // length                   # slab of the wavefront
    # edges (row -> earlier-ordered neighbour), in the new numbering
    coo = csr.tocoo()
    r, c = new_of_old[coo.row], new_of_old[coo.col]
    keep = coo.col < coo.row                            # earlier-ordered in the reference's (original) order
    r, c = r[keep], c[keep]
    lev_new = level[order]
    cross_edge = block[r] != block[c]
    frac_cross = cross_edge.mean()
    er, ec = r[cross_edge], c[cross_edge]
    o = np.argsort(lev_new[er], kind="stable")
    er, ec = er[o], ec[o]
    eptr = np.searchsorted(lev_new[er], np.arange(nlev + 1))
    T = np.zeros((nblocks, nlev))                       # finish time of step (b, w)
    prev = np.zeros(nblocks)
    for w in range(nlev):
        start = prev.copy()
        a, b = eptr[w], eptr[w + 1]
        if b > a:
            cand = T[block[ec[a:b]], lev_new[ec[a:b]]] + cross
            np.maximum.at(start, block[er[a:b]], cand)
        T[:, w] = start + intra
        # a CTA without rows in this wavefront does not pay the step
        has = np.zeros(nblocks, bool)
        has[np.unique(block[lvlptr[w]:lvlptr[w + 1]])] = True
        T[~has, w] = prev[~has]
        prev = T[:, w]
    return nlev, float(T[:, -1].max()), float(frac_cross)


def model_index_blocks(csr, nblocks, intra, cross):
    """Blocks = contiguous ranges of the ORIGINAL numbering: an earlier-ordered neighbour is always in the same or an
    EARLIER block, so cross-CTA dependencies flow one way and a CTA may trail its predecessors instead of marching in
    lock-step.  Inside a block: local level schedule (intra-block edges only), one step per local level."""
    n = csr.shape[0]
    ptr, idx = csr.indptr, csr.indices
    bounds = (np.arange(nblocks + 1) * n) // nblocks
    block = np.searchsorted(bounds, np.arange(n), side="right") - 1
    local = np.zeros(n, np.int64)
    for i in range(n):
        nb = idx[ptr[i]:ptr[i + 1]]
        nb = nb[(nb < i) & (nb >= bounds[block[i]])]
        if len(nb):
            local[i] = local[nb].max() + 1
    nsteps = np.zeros(nblocks, np.int64)
    np.maximum.at(nsteps, block, local + 1)
    coo = csr.tocoo()
    keep = (coo.col < coo.row) & (block[coo.col] != block[coo.row])
    er, ec = coo.row[keep], coo.col[keep]
    frac_cross = keep.sum() / max((coo.col < coo.row).sum(), 1)
    # process blocks in order (dependencies only point to earlier blocks)
    T = [np.zeros(int(s)) for s in nsteps]
    o = np.argsort(block[er], kind="stable")
    er, ec = er[o], ec[o]
    bptr = np.searchsorted(block[er], np.arange(nblocks + 1))
    for b in range(nblocks):
        start = np.zeros(int(nsteps[b]))
        a, e = bptr[b], bptr[b + 1]
        if e > a:
            cand = np.array([T[block[c]][local[c]] for c in ec[a:e]]) + cross
            np.maximum.at(start, local[er[a:e]], cand)
        t = 0.0
        for s_ in range(int(nsteps[b])):
            t = max(t, start[s_]) + intra
            T[b][s_] = t
    return int(nsteps.max()), float(max(t_[-1] for t_ in T if len(t_))), float(frac_cross)


A = amg.poisson((args.size,) * 3)
ml = amg.ruge_stuben(A)
print(f"model: intra-CTA step {args.intra} us, hand-off through L2 {args.cross} us; poisson {args.size}^3 RS hierarchy")
for lv, level in enumerate(ml.levels[: args.levels]):
    csr = level.A.to_scipy().tocsr()
    lev = wavefronts(csr)
    for nb in [int(v) for v in args.blocks.split(",")]:
        nlev, t, fc = model(csr, lev, nb, args.intra, args.cross)
        print(f"  level {lv}: n={csr.shape[0]} nnz/row={csr.nnz / csr.shape[0]:.1f} wavefronts={nlev} blocks={nb}: "
              f"cross-CTA edges {100 * fc:.1f} %; sweep {t:8.1f} us blocked vs {nlev * args.cross:8.1f} us all-through-L2 "
              f"({nlev * args.cross / t:.2f}x); floor {nlev * args.intra:.1f} us")
        ns, t2, fc2 = model_index_blocks(csr, nb, args.intra, args.cross)
        print(f"           blocks = index ranges: cross-CTA edges {100 * fc2:.1f} %, longest local schedule {ns} steps; sweep {t2:8.1f} us "
              f"({nlev * args.cross / t2:.2f}x of all-through-L2)")
