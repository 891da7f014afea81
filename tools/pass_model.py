#!/usr/bin/env python
"""Timing model of the pass sweep's schedule (csrc/device/pass_plan.h: simulate_pass_sweep) on the levels of an RS hierarchy —
CPU only.  Prints, per level, the modelled forward-sweep time for 148 CTAs, for a zero hand-off latency and for unlimited CTAs
(= the critical path of the plan).  B200AMG_MODEL_TPASS_NS sets the time of one pass (default 250 ns).

    python tools/pass_model.py [size]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import algebraicmultigrid_jl_b200 as amg
from algebraicmultigrid_jl_b200 import _devlib
size=int(sys.argv[1]) if len(sys.argv)>1 else 128
A=amg.poisson((size,)*3)
ml=amg.ruge_stuben(A)
for lv in ml.levels[:4]:
    M=lv.A
    r=np.random.default_rng(0); x=r.random(M.n); b=r.random(M.n)
    st,msg,_,_=_devlib.block_plan_check(M,x,b,sweep=1,emulate_pass=True,verbose=1)
    print({k:st[k] for k in ("ok","tiles","a","b","max_tile_steps")}, flush=True)
