"""CPU oracle for the AMG solve phase — TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  See ``amg_oracle.c`` for the
reference file:line each routine follows and for how the oracle is pinned.
Parity status: PINNED against the reference's own golden vectors (tests/test_oracle_goldens.py).
"""
from .oracle import OracleHierarchy, build, mul, norm, smooth  # noqa: F401
