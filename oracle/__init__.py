"""CPU oracle for the vectorAdd path -- TEST INFRASTRUCTURE ONLY (see vadd_oracle.c).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; the product package never imports it.  PARITY UNPINNED by the
reference (it ships no source, tests or vectors for this path): the oracle is pinned by
IEEE-754 (soft-float cross-check) and the glibc seed-1 known answers of SURVEY.md 8(c).
"""
from .oracle import *  # noqa: F401,F403
