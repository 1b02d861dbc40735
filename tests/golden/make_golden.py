#!/usr/bin/env python
"""Regenerates tests/golden/*.json|npz from the CPU oracle.

The reference has no test vectors for this path: these
fixtures are OUR known answers, produced by oracle/vadd_oracle.c (hardware IEEE-754 add,
cross-checked against the integer soft-float implementation while generating) on:
  * the sample's input recipe  (glibc rand(), seed 1, N = 50000)   -> rand_50000.json
  * the counter generator      (N = 2^20 + 5 at index offset 12345) -> ctr_1m.json
  * special values             (+-0, subnormals, +-Inf, NaN, cancellation, overflow)
                                                                    -> special_values.npz
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def special_values() -> tuple[np.ndarray, np.ndarray]:
    pats = [0x00000000, 0x80000000, 0x00000001, 0x80000001, 0x007FFFFF, 0x807FFFFF, 0x00800000, 0x80800000,
            0x3F800000, 0xBF800000, 0x3F7FFFFF, 0x3F800001, 0x33800000, 0x34000000, 0x4B800000, 0x4B7FFFFF,
            0x7F7FFFFF, 0xFF7FFFFF, 0x7F000000, 0x7F800000, 0xFF800000, 0x7FC00000, 0xFFC00001, 0x7F800001,
            0x00400000, 0x80400000, 0x3EAAAAAB, 0x40490FDB, 0x7E967699, 0x00FFFFFF]
    ua = np.repeat(np.array(pats, dtype=np.uint32), len(pats))
    ub = np.tile(np.array(pats, dtype=np.uint32), len(pats))
    return ua, ub


def main() -> None:
    # 1. the sample's recipe
    a, b = oracle.fill_rand(50000)
    c = oracle.vadd(a, b)
    assert oracle.first_mismatch(c, oracle.softfloat_vadd_bits(a.view(np.uint32), b.view(np.uint32)).view(np.float32)) < 0
    s, x = oracle.bits_digest(c)
    json.dump({
        "n": 50000, "recipe": "h_A[i]=rand()/(float)RAND_MAX; h_B[i]=rand()/(float)RAND_MAX; no srand (glibc)",
        "first_rand": [1804289383, 846930886, 1681692777, 1714636915],
        "A_bits": {str(i): f"{a.view(np.uint32)[i]:08x}" for i in (0, 1, 2, 3, 49999)},
        "B_bits": {str(i): f"{b.view(np.uint32)[i]:08x}" for i in (0, 1, 2, 3, 49999)},
        "C_bits": {str(i): f"{c.view(np.uint32)[i]:08x}" for i in (0, 1, 2, 3, 49999)},
        "fnv1a64": {"A": f"{oracle.fnv1a64(a):016x}", "B": f"{oracle.fnv1a64(b):016x}", "C": f"{oracle.fnv1a64(c):016x}"},
        "C_bits_sum": s, "C_bits_xor": f"{x:08x}", "C_sum_f64": float(c.astype(np.float64).sum()),
    }, open(os.path.join(HERE, "rand_50000.json"), "w"), indent=1)

    # 2. the counter generator
    n, first = (1 << 20) + 5, 12345
    a, b = oracle.fill_ctr(n, 0x0A, first), oracle.fill_ctr(n, 0x0B, first)
    c = oracle.vadd(a, b)
    s, x = oracle.bits_digest(c)
    json.dump({
        "n": n, "first": first, "seed_a": 0x0A, "seed_b": 0x0B,
        "A_head_bits": [f"{v:08x}" for v in a.view(np.uint32)[:8]],
        "B_head_bits": [f"{v:08x}" for v in b.view(np.uint32)[:8]],
        "C_head_bits": [f"{v:08x}" for v in c.view(np.uint32)[:8]],
        "fnv1a64": {"A": f"{oracle.fnv1a64(a):016x}", "B": f"{oracle.fnv1a64(b):016x}", "C": f"{oracle.fnv1a64(c):016x}"},
        "C_bits_sum": s, "C_bits_xor": f"{x:08x}",
    }, open(os.path.join(HERE, "ctr_1m.json"), "w"), indent=1)

    # 3. special values: hardware add and soft-float must agree before anything is written
    ua, ub = special_values()
    hw = oracle.vadd(ua.view(np.float32), ub.view(np.float32)).view(np.uint32)
    sf = oracle.softfloat_vadd_bits(ua, ub)
    assert oracle.first_mismatch(hw.view(np.float32), sf.view(np.float32)) < 0
    np.savez_compressed(os.path.join(HERE, "special_values.npz"), a_bits=ua, b_bits=ub, c_bits=sf)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
