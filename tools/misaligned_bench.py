#!/usr/bin/env python
"""Throughput of the C-ABI call when A, B, C do not share a 16-byte phase (mixed misalignment)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from k8s_gpu_hpa_b200 import vector_add as va  # noqa: E402

n = 1 << 28
a = torch.empty(n + 8, dtype=torch.float32, device="cuda")
b = torch.empty(n + 8, dtype=torch.float32, device="cuda")
c = torch.empty(n + 8, dtype=torch.float32, device="cuda")
va.fill_ctr(a, 0x0A); va.fill_ctr(b, 0x0B)
for name, (oa, ob, oc), variant in (("aligned auto", (0, 0, 0), "auto"), ("same phase +1 auto", (1, 1, 1), "auto"),
                                    ("mixed (1,2,3) auto", (1, 2, 3), "auto"), ("mixed (0,0,1) auto", (0, 0, 1), "auto"),
                                    ("mixed (1,2,3) k0 control", (1, 2, 3), "k0")):
    x, y, z = a[oa:oa + n], b[ob:ob + n], c[oc:oc + n]
    for _ in range(3):
        va.add(x, y, z, variant=variant)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        va.add(x, y, z, variant=variant)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    bad, _ = va.verify(x, y, z)
    print(json.dumps({"case": name, "ms": ms, "GBps": 12 * n / ms / 1e6, "mismatches": bad}), flush=True)
