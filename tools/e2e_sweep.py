#!/usr/bin/env python
"""Host-buffer path sweep on a GPU box: chunk size x depth x mode x host-memory placement.
Prints one JSON line per configuration (ms per 2^28-element step, elements/s, GB/s per PCIe direction)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from k8s_gpu_hpa_b200 import vector_add as va  # noqa: E402


def run(n, ha, hb, hc, chunk, depth, mode, steps=5):
    with va.Stager(0, chunk, depth) as st:
        st.add(ha, hb, hc, mode=int(mode))
        ms = [st.add(ha, hb, hc, mode=int(mode)) for _ in range(steps)]
    ms.sort()
    return ms[len(ms) // 2], ms[0]


def main():
    n = 1 << 28
    dev = torch.device("cuda:0")
    a = torch.empty(n, dtype=torch.float32, device=dev)
    b = torch.empty_like(a)
    va.fill_ctr(a, 0x0A)
    va.fill_ctr(b, 0x0B)
    placements = [("b200va_host_alloc(auto numa)", None)]
    for label, node in placements:
        if node == "torch":
            bufs = None
            ha, hb, hc = (torch.empty(n, dtype=torch.float32, pin_memory=True) for _ in range(3))
        else:
            if node is None:
                os.environ.pop("B200VA_NUMA_NODE", None)
            else:
                os.environ["B200VA_NUMA_NODE"] = node
            try:
                bufs = [va.PinnedBuffer(n) for _ in range(3)]
            except Exception as e:
                print(json.dumps({"placement": label, "error": repr(e)}), flush=True)
                continue
            ha, hb, hc = (torch.from_numpy(p.array) for p in bufs)
        ha.copy_(a); hb.copy_(b)
        torch.cuda.synchronize()
        grid = [(1 << 22, 3, False), (0, 0, True)] if label != "b200va_host_alloc(auto numa)" else \
            ([(c, d, m) for m in (0, 2) for c in (1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24, 1 << 25) for d in (2, 3, 4, 6)] + [(0, 0, 1)]
             if os.environ.get("E2E_GRID", "full") == "full" else
             [(c, d, 2) for c in (1 << 25, 1 << 26, 1 << 27, 0) for d in (2, 3)])
        for chunk, depth, zc in grid:
            med, best = run(n, ha, hb, hc, chunk, depth, zc)
            c = torch.empty_like(a); c.copy_(hc)
            bad, _ = va.verify(a, b, c)
            del c
            print(json.dumps({"placement": label, "chunk_elems": chunk, "depth": depth, "zero_copy": zc, "ms_median": med,
                              "ms_best": best, "elements_per_s": n / (med * 1e-3), "h2d_GBps": 8 * n / med / 1e6,
                              "d2h_GBps": 4 * n / med / 1e6, "mismatches": bad}), flush=True)
        del ha, hb, hc
        if bufs:
            for p in bufs:
                p.free()


if __name__ == "__main__":
    main()
