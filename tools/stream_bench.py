#!/usr/bin/env python
"""STREAM table of the generalised core on one GPU (row f4): copy/scale/add/triad x
f32/f64/f16/bf16 at 1 GiB per array, CUDA-event batch timing, bit-exactness spot check."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402
from k8s_gpu_hpa_b200 import vector_add as va  # noqa: E402

TORCH = {"f32": torch.float32, "f64": torch.float64, "f16": torch.float16, "bf16": torch.bfloat16}
ES = {"f32": 4, "f64": 8, "f16": 2, "bf16": 2}
ARRAYS = {"copy": 2, "scale": 2, "add": 3, "triad": 3}


def main():
    nbytes = 1 << 30
    reps, rounds = 20, 5
    only = os.environ.get("STREAM_BENCH_DTYPES", "f32,f64,f16,bf16").split(",")
    for dt, tdt in TORCH.items():
        if dt not in only:
            continue
        n = nbytes // ES[dt]
        a = (torch.rand(n, device="cuda", dtype=torch.float32) * 2 - 1).to(tdt) if dt != "f64" else torch.rand(n, device="cuda", dtype=tdt)
        b = (torch.rand(n, device="cuda", dtype=torch.float32) * 2 - 1).to(tdt) if dt != "f64" else torch.rand(n, device="cuda", dtype=tdt)
        c = torch.empty_like(a)
        for op in ("copy", "scale", "add", "triad"):
            s = 3.0 if op in ("scale", "triad") else 0.0
            bb = b if op in ("add", "triad") else None
            va.stream(op, a, bb, c, scalar=s)
            torch.cuda.synchronize()
            # spot-check a window against the oracle
            lo, m = n // 3 + 5, 1 << 18
            def host(t):
                return t[lo:lo + m].view(torch.int16).cpu().numpy().view(np.uint16) if dt in ("f16", "bf16") else t[lo:lo + m].cpu().numpy()
            want = oracle.stream(op, dt, host(a), host(b) if bb is not None else None, s)
            bad = oracle.first_mismatch_bits(host(c), want, dt)
            ms = []
            for _ in range(rounds):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for _ in range(3):
                    va.stream(op, a, bb, c, scalar=s)
                e0.record()
                for _ in range(reps):
                    va.stream(op, a, bb, c, scalar=s)
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1) / reps)
            ms.sort()
            med = ms[len(ms) // 2]
            moved = ARRAYS[op] * nbytes
            print(json.dumps({"geometry": os.environ.get("B200VA_STREAM_GEOMETRY", "default"), "dtype": dt, "op": op, "n": n, "bytes_per_launch": moved, "ms_median": med, "ms_best": ms[0],
                              "GBps": moved / med / 1e6, "elements_per_s": n / (med * 1e-3), "first_mismatch": bad}), flush=True)
        del a, b, c


if __name__ == "__main__":
    main()
