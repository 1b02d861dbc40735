#!/usr/bin/env python
"""Host-buffer entry points at 2^28 elements: pinned stager (lanes), one-shot b200va_add_f32_host on
pageable numpy arrays (what a fresh ./vectorAdd-like caller has), and the sample-mode CLI."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from k8s_gpu_hpa_b200 import vector_add as va  # noqa: E402

n = 1 << 28
a = va.fill_ctr_host(n, 0x0A)
b = va.fill_ctr_host(n, 0x0B)
c = np.empty_like(a)
torch.cuda.init()
for _ in range(3):
    t0 = time.perf_counter()
    va.add_host(a, b, c)
    dt = time.perf_counter() - t0
    print(json.dumps({"case": "b200va_add_f32_host, pageable arrays, one-shot (alloc+pipeline+free)", "ms": dt * 1e3,
                      "elements_per_s": n / dt}), flush=True)
for ch in (1 << 20, 1 << 21, 1 << 22):
    with va.Stager(0, ch, 3) as st:
        st.add(a, b, c, mode=3)
        ms = st.add(a, b, c, mode=3)
        print(json.dumps({"case": f"stager mode 3, chunk {ch}, steady", "ms": ms}), flush=True)
with va.Stager(0, 1 << 23, 3) as st:
    for _ in range(3):
        ms = st.add(a, b, c, mode=3)
        print(json.dumps({"case": "stager mode 3 (pageable bounce), persistent stager", "ms": ms, "elements_per_s": n / (ms * 1e-3)}), flush=True)
    for mode in (2, 0):
        ms = st.add(a, b, c, mode=mode)
        print(json.dumps({"case": f"stager mode {mode} fed pageable arrays (driver-staged cudaMemcpyAsync)", "ms": ms}), flush=True)
with va.Stager(0) as st:          # register-once (mode AUTO -> 4): the first call page-locks a, b, c in place
    t0 = time.perf_counter()
    ms = st.add(a, b, c, mode=-1)
    dt = time.perf_counter() - t0
    print(json.dumps({"case": "stager AUTO on pageable arrays, FIRST call (cudaHostRegister x3 + pipeline)", "wall_ms": dt * 1e3,
                      "pipeline_ms": ms, "stage_mode": st.last_mode}), flush=True)
    for _ in range(3):
        t0 = time.perf_counter()
        ms = st.add(a, b, c, mode=-1)
        dt = time.perf_counter() - t0
        print(json.dumps({"case": "stager AUTO on pageable arrays, steady (registration cached)", "wall_ms": dt * 1e3, "ms": ms,
                          "elements_per_s": n / (ms * 1e-3), "stage_mode": st.last_mode}), flush=True)
assert va.verify_host(a, b, c) == -1
bufs = [va.PinnedBuffer(n) for _ in range(3)]
for p, src in zip(bufs[:2], (a, b)):
    p.array[:] = src
with va.Stager(0) as st:
    st.add(bufs[0].array, bufs[1].array, bufs[2].array, mode=2)
    ms = st.add(bufs[0].array, bufs[1].array, bufs[2].array, mode=2)
print(json.dumps({"case": "stager lanes, pinned", "ms": ms, "elements_per_s": n / (ms * 1e-3)}), flush=True)
assert va.verify_host(a, b, bufs[2].array) == -1 and va.verify_host(a, b, c) == -1
t0 = time.perf_counter()
p = va.run_cli("--n", "2^28", "--gen", "ctr")
dt = time.perf_counter() - t0
print(json.dumps({"case": "./vectorAdd --n 2^28 --gen ctr (sample mode: malloc, host fill, cudaMemcpy, add, cudaMemcpy, verify)",
                  "wall_s": dt, "ok": p.returncode == 0 and "Test PASSED" in p.stdout}), flush=True)
