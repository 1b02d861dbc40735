#!/usr/bin/env python
"""What the PCIe link gives on this box: H2D alone, D2H alone, both directions, and two
H2D copies on two streams at once (does a second copy engine add bandwidth?)."""
import json
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from k8s_gpu_hpa_b200 import vector_add as va  # noqa: E402

n = 1 << 28
bufs = [va.PinnedBuffer(n) for _ in range(3)]
ha, hb, hc = (torch.from_numpy(p.array) for p in bufs)
ha.fill_(1.0); hb.fill_(2.0)
da, db, dc = (torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(3))
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    out = []
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in (s1, s2, s3):
            s.wait_event(e0)
        fn()
        for s in (s1, s2, s3):
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    out = sorted(out[1:])
    return out[len(out) // 2]


def h2d_one():
    with torch.cuda.stream(s1):
        da.copy_(ha, non_blocking=True); db.copy_(hb, non_blocking=True)


def h2d_two():
    with torch.cuda.stream(s1):
        da.copy_(ha, non_blocking=True)
    with torch.cuda.stream(s2):
        db.copy_(hb, non_blocking=True)


def d2h():
    with torch.cuda.stream(s3):
        hc.copy_(dc, non_blocking=True)


def both():
    h2d_one(); d2h()


def both_two():
    h2d_two(); d2h()


GiB = 1 << 30
for name, fn, gib_in, gib_out in (("h2d_one_stream", h2d_one, 2, 0), ("h2d_two_streams", h2d_two, 2, 0), ("d2h", d2h, 0, 1),
                                  ("h2d_one_stream+d2h", both, 2, 1), ("h2d_two_streams+d2h", both_two, 2, 1)):
    ms = timed(fn)
    print(json.dumps({"case": name, "ms": ms, "h2d_GBps": gib_in * GiB / ms / 1e6, "d2h_GBps": gib_out * GiB / ms / 1e6}), flush=True)
