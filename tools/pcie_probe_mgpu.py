#!/usr/bin/env python
"""The host<->device ceiling of this box as a function of which GPUs copy at once.

For a list of GPU sets (one socket's GPUs, socket-interleaved pairs, all eight) every GPU of the
set copies PLAIN whole arrays concurrently -- H2D only (2 arrays in), D2H only (1 array out), and
the vectorAdd step's mix (2 in + 1 out at once) -- from pinned host memory placed on the GPU's own
NUMA node (or on a forced node: --placement remote).  One host thread per GPU, a barrier before
every repetition, device-side events, max over the GPUs of a set.  torch copies only: none of the
product's code is on this path, so the numbers are the platform's, not the pipeline's.

    python tools/pcie_probe_mgpu.py [--log2n 27] [--reps 3] [--sets "0;0,1;0,1,2,3;0,4;0,4,1,5;all"] [--placement local|remote]

One JSON line per (set, direction): ms, aggregate and per-GPU GB/s, per-socket GB/s.
"""
import argparse
import json
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from k8s_gpu_hpa_b200 import capi, vector_add as va  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=27)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--sets", default="0;0,1;0,1,2,3;0,4;0,4,1,5;all")
    ap.add_argument("--placement", choices=["local", "remote"], default="local")
    args = ap.parse_args()
    ndev = torch.cuda.device_count()
    n = 1 << args.log2n
    nodes = [int(capi.lib.b200va_device_numa_node_of(d)) for d in range(ndev)]
    all_nodes = sorted(set(nodes))
    sets = []
    for spec in args.sets.split(";"):
        devs = list(range(ndev)) if spec == "all" else [int(x) for x in spec.split(",")]
        if all(d < ndev for d in devs) and devs not in sets:
            sets.append(devs)
    used = sorted({d for s in sets for d in s})
    host, devbuf = {}, {}
    for d in used:
        torch.cuda.set_device(d)
        if args.placement == "remote" and len(all_nodes) > 1:
            os.environ["B200VA_NUMA_NODE"] = str([x for x in all_nodes if x != nodes[d]][0])
        bufs = [va.PinnedBuffer(n) for _ in range(3)]
        os.environ.pop("B200VA_NUMA_NODE", None)
        host[d] = [torch.from_numpy(p.array) for p in bufs]
        host[d].append(bufs)                                     # keep alive
        host[d][0].fill_(1.0); host[d][1].fill_(2.0)
        devbuf[d] = [torch.empty(n, dtype=torch.float32, device=f"cuda:{d}") for _ in range(3)]
    print(json.dumps({"gpus": ndev, "numa_node_of_gpu": nodes, "elements_per_array": n, "placement": args.placement,
                      "host_node_of_A": {d: bufs_[3][0].numa_node for d, bufs_ in host.items()}}), flush=True)

    def run(devs, direction):
        bar = threading.Barrier(len(devs))
        out = {}

        def worker(d):
            torch.cuda.set_device(d)
            ha, hb, hc = host[d][:3]
            da, db, dc = devbuf[d]
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            best = float("inf")
            for _ in range(args.reps + 1):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                torch.cuda.synchronize()
                bar.wait()
                e0.record()
                s_in.wait_event(e0); s_out.wait_event(e0)
                if direction in ("h2d", "both"):
                    with torch.cuda.stream(s_in):
                        da.copy_(ha, non_blocking=True); db.copy_(hb, non_blocking=True)
                if direction in ("d2h", "both"):
                    with torch.cuda.stream(s_out):
                        hc.copy_(dc, non_blocking=True)
                e1.record(s_in); e2.record(s_out)
                torch.cuda.synchronize()
                best = min(best, max(e0.elapsed_time(e1), e0.elapsed_time(e2)))
            out[d] = best

        th = [threading.Thread(target=worker, args=(d,)) for d in devs]
        for t in th:
            t.start()
        for t in th:
            t.join()
        return out

    for devs in sets:
        for direction in ("h2d", "d2h", "both"):
            per = run(devs, direction)
            ms = max(per.values())
            bytes_per_gpu = 4 * n * {"h2d": 2, "d2h": 1, "both": 3}[direction]
            by_socket = {}
            for d in devs:
                by_socket[nodes[d]] = by_socket.get(nodes[d], 0) + bytes_per_gpu / ms / 1e6
            print(json.dumps({"gpus": devs, "direction": direction, "ms": ms, "ms_per_gpu": per,
                              "GBps_total": len(devs) * bytes_per_gpu / ms / 1e6, "GBps_per_gpu": bytes_per_gpu / ms / 1e6,
                              "GBps_per_socket": by_socket,
                              "ms_for_a_2p28_step": ms * (1 << 28) / n if direction == "both" else None}), flush=True)


if __name__ == "__main__":
    main()
