#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline: fp32 elements/sec on 2^28-element vectorAdd.

    python bench.py --gpus N --steps K --warmup W            # one rank per GPU under torchrun for N>1
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path: C = A + B over this rank's 2^28-element shard
(weak scaling: every GPU holds 2^28 elements, 3 GiB of operands, far above the 126 MB L2,
so no L2 flush is needed between steps).  One step == one launch of OUR kernel through the
C ABI (include/b200va.h: b200va_add_f32).

value    whole-job elements/s with operands resident in HBM (CUDA events, max over ranks)
e2e      the same metric through the host-buffer C-ABI call (b200va_stager_add_f32): every
         step copies A and B from pinned host memory to the GPU, adds, and copies C back
roofline dominant (only) kernel vs the measured HBM copy peak (MEASURED_PEAKS.json)
cpu_baseline / --impl reference
         the oracle port (oracle/vadd_oracle.c: the reference ships no source, so kind =
         "port") timed on this box's host cores.  The oracle is never on the product path.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fp32 elements/sec on 2^28-elem vectorAdd"
UNIT = "elements/s"
N_PER_GPU = 1 << 28
CLOCK_PERIOD_MS = 2.0
STAGE_MODE = 2       # host-path pipeline used for e2e (0 slot streams, 2 lanes: 42.25 vs 42.56 ms, profiles/r01/m_*)
BYTES_PER_ELEM = 12  # 4 read A + 4 read B + 4 write C


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while a timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, cuda_index: int, period_s: float = 0.002):
        self.period_s = period_s
        self.samples: list[int] = []
        self.reason_bits = 0
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            self._nv = pynvml
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(cuda_index).uuid)
                self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # NVML missing: report clocks as unknown, do not fail the bench
            self._err = repr(e)

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            except Exception:
                pass
            time.sleep(self.period_s)

    def __enter__(self):
        if self._h is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self) -> dict:
        if self._h is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in self.REASONS.items() if self.reason_bits & b], "samples": len(s)}


def measured_peak() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic() -> float | None:
    """dram read+write bytes per launch of the dominant kernel from the committed ncu capture."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))["dram_bytes_per_launch"])
    except Exception:
        return None


# --------------------------------------------------------------------------- reference arm
def cpu_time_passes(n: int, threads: int, warmup: int, steps: int) -> list[float]:
    import oracle  # the checker/baseline; never imported by the product package

    return oracle.time_vadd_mt(n, threads, warmup, steps)


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle

    threads, rates = oracle.best_thread_count()   # all the host threads the add can use (quota-aware)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    steps = min(steps, 50)  # 2^28 elements per pass on host cores: keep the run within minutes
    secs = cpu_time_passes(N_PER_GPU, threads, warmup, steps)
    total = sum(secs)
    value = N_PER_GPU * steps / total
    sample = (f"{steps} passes of C=A+B over 2^28 fp32 elements (one GPU's shard of the workload), "
              f"{threads} host threads (fastest of {sorted(rates)} tried; {oracle.num_cpus()} CPUs in the affinity mask, "
              f"cgroup quota {oracle.cpu_quota() or 'none'}), contiguous static partition, regular stores")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * total / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "vectorAdd N=2^28 fp32 (BASELINE.json configs[1]) on host cores",
                   "n_per_step": N_PER_GPU, "inputs": "ctr generator seeds 0x0A/0x0B"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "algorithmic_GBps": value * BYTES_PER_ELEM / 1e9},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference ships no source for this path (image k8s.gcr.io/cuda-vector-add:v0.1); "
                "this is the oracle port of its arithmetic on all host threads",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def run_ours(args, emit=print) -> None:
    import torch

    import k8s_gpu_hpa_b200 as pkg
    from k8s_gpu_hpa_b200 import sharding, vector_add as va

    rank, ws, local_rank = sharding.world()
    if ws != args.gpus:
        if ws == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torchrun (one rank per GPU)")
        args.gpus = ws
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the vectorAdd hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sharding.init("nccl")

    n = args.n_per_gpu
    first = rank * n  # this rank's shard of the global index space [0, ws*n)
    variant = pkg.VARIANTS[args.kernel]
    a = torch.empty(n, dtype=torch.float32, device=dev)
    b = torch.empty(n, dtype=torch.float32, device=dev)
    c = torch.empty(n, dtype=torch.float32, device=dev)
    va.fill_ctr(a, 0x0A, first)
    va.fill_ctr(b, 0x0B, first)
    stream = torch.cuda.current_stream()
    tune = pkg.resolve(variant, n)

    # ---- device-resident timing: W warm-up, K timed steps, barrier + sync both sides
    for _ in range(max(3, args.warmup)):
        va.add(a, b, c, variant=variant)
    torch.cuda.synchronize()
    sharding.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank, args.clock_period_ms * 1e-3)
    torch.cuda.synchronize()
    with sampler:
        ev0.record(stream)
        for _ in range(args.steps):
            va.add(a, b, c, variant=variant)      # one C-ABI call == one kernel launch
        ev1.record(stream)
        torch.cuda.synchronize()
    sharding.barrier()
    ms_local = ev0.elapsed_time(ev1)
    ms_total = sharding.max_over_ranks(ms_local)
    launches = int(sharding.sum_over_ranks(args.steps))

    # ---- correctness of what was timed (outside the timed region): bit-exact recompute in
    # HBM plus the order-independent digest, combined over shards
    bad, first_bad = va.verify(a, b, c)
    bad_total = int(sharding.sum_over_ranks(bad))
    dig = sharding.combine_digests(va.digest(c))

    # ---- e2e: host buffers through the C ABI, H2D + add + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        # pinned host buffers from the C ABI (pages on the GPU's NUMA node), filled from the
        # device arrays outside the timed region
        pa, pb, pc = va.PinnedBuffer(n, args.wc_inputs), va.PinnedBuffer(n, args.wc_inputs), va.PinnedBuffer(n)
        # [gpu's NUMA node, node of A, B, C] for every rank: the pinned buffers should sit next to their GPU
        host_nodes = sharding.gather_ints([int(pkg.capi.lib.b200va_device_numa_node())] + [p.numa_node for p in (pa, pb, pc)])
        ha, hb, hc = (torch.from_numpy(p.array) for p in (pa, pb, pc))
        ha.copy_(a); hb.copy_(b)
        torch.cuda.synchronize()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        with va.Stager(local_rank, args.chunk_elems, args.depth) as stg:
            mode = 1 if args.zero_copy else args.stage_mode
            for _ in range(2):
                stg.add(ha, hb, hc, variant=variant, mode=mode)
            sharding.barrier()
            ms_e2e = 0.0
            for _ in range(e2e_steps):
                ms_e2e += stg.add(ha, hb, hc, variant=variant, mode=mode)
            sharding.barrier()
        ms_e2e = sharding.max_over_ranks(ms_e2e)
        # the step's result must be the right one
        c2 = torch.empty_like(c)
        c2.copy_(hc)
        bad2, _ = va.verify(a, b, c2)
        bad_total += int(sharding.sum_over_ranks(bad2))
        del c2
        e2e = {"value": ws * n * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * n * ws,
               "d2h_bytes_per_step": 4 * n * ws, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
               "path": "b200va_stager_add_f32 " + {0: "copy-engine pipeline, one stream per slot: H2D(A,B) -> add -> D2H(C) per chunk",
                                                   1: "zero-copy kernel over PCIe",
                                                   2: "copy-engine pipeline, one stream per direction (lanes): H2D(A,B) | add | D2H(C)"}[mode],
               "host_memory": "pinned" + (", write-combined inputs" if args.wc_inputs else ""),
               "numa_gpu_A_B_C_per_rank": host_nodes}
        del ha, hb, hc
        for p in (pa, pb, pc):
            p.free()

    if rank != 0:
        return
    if bad_total:
        raise SystemExit(f"bit-exactness check failed: {bad_total} mismatching elements (first at {first_bad})")

    value = ws * n * args.steps / (ms_total * 1e-3)
    ms_per_step = ms_total / args.steps
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_ELEM * n / (ms_per_step * 1e-3) / 1e9  # per GPU, GB/s
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "vectorAdd N=2^28 fp32 per B200 (BASELINE.json configs[1]); weak scaling: "
                               f"global N = {ws} x 2^28, contiguous shards, no collective on the data path",
                   "n_per_gpu": n, "global_n": ws * n, "kernel": tune.as_dict(),
                   "inputs": "ctr generator (splitmix64 of the global index), seeds 0x0A/0x0B, uniform [0,1)",
                   "l2": "operands 3 GiB per GPU >> 126 MB L2: inputs larger than L2, no flush between steps",
                   "verified": "bit-exact recompute in HBM after the timed region",
                   "digest_sum": f"{dig[0]:016x}", "digest_xor": f"{dig[1]:08x}"},
        "algorithmic_GBps": value * BYTES_PER_ELEM / 1e9,
        "frac_of_8TBps_nameplate_per_gpu": achieved / 8000.0,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_ELEM * n,
                     "kernel": "b200va::vadd_vec / vadd_tma (one launch per step)"},
        "clocks": sampler.summary(),
        "gpu_launches": launches,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if ws == 1 and not args.no_cpu_baseline:
        import oracle

        threads, rates = oracle.best_thread_count()
        secs = cpu_time_passes(n, threads, 1, 5)
        v = n * len(secs) / sum(secs)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"5 passes over the full 2^28-element workload, {threads} host threads (fastest of "
                                          f"{sorted(rates)} tried; affinity {oracle.num_cpus()} CPUs, cgroup quota "
                                          f"{oracle.cpu_quota() or 'none'}), oracle/vadd_oracle.c (reference ships no source: "
                                          "port of its arithmetic)",
                                "algorithmic_GBps": v * BYTES_PER_ELEM / 1e9}
    emit(json.dumps(line))


class SingleLineStdout:
    """stdout carries exactly ONE JSON line: while the run is in progress fd 1 points at stderr,
    so library banners (NCCL prints its version on stdout) cannot get in front of it."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self.emit

    def emit(self, line: str) -> None:
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--kernel", choices=["auto", "k0", "k1", "k2", "k3"], default="auto")
    ap.add_argument("--n-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--chunk-elems", type=int, default=0)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--zero-copy", action="store_true")
    ap.add_argument("--stage-mode", type=int, choices=[0, 2], default=STAGE_MODE)
    ap.add_argument("--wc-inputs", action="store_true", help="write-combined pinned memory for the H2D sources")
    ap.add_argument("--clock-period-ms", type=float, default=CLOCK_PERIOD_MS, help="NVML clock sampling period during the timed region")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    try:
        with SingleLineStdout() as emit:
            run_ours(args, emit)
    finally:
        try:
            import torch.distributed as dist

            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
