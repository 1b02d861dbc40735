#!/usr/bin/env python
"""Summarise .ncu-rep captures (read here, no GPU needed) into JSON + markdown for profiles/.

    python tools/ncu_summary.py gpurun_out/r01/prof_auto.ncu-rep [more.ncu-rep ...] --out profiles/r01 [--n N] [--table]

--table also merges one row per (kernel, n) into profiles/ncu_summary.json, the per-kernel table
bench.py reads roofline.traffic from (key "<kernel>@<n>", kernel as Tune.kernel_name()).
"""
import argparse
import re
import csv
import io
import json
import os
import subprocess

KEEP = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "dram__cycles_active.max.pct_of_peak_sustained_elapsed": "dram_busiest_channel_pct",
    "dram__cycles_active.min.pct_of_peak_sustained_elapsed": "dram_idlest_channel_pct",
    "dram__bytes.sum.peak_sustained": "dram_peak_bytes_per_cycle",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed": "dram_cycles_active_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_bytes",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "launch__occupancy_limit_warps": "occ_limit_warps",
    "launch__waves_per_multiprocessor": "waves_per_sm",
    "sm__cycles_elapsed.avg.per_second": "sm_clock_hz",
    "dram__cycles_elapsed.avg.per_second": "dram_clock_hz",
    "smsp__inst_executed.sum": "warp_insts",
    "l1tex__t_bytes.sum": "l1tex_bytes",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
}


def read_raw(rep: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {"kernel": r[header.index("Kernel Name")], "id": r[header.index("ID")]}
        for col, unit, val in zip(header, units, r):
            if col in KEEP:
                try:
                    v = float(val.replace(",", ""))
                except ValueError:
                    continue
                scale = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "us": 1e3, "ms": 1e6, "ns": 1,
                         "usecond": 1e3, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9, "Ghz": 1e9, "Mhz": 1e6,
                         "hz": 1}.get(unit, 1)
                if KEEP[col].endswith(("_bytes", "_ns", "_hz")):
                    v *= scale
                d[KEEP[col]] = v
        if "dram_read_bytes" in d and "dram_write_bytes" in d:
            d["dram_bytes"] = d["dram_read_bytes"] + d["dram_write_bytes"]
            if d.get("duration_ns"):
                d["dram_GBps"] = d["dram_bytes"] / d["duration_ns"]
            if d.get("dram_peak_bytes_per_cycle") and d.get("dram_clock_hz"):
                # ncu reports Kbyte/cycle for the chip-wide sum (2.048 -> 2048 B/cycle)
                bpc = d["dram_peak_bytes_per_cycle"] * (1e3 if d["dram_peak_bytes_per_cycle"] < 100 else 1)
                d["dram_pin_peak_GBps"] = bpc * d["dram_clock_hz"] / 1e9
        res.append(d)
    return res


def canonical(name: str) -> str:
    """'void b200va::vadd_vec<(int)4, (int)1, (int)0, (int)1, (bool)0>(const float *, ...)' -> 'vadd_vec<4,1,0,1,0>'"""
    name = re.sub(r"\(.*?\)(?=\s*\d|\s*true|\s*false)", "", name.split("(const")[0].split("(float")[0])
    name = name.replace("void ", "").replace("b200va::", "").replace("true", "1").replace("false", "0")
    return re.sub(r"\s+", "", name)


def merge_table(rows, n: int, path: str) -> None:
    try:
        doc = json.load(open(path))
    except Exception:
        doc = {"kernels": {}}
    groups = {}
    for d in rows:
        if "dram_bytes" in d:
            groups.setdefault(canonical(d["kernel"]), []).append(d)
    for k, g in groups.items():
        mean = lambda key: sum(x[key] for x in g if key in x) / max(1, sum(1 for x in g if key in x))  # noqa: E731
        doc["kernels"][f"{k}@{n}"] = {
            "kernel": "b200va::" + k, "source": "profiles/" + "/".join(g[0]["report_path"].split("/")[-2:]) + " (ncu --set full --clock-control none)",
            "launches_captured": len(g), "dram_bytes_per_launch": mean("dram_bytes"), "dram_read_bytes_per_launch": mean("dram_read_bytes"),
            "dram_write_bytes_per_launch": mean("dram_write_bytes"), "algorithmic_bytes_per_launch": 12 * n,
            "traffic_over_algorithmic": mean("dram_bytes") / (12 * n), "duration_us": [x["duration_ns"] / 1e3 for x in g],
            "dram_pct_of_pin_peak": [x.get("dram_pct_of_peak") for x in g], "busiest_channel_pct": [x.get("dram_busiest_channel_pct") for x in g],
            "idlest_channel_pct": [x.get("dram_idlest_channel_pct") for x in g], "l2_hit_rate_pct": [x.get("l2_hit_rate_pct") for x in g],
            "registers_per_thread": g[0].get("registers_per_thread"), "grid": g[0].get("grid"), "block": g[0].get("block"),
            "achieved_occupancy_pct": [x.get("achieved_occupancy_pct") for x in g]}
    json.dump(doc, open(path, "w"), indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--table", action="store_true", help="merge into profiles/ncu_summary.json")
    ap.add_argument("reps", nargs="+")
    ap.add_argument("--out", required=True, help="output prefix (writes <out>_ncu.json and <out>_ncu.md)")
    ap.add_argument("--n", type=int, default=1 << 28)
    args = ap.parse_args()
    allk = []
    for rep in args.reps:
        for d in read_raw(rep):
            d["report"] = os.path.basename(rep)
            d["report_path"] = os.path.abspath(rep)
            d["algorithmic_bytes"] = 12 * args.n
            if d.get("duration_ns"):
                d["algorithmic_GBps"] = d["algorithmic_bytes"] / d["duration_ns"]
            if d.get("dram_bytes"):
                d["traffic_over_algorithmic"] = d["dram_bytes"] / d["algorithmic_bytes"]
            allk.append(d)
    if args.table:
        merge_table(allk, args.n, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_summary.json"))
    for d in allk:
        d.pop("report_path", None)
    json.dump(allk, open(args.out + "_ncu.json", "w"), indent=1)
    cols = ["report", "kernel", "grid", "block", "registers_per_thread", "dyn_smem_bytes", "duration_ns", "algorithmic_GBps",
            "dram_GBps", "dram_bytes", "traffic_over_algorithmic", "dram_pct_of_peak", "l2_hit_rate_pct",
            "dram_busiest_channel_pct", "dram_idlest_channel_pct", "achieved_occupancy_pct", "sm_clock_hz", "dram_clock_hz"]
    with open(args.out + "_ncu.md", "w") as f:
        f.write("| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
        for d in allk:
            def fmt(c):
                v = d.get(c, "")
                if isinstance(v, float):
                    return f"{v:.4g}" if abs(v) < 1e6 else f"{v:.6g}"
                return str(v)[:60]
            f.write("| " + " | ".join(fmt(c) for c in cols) + " |\n")
    print(open(args.out + "_ncu.md").read())


if __name__ == "__main__":
    main()
