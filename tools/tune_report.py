#!/usr/bin/env python
"""Rank a b200va_tune JSONL sweep and print the best geometries per kernel family."""
import json
import sys

KIND = {1: "k0_scalar", 2: "k1_vec128", 3: "k2_tma", 4: "k3_vec256", 5: "k4_scalar_mlp"}


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    rows, meta = [], None
    for line in open(path):
        d = json.loads(line)
        if "device" in d:
            meta = d
        else:
            rows.append(d)
    print("#", meta)
    bad = [r for r in rows if r.get("error") or r.get("mismatches") or not r.get("digest_ok", True)]
    print(f"# {len(rows)} geometries, {len(bad)} with errors/mismatches")
    for r in bad[:20]:
        print("BAD", r)
    ok = [r for r in rows if r not in bad]
    for kind in (1, 2, 4, 3, 5):
        sel = sorted((r for r in ok if r["kind"] == kind), key=lambda r: r["ms_median"])
        print(f"\n## {KIND[kind]}: {len(sel)} geometries")
        print("threads unroll cps ld st stages tile mode | ms_med ms_best ms_mean | GB/s(med) GB/s(mean)")
        for r in sel[:top]:
            print(f"{r['threads']:5d} {r['unroll']:3d} {r['ctas_per_sm']:3d} {r['ld']} {r['st']} {r['stages']:3d} {r['tile_bytes']:6d} {r['store_mode']} | "
                  f"{r['ms_median']:.4f} {r['ms_best']:.4f} {r['ms_mean']:.4f} | {r['GBps_median']:.0f} {r['GBps_mean']:.0f}")
        if sel:
            w = sel[-1]
            print(f"worst: {w['threads']} {w['unroll']} {w['ctas_per_sm']} {w['ld']} {w['st']} {w['stages']} {w['tile_bytes']} {w['store_mode']} -> {w['GBps_median']:.0f} GB/s")


if __name__ == "__main__":
    main()
