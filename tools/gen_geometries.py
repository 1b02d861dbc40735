#!/usr/bin/env python
"""Emit geometry lines for b200va_tune (kind threads unroll ctas_per_sm ld st stages tile_bytes store_mode)."""
import sys

K1, K2, K3 = 2, 3, 4


def vec_geometry():
    for kind in (K1, K3):
        for threads in (128, 256, 512, 1024):
            for unroll in (1, 2, 4, 8):
                for cps in (0, 1, 2, 4, 8, 16):
                    if cps * threads > 2048:
                        continue
                    yield (kind, threads, unroll, cps, 1, 1, 0, 0, 0)


def vec_hints():
    for kind in (K1, K3):
        for threads, unroll, cps in ((256, 4, 0), (512, 2, 0), (256, 4, 4)):
            for ld in range(5):
                for st in range(4):
                    if (ld, st) != (1, 1):
                        yield (kind, threads, unroll, cps, ld, st, 0, 0, 0)


def tma():
    for mode in (0, 1):
        for threads in (128, 256, 512):
            for stages in (2, 3, 4, 6, 8, 12):
                for tile in (4096, 8192, 16384, 32768):
                    for cps in (1, 2):
                        smem = stages * 2 * tile + 16 * stages
                        if smem * cps > 227 * 1024:
                            continue
                        for ld in (0, 3):
                            yield (K2, threads, 0, cps, ld, 1, stages, tile, mode)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    gens = {"vec": [vec_geometry, vec_hints], "tma": [tma], "all": [vec_geometry, vec_hints, tma]}[which]
    print("1 256 1 0 0 0 0 0 0")  # K0 control first
    for g in gens:
        for t in g():
            print(*t)


if __name__ == "__main__":
    main()
