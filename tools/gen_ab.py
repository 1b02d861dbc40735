#!/usr/bin/env python
"""Short-list geometries for the interleaved A/B pass of b200va_tune."""
import sys

K1, K2, K3 = 2, 3, 4
which = sys.argv[1] if len(sys.argv) > 1 else "round2"
print("1 256 1 0 0 0 0 0 0")
if which == "round1":
    for kind in (K1, K3):
        for threads, unroll in ((128, 1), (256, 1), (512, 1), (1024, 1), (128, 2), (256, 2), (512, 2), (1024, 2), (256, 4), (512, 4), (1024, 4), (256, 8)):
            for ld, st in ((1, 1), (0, 0), (3, 3), (1, 0), (3, 0), (0, 1), (3, 1)):
                print(kind, threads, unroll, 0, ld, st, 0, 0, 0)
    for threads, cps, stages, tile, mode in ((128, 1, 4, 8192, 0), (128, 2, 2, 8192, 0), (128, 1, 8, 4096, 0), (256, 1, 12, 4096, 1), (256, 1, 6, 16384, 1), (256, 1, 4, 8192, 0)):
        for ld in (0, 3):
            print(K2, threads, 0, cps, ld, 1, stages, tile, mode)
elif which == "width":
    # access width vs loads in flight: 4-byte x {4,8,16}, 16-byte x {1,2,4}, 32-byte x {1,2}
    for threads in (128, 256, 512, 1024):
        for u in (4, 8, 16):
            print(5, threads, u, 0, 0, 0, 0, 0, 0)
        for u in (1, 2, 4):
            print(K1, threads, u, 0, 0, 1, 0, 0, 0)
        for u in (1, 2):
            print(K3, threads, u, 0, 0, 1, 0, 0, 0)
elif which == "clc":
    print(K1, 512, 1, 0, 0, 1, 0, 0, 0)
    print(K3, 768, 2, 0, 0, 1, 0, 0, 0)
    print(K2, 128, 0, 1, 0, 1, 4, 8192, 0)
    for threads in (64, 128, 256, 512):
        for stages, tile in ((2, 8192), (3, 8192), (4, 8192), (6, 8192), (8, 8192), (2, 16384), (3, 16384), (4, 16384), (6, 16384),
                             (4, 4096), (8, 4096), (2, 32768), (3, 32768), (12, 8192), (6, 32768)):
            for ld, st in ((0, 1), (3, 1), (0, 0)):
                print(K2, threads, 0, 1, ld, st, stages, tile, 2)
elif which == "cold":
    # r02: cold (rotating-buffer) mid sizes -- what closes the launch-boundary bubble: early loads,
    # the CLC scheduler (K1c), whole-wave persistent grids, next to the r01 AUTO classes
    for early in (0, 1, 2):
        for threads, unroll in ((128, 1), (256, 1), (512, 1), (1024, 1), (128, 2), (256, 2), (512, 2), (128, 4), (256, 4)):
            for ld, st in ((0, 0), (0, 1), (3, 0)):
                print(K1, threads, unroll, 0, ld, st, 0, 0, 0, early, 0)
        for threads, unroll in ((256, 1), (512, 1), (256, 2)):
            print(K3, threads, unroll, 0, 0, 1, 0, 0, 0, early, 0)
        for threads, unroll in ((128, 2), (256, 2), (512, 2), (128, 4), (256, 4), (512, 4), (256, 8)):       # K1c
            for st in (0, 1):
                print(K1, threads, unroll, 0, 0, st, 0, 0, 0, early, 1)
        for threads, unroll, cps in ((256, 2, 8), (512, 1, 4), (256, 4, 8), (512, 2, 4), (256, 2, 4)):   # whole-wave persistent
            print(K1, threads, unroll, cps, 0, 1, 0, 0, 0, early, 0)
    for threads, stages in ((128, 3), (256, 4), (512, 8)):
        print(K2, threads, 0, 1, 0, 1, stages, 8192, 2)
elif which == "headline":
    # r02: 2^28 -- the production geometry, its early-load twin, K1c, and the r01 runners-up
    for early in (0, 1, 2):
        for threads, unroll in ((512, 1), (384, 1), (256, 1), (1024, 1), (256, 2), (512, 2)):
            print(K1, threads, unroll, 0, 0, 1, 0, 0, 0, early, 0)
        for threads, unroll in ((256, 2), (512, 2), (256, 4), (512, 4), (128, 4), (256, 8)):
            print(K1, threads, unroll, 0, 0, 1, 0, 0, 0, early, 1)
        print(K3, 768, 2, 0, 0, 1, 0, 0, 0, early, 0)
    print(K2, 512, 0, 1, 0, 1, 8, 8192, 2)
elif which == "skew":
    # r02 channel-phase experiment: only the production kernel and the control
    print(K1, 512, 1, 0, 0, 1, 0, 0, 0, 0, 0)
    print(K1, 512, 1, 0, 0, 1, 0, 0, 0, 1, 0)
    print(K3, 768, 2, 0, 0, 1, 0, 0, 0, 0, 0)
else:
    # round 2: thread counts around the winner, the L2::256B load hint, a few TMA shapes
    for kind in (K1, K3):
        for threads in (128, 192, 256, 320, 384, 448, 512, 576, 640, 768, 1024):
            for unroll in (1, 2):
                for ld, st in ((0, 1), (0, 0), (5, 1), (5, 0), (3, 0), (3, 1)):
                    print(kind, threads, unroll, 0, ld, st, 0, 0, 0)
    for threads, cps, stages, tile, mode in ((64, 2, 4, 8192, 0), (64, 4, 2, 8192, 0), (128, 1, 4, 8192, 0), (128, 2, 3, 8192, 0), (128, 1, 6, 8192, 0), (96, 2, 4, 8192, 0)):
        print(K2, threads, 0, cps, 0, 1, stages, tile, mode)
