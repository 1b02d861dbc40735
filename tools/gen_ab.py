#!/usr/bin/env python
"""Short-list geometries for the interleaved A/B pass of b200va_tune."""
K1, K2, K3 = 2, 3, 4
print("1 256 1 0 0 0 0 0 0")
for kind in (K1, K3):
    for threads, unroll in ((128, 1), (256, 1), (512, 1), (1024, 1), (128, 2), (256, 2), (512, 2), (1024, 2), (256, 4), (512, 4), (1024, 4), (256, 8)):
        for ld, st in ((1, 1), (0, 0), (3, 3), (1, 0), (3, 0), (0, 1), (3, 1)):
            print(kind, threads, unroll, 0, ld, st, 0, 0, 0)
for threads, cps, stages, tile, mode in ((128, 1, 4, 8192, 0), (128, 2, 2, 8192, 0), (128, 1, 8, 4096, 0), (256, 1, 12, 4096, 1), (256, 1, 6, 16384, 1), (256, 1, 4, 8192, 0)):
    for ld in (0, 3):
        print(K2, threads, 0, cps, ld, 1, stages, tile, mode)
