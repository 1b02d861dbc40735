// b200va_tune -- geometry sweep for the vectorAdd kernels, through the C ABI.
//
// Reads one geometry per line from stdin:
//     kind threads unroll ctas_per_sm ld_hint st_hint stages tile_bytes store_mode
// and prints one JSON line per geometry: bit-exactness (device-side recompute +
// digest against a K0 run), median / best per-launch time over P interleaved rounds of
// R back-to-back launches between two CUDA events, algorithmic GB/s (12 B/element).
//
//     b200va_tune [--n ELEMS] [--reps R] [--warmup W] [--rounds P] < geometries.txt
//
// Not part of the reference's surface: a development tool for profiles/*.
#include <cuda_runtime.h>

#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200va.h"

#define CK(expr)                                                                           \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            std::fprintf(stderr, "%s failed: %s\n", #expr, cudaGetErrorString(e__));       \
            std::exit(1);                                                                  \
        }                                                                                  \
    } while (0)

#define VA(expr)                                                                           \
    do {                                                                                   \
        int rc__ = (expr);                                                                 \
        if (rc__ != B200VA_OK) {                                                           \
            std::fprintf(stderr, "%s failed: %s (%d)\n", #expr, b200va_strerror(rc__), rc__); \
            std::exit(1);                                                                  \
        }                                                                                  \
    } while (0)

int main(int argc, char** argv)
{
    size_t n = size_t{1} << 28;
    int reps = 20, warmup = 3, rounds = 5;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--n") && i + 1 < argc) n = std::strtoull(argv[++i], nullptr, 0);
        else if (!std::strcmp(argv[i], "--reps") && i + 1 < argc) reps = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--warmup") && i + 1 < argc) warmup = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--rounds") && i + 1 < argc) rounds = std::atoi(argv[++i]);
        else { std::fprintf(stderr, "usage: %s [--n N] [--reps R] [--warmup W] [--rounds P] < geometries\n", argv[0]); return 2; }
    }
    b200va_devinfo_t di;
    VA(b200va_query(0, &di));
    CK(cudaSetDevice(0));
    std::printf("{\"device\": \"%s\", \"sm_count\": %d, \"l2_bytes\": %d, \"n\": %zu}\n", di.name, di.sm_count,
                di.l2_bytes, n);

    float *dA, *dB, *dC;
    uint64_t *dRes, hRes[2], refDig[2];
    CK(cudaMalloc(&dA, n * sizeof(float)));
    CK(cudaMalloc(&dB, n * sizeof(float)));
    CK(cudaMalloc(&dC, n * sizeof(float)));
    CK(cudaMalloc(&dRes, 2 * sizeof(uint64_t)));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    VA(b200va_fill_ctr_f32(dA, n, 0x0A, 0, st));
    VA(b200va_fill_ctr_f32(dB, n, 0x0B, 0, st));

    // reference digest from the scalar control kernel
    VA(b200va_add_f32(dA, dB, dC, n, B200VA_K0_SCALAR, st));
    VA(b200va_digest_f32(dC, n, dRes, st));
    CK(cudaMemcpyAsync(refDig, dRes, sizeof refDig, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    // read every geometry, check each once for bit-exactness, then time them in `rounds`
    // interleaved passes (so drift hits all geometries alike).  One timing sample = `reps`
    // back-to-back launches between two events (no event inside the batch: the launches
    // chain through programmatic dependent launch exactly as in production).
    struct Geo { b200va_tune_t t; unsigned long long bad = 0; bool dig_ok = false; std::string err; std::vector<double> ms; };
    std::vector<Geo> geos;
    char line[512];
    while (std::fgets(line, sizeof line, stdin)) {
        if (line[0] == '#' || line[0] == '\n') continue;
        Geo g{};
        b200va_tune_t& t = g.t;
        if (std::sscanf(line, "%d %d %d %d %d %d %d %d %d", &t.kind, &t.threads, &t.unroll, &t.ctas_per_sm,
                        &t.ld_hint, &t.st_hint, &t.stages, &t.tile_bytes, &t.store_mode) != 9) {
            std::fprintf(stderr, "bad geometry line: %s", line);
            continue;
        }
        geos.push_back(g);
    }
    for (auto& g : geos) {
        CK(cudaMemsetAsync(dC, 0xff, n * sizeof(float), st));
        int rc = b200va_add_f32_tuned(dA, dB, dC, n, &g.t, st);
        cudaError_t se = cudaStreamSynchronize(st);
        if (rc != B200VA_OK || se != cudaSuccess) {
            g.err = rc != B200VA_OK ? b200va_strerror(rc) : cudaGetErrorString(se);
            if (se != cudaSuccess) { std::fprintf(stderr, "sticky CUDA error: %s\n", g.err.c_str()); return 1; }
            continue;
        }
        VA(b200va_verify_f32(dA, dB, dC, n, dRes, st));
        CK(cudaMemcpyAsync(hRes, dRes, sizeof hRes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        g.bad = hRes[0];
        VA(b200va_digest_f32(dC, n, dRes, st));
        CK(cudaMemcpyAsync(hRes, dRes, sizeof hRes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        g.dig_ok = hRes[0] == refDig[0] && hRes[1] == refDig[1];
    }
    for (int r = 0; r < rounds; ++r) {
        for (auto& g : geos) {
            if (!g.err.empty()) continue;
            for (int i = 0; i < warmup; ++i) VA(b200va_add_f32_tuned(dA, dB, dC, n, &g.t, st));
            CK(cudaEventRecord(e0, st));
            for (int i = 0; i < reps; ++i) VA(b200va_add_f32_tuned(dA, dB, dC, n, &g.t, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float total = 0.f;
            CK(cudaEventElapsedTime(&total, e0, e1));
            g.ms.push_back(total / reps);
        }
    }
    const double bytes = 12.0 * static_cast<double>(n);
    for (auto& g : geos) {
        const b200va_tune_t& t = g.t;
        std::printf("{\"kind\": %d, \"threads\": %d, \"unroll\": %d, \"ctas_per_sm\": %d, \"ld\": %d, \"st\": %d, "
                    "\"stages\": %d, \"tile_bytes\": %d, \"store_mode\": %d, ", t.kind, t.threads, t.unroll, t.ctas_per_sm,
                    t.ld_hint, t.st_hint, t.stages, t.tile_bytes, t.store_mode);
        if (!g.err.empty()) { std::printf("\"error\": \"%s\"}\n", g.err.c_str()); continue; }
        std::sort(g.ms.begin(), g.ms.end());
        double mean = 0;
        for (double v : g.ms) mean += v;
        mean /= static_cast<double>(g.ms.size());
        const double med = g.ms[g.ms.size() / 2], best = g.ms[0];
        std::printf("\"mismatches\": %llu, \"digest_ok\": %s, \"samples\": %zu, \"launches_per_sample\": %d, "
                    "\"ms_median\": %.5f, \"ms_best\": %.5f, \"ms_mean\": %.5f, \"GBps_median\": %.1f, "
                    "\"GBps_best\": %.1f, \"GBps_mean\": %.1f, \"elems_per_s\": %.4e}\n",
                    g.bad, g.dig_ok ? "true" : "false", g.ms.size(), reps, med, best, mean, bytes / med / 1e6,
                    bytes / best / 1e6, bytes / mean / 1e6, static_cast<double>(n) / (med * 1e-3));
    }
    return 0;
}
