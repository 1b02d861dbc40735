// b200va_sweep -- BASELINE.json configs[3]: vectorAdd N sweep 2^lo..2^hi on one GPU
// (launch-overhead vs bandwidth-bound crossover), through the C ABI.
//
// For every N it reports the per-launch time of back-to-back launches in three regimes:
//   hot    the same three buffers every launch (for 12*N <= L2 this is an L2 number,
//          flagged "l2_resident": true -- NOT an HBM figure)
//   cold   rotating through enough buffer sets that the footprint is >= 4x L2, so every
//          launch streams from HBM (launched with B200VA_F_COLD: the cold-tuned AUTO class)
//   cold_chain  the cold rotation launched with B200VA_F_INPUTS_STABLE (consecutive launches never
//          write each other's inputs): loads run ahead of the programmatic dependency, so the
//          next launch's DRAM ramp overlaps the previous launch's tail
//   graph  hot buffers, launches captured 100 per CUDA graph (launch latency amortised; launches
//          2..100 of a graph run with early loads -- the loop API knows its own predecessor)
// plus the reference-shape control K0 (hot).  Times are CUDA-event batch times / launches.
//
//     b200va_sweep [--lo 16] [--hi 30] [--kernel auto|k0|k1|k2|k3]
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

static double median(std::vector<double> v)
{
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
}

int main(int argc, char** argv)
{
    int lo = 16, hi = 30, variant = B200VA_K_AUTO;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--lo") && i + 1 < argc) lo = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--hi") && i + 1 < argc) hi = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--kernel") && i + 1 < argc) {
            const char* k = argv[++i];
            variant = !std::strcmp(k, "k0") ? 1 : !std::strcmp(k, "k1") ? 2 : !std::strcmp(k, "k2") ? 3 : !std::strcmp(k, "k3") ? 4 : 0;
        } else { std::fprintf(stderr, "usage: %s [--lo 16] [--hi 30] [--kernel auto|k0|k1|k2|k3]\n", argv[0]); return 2; }
    }
    b200va_devinfo_t di;
    VA(b200va_query(0, &di));
    CK(cudaSetDevice(0));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    uint64_t* dRes;
    CK(cudaMalloc(&dRes, 2 * sizeof(uint64_t)));
    const size_t cold_bytes = 4 * static_cast<size_t>(di.l2_bytes);

    for (int k = lo; k <= hi; ++k) {
        const size_t n = size_t{1} << k;
        const size_t set_bytes = 12 * n;
        const int sets = static_cast<int>(std::max<size_t>(1, (cold_bytes + set_bytes - 1) / set_bytes));
        float* pool = nullptr;
        CK(cudaMalloc(&pool, static_cast<size_t>(sets) * 3 * n * sizeof(float)));
        for (int s = 0; s < sets; ++s) {
            VA(b200va_fill_ctr_f32(pool + (static_cast<size_t>(s) * 3 + 0) * n, n, 0x0A, 0, st));
            VA(b200va_fill_ctr_f32(pool + (static_cast<size_t>(s) * 3 + 1) * n, n, 0x0B, 0, st));
        }
        auto A = [&](int s) { return pool + (static_cast<size_t>(s) * 3 + 0) * n; };
        auto B = [&](int s) { return pool + (static_cast<size_t>(s) * 3 + 1) * n; };
        auto C = [&](int s) { return pool + (static_cast<size_t>(s) * 3 + 2) * n; };

        // launches per timing sample: ~20 ms of work, at least 20, at most 2000
        const double est_us = std::max(3.0, 12.0 * static_cast<double>(n) / 7.0e6);
        const int iters = static_cast<int>(std::min(2000.0, std::max(20.0, 20000.0 / est_us)));
        const int samples = 7;

        auto time_batches = [&](auto&& body) {
            std::vector<double> us;
            for (int sidx = 0; sidx < samples + 1; ++sidx) {
                CK(cudaEventRecord(e0, st));
                body();
                CK(cudaEventRecord(e1, st));
                CK(cudaStreamSynchronize(st));
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (sidx) us.push_back(1e3 * ms / iters);   // first batch = warm-up
            }
            return median(us);
        };

        const double hot = time_batches([&] { for (int i = 0; i < iters; ++i) VA(b200va_add_f32(A(0), B(0), C(0), n, variant, st)); });
        const double cold = sets > 1 ? time_batches([&] { for (int i = 0; i < iters; ++i) { const int s = i % sets; VA(b200va_add_f32_ex(A(s), B(s), C(s), n, variant, B200VA_F_COLD, st)); } })
                                     : hot;
        const double cold_chain = time_batches([&] { for (int i = 0; i < iters; ++i) { const int s = sets > 1 ? i % sets : 0; VA(b200va_add_f32_ex(A(s), B(s), C(s), n, variant, B200VA_F_COLD | B200VA_F_INPUTS_STABLE, st)); } });
        b200va_loop_t* loop = nullptr;
        VA(b200va_loop_create(&loop, A(0), B(0), C(0), n, variant, 100));
        const double graph = time_batches([&] { VA(b200va_loop_run(loop, iters, st)); });
        CK(cudaStreamSynchronize(st));
        VA(b200va_loop_destroy(loop));
        const double k0 = time_batches([&] { for (int i = 0; i < iters; ++i) VA(b200va_add_f32(A(0), B(0), C(0), n, B200VA_K0_SCALAR, st)); });

        // bit-exactness of the last cold launch's set
        uint64_t h[2];
        VA(b200va_verify_f32(A(0), B(0), C(0), n, dRes, st));
        CK(cudaMemcpyAsync(h, dRes, sizeof h, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));

        b200va_tune_t t;
        VA(b200va_resolve(variant, n, &t));
        auto gbps = [&](double us) { return 12.0 * static_cast<double>(n) / us / 1e3; };
        std::printf("{\"log2_n\": %d, \"n\": %zu, \"algorithmic_bytes\": %zu, \"l2_resident\": %s, \"buffer_sets\": %d, "
                    "\"launches_per_sample\": %d, \"kind\": %d, \"threads\": %d, \"unroll\": %d, \"mismatches\": %llu, "
                    "\"us_hot\": %.3f, \"us_cold\": %.3f, \"us_cold_chain\": %.3f, \"us_graph\": %.3f, \"us_k0\": %.3f, "
                    "\"GBps_hot\": %.1f, \"GBps_cold\": %.1f, \"GBps_cold_chain\": %.1f, \"GBps_graph\": %.1f, \"GBps_k0\": %.1f, "
                    "\"elems_per_s_cold\": %.4e, \"elems_per_s_graph\": %.4e}\n",
                    k, n, set_bytes, set_bytes <= static_cast<size_t>(di.l2_bytes) ? "true" : "false", sets, iters, t.kind,
                    t.threads, t.unroll, static_cast<unsigned long long>(h[0]), hot, cold, cold_chain, graph, k0, gbps(hot), gbps(cold),
                    gbps(cold_chain), gbps(graph), gbps(k0), static_cast<double>(n) / (cold * 1e-6), static_cast<double>(n) / (graph * 1e-6));
        std::fflush(stdout);
        CK(cudaFree(pool));
    }
    return 0;
}
