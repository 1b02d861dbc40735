// b200va_kernels.cuh -- the sm_100a kernels of the vectorAdd hot path.
//
// What they replace: the image's `vectorAdd` kernel (C[i] = A[i] + B[i], one element
// per thread, 256-thread CTAs), which the reference only invokes
// (cuda-test-deployment.yaml:18-19).
//
// Roofline: pure HBM stream, 12 algorithmic bytes per element (4 read A + 4 read B +
// 4 write C), 1 FLOP.  No data reuse, so no tensor cores and no smem tiling for reuse;
// shared memory appears only as the landing zone of the TMA ring (K2).
//
//   K0  vadd_scalar      reference-shape control (scalar LDG/STG).
//   K1  vadd_vec<4,...>  128-bit LDG/STG, UNROLL independent vectors per thread per
//                        array in flight, cache-hinted, tile-strided (optionally
//                        persistent).
//   K3  vadd_vec<8,...>  same with 256-bit LDG.E.256/STG.E.256 (PTX 8.8, sm_100).
//   K1c vadd_vec_clc     the K1/K3 tile body with Blackwell's cluster-launch-control
//                        scheduler: resident CTAs cancel not-yet-started CTAs of the grid
//                        and take over their tiles (no CTA relaunch per tile).
//   EARLY (K1/K3/K1c)    what a CTA does about its first tile BEFORE the programmatic dependency
//                        on the previous launch resolves (griddepcontrol.wait):
//                        1  issues the loads themselves; only the stores wait.  Legal when the
//                           previous launch on the stream does not write A or B (the launch
//                           loop, a1: the previous launch is the same add, which writes only C).
//                        2  one thread bulk-prefetches the A and B tiles into L2
//                           (cp.async.bulk.prefetch.L2), loads and stores wait.  ALWAYS legal:
//                           L2 is the coherence point, a line prefetched early and then written
//                           by the previous launch is simply up to date when it is read.
//                        Both overlap this launch's DRAM ramp with the previous launch's tail.
//   K2  vadd_tma         persistent CTAs; one producer lane issues 1-D cp.async.bulk
//                        copies of an A tile and a B tile into a `stages`-deep smem ring
//                        (mbarrier complete_tx); consumer warps add from smem and either
//                        store from registers or write back in place and bulk-store.
//   K2c vadd_tma_clc     the same ring with Blackwell's cluster-launch-control scheduler:
//                        one CTA per tile in the grid, resident CTAs cancel and take over
//                        the not-yet-started ones (dynamic balance, persistent ring).
//
// Ragged sizes / alignment: pointers need only 4-byte alignment.  `head` scalar
// elements are peeled so the vector body is 16/32-byte aligned, the < VW tail is
// scalar; both are done by CTA 0 of the same launch (one launch per call, always).
// If A, B and C share no 16-byte phase there is no vector body: vadd_scalar_unrolled
// keeps 16 independent 4-byte loads per thread in flight instead (7.19 TB/s at 2^28 --
// bytes in flight, not access width, is what saturates the HBM).
#pragma once
#include <cstddef>
#include <cstdint>

#include "b200va_ptx.cuh"

namespace b200va {

// ------------------------------------------------------------------------------ K0
__global__ void vadd_scalar(const float* A, const float* B, float* C, size_t n)
{
    const size_t i = static_cast<size_t>(blockDim.x) * blockIdx.x + threadIdx.x;
    if (i < n) C[i] = __fadd_rn(A[i], B[i]);
}

// Fallback when A, B and C are misaligned *differently* (no common 16-byte phase, so no
// vector body exists): 4-byte accesses, but U independent coalesced loads per array per
// thread in flight instead of the control's one.  Tile = blockDim.x * U elements.
template <int U>
__global__ void vadd_scalar_unrolled(const float* A, const float* B, float* C, size_t n)
{
    const size_t base = static_cast<size_t>(blockIdx.x) * blockDim.x * U + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    float a[U], b[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const size_t i = base + static_cast<size_t>(j) * blockDim.x;
        a[j] = i < n ? A[i] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const size_t i = base + static_cast<size_t>(j) * blockDim.x;
        b[j] = i < n ? B[i] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const size_t i = base + static_cast<size_t>(j) * blockDim.x;
        if (i < n) C[i] = __fadd_rn(a[j], b[j]);
    }
}

// --------------------------------------------------------------------------- K1/K3
template <int VW> struct vec_t;
template <> struct vec_t<4> { using type = f32x4; };
template <> struct vec_t<8> { using type = f32x8; };

template <int VW, int LD>
__device__ __forceinline__ typename vec_t<VW>::type ldg_vec(const float* p, uint64_t pol)
{
    if constexpr (VW == 4) return ldg128<LD>(p, pol);
    else return ldg256<LD>(p, pol);
}

template <int VW, int ST>
__device__ __forceinline__ void stg_vec(float* p, const typename vec_t<VW>::type& v, uint64_t pol)
{
    if constexpr (VW == 4) stg128<ST>(p, v, pol);
    else stg256<ST>(p, v, pol);
}

__device__ __forceinline__ f32x4 add_vec(const f32x4& a, const f32x4& b)
{
    return f32x4{__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w)};
}

__device__ __forceinline__ f32x8 add_vec(const f32x8& a, const f32x8& b)
{
    f32x8 r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = __fadd_rn(a.v[k], b.v[k]);
    return r;
}

// Scalar head [0, head) and tail [tail0, n), done by the first threads of one CTA.
__device__ __forceinline__ void add_edges(const float* A, const float* B, float* C, size_t n,
                                          size_t head, size_t tail0, unsigned t)
{
    if (t < head) C[t] = __fadd_rn(A[t], B[t]);
    if (tail0 + t < n) C[tail0 + t] = __fadd_rn(A[tail0 + t], B[tail0 + t]);
}

// Tile = blockDim.x * UNROLL vectors; thread t owns vectors t + j*blockDim.x of the tile,
// so each warp-level access is one contiguous 512 B (VW=4) or 1 KiB (VW=8) run.  All
// 2*UNROLL loads are issued before the first add: that is the memory-level parallelism.
template <int VW, int UNROLL, int LD, int ST>
__device__ __forceinline__ void vec_tile(const float* a, const float* b, float* c, size_t tile, size_t tile_vecs,
                                         size_t nvec, uint64_t pol, bool& waited)
{
    using V = typename vec_t<VW>::type;
    const size_t v0 = tile * tile_vecs + threadIdx.x;
    if ((tile + 1) * tile_vecs <= nvec) {
        V ra[UNROLL], rb[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j)
            ra[j] = ldg_vec<VW, LD>(a + (v0 + static_cast<size_t>(j) * blockDim.x) * VW, pol);
#pragma unroll
        for (int j = 0; j < UNROLL; ++j)
            rb[j] = ldg_vec<VW, LD>(b + (v0 + static_cast<size_t>(j) * blockDim.x) * VW, pol);
        if (!waited) { pdl_wait(); waited = true; }      // EARLY: only the stores wait for the previous launch
#pragma unroll
        for (int j = 0; j < UNROLL; ++j)
            stg_vec<VW, ST>(c + (v0 + static_cast<size_t>(j) * blockDim.x) * VW,
                            add_vec(ra[j], rb[j]), pol);
    } else {
        if (!waited) { pdl_wait(); waited = true; }
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            const size_t v = v0 + static_cast<size_t>(j) * blockDim.x;
            if (v < nvec) {
                const V x = ldg_vec<VW, LD>(a + v * VW, pol);
                const V y = ldg_vec<VW, LD>(b + v * VW, pol);
                stg_vec<VW, ST>(c + v * VW, add_vec(x, y), pol);
            }
        }
    }
}

// EARLY == 2: thread 0 asks the L2 to fetch the CTA's first A and B tiles (two bulk-prefetch
// instructions) while the previous launch is still draining.
template <int VW>
__device__ __forceinline__ void prefetch_first_tile(const float* a, const float* b, size_t tile, size_t tile_vecs, size_t nvec)
{
    if (threadIdx.x != 0 || tile * tile_vecs >= nvec) return;
    const size_t v0 = tile * tile_vecs;
    const size_t nv = nvec - v0 < tile_vecs ? nvec - v0 : tile_vecs;
    const uint32_t bytes = static_cast<uint32_t>(nv * VW * sizeof(float));      // a multiple of 16, <= 256 KiB
    bulk_prefetch_l2(a + v0 * VW, bytes);
    bulk_prefetch_l2(b + v0 * VW, bytes);
}

template <int VW, int UNROLL, int LD, int ST, int EARLY>
__global__ void vadd_vec(const float* A, const float* B, float* C, size_t n, size_t head,
                         size_t nvec, size_t ntiles)
{
    uint64_t pol = 0;
    if constexpr (LD == LD_NA_EF || ST == ST_NA_EF) pol = l2_evict_first_policy();

    const float* a = A + head;
    const float* b = B + head;
    float* c = C + head;
    const size_t tile_vecs = static_cast<size_t>(blockDim.x) * UNROLL;
    pdl_launch_dependents();
    if constexpr (EARLY == 2) prefetch_first_tile<VW>(a, b, blockIdx.x, tile_vecs, nvec);
    bool waited = EARLY != 1;
    if constexpr (EARLY != 1) pdl_wait();   // everything above overlapped the previous launch's tail

    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        vec_tile<VW, UNROLL, LD, ST>(a, b, c, tile, tile_vecs, nvec, pol, waited);
    if (!waited) pdl_wait();
    if (blockIdx.x == 0) add_edges(A, B, C, n, head, head + nvec * VW, threadIdx.x);
}

// ------------------------------------------------------------------------------ K1c
// The vec tile body under the cluster-launch-control scheduler.  Grid = one CTA per tile;
// a running CTA, before it touches its tile, asks the hardware to cancel one CTA that has
// not started yet (try_cancel, answered through the async proxy into shared memory and
// counted on an mbarrier) and -- after its own tile -- processes the cancelled CTA's tile
// instead of exiting.  The request's latency hides behind the tile's loads.  One response
// slot: every thread reads it, the CTA barrier orders those reads before thread 0 re-arms it.
template <int VW, int UNROLL, int LD, int ST, int EARLY>
__global__ void vadd_vec_clc(const float* A, const float* B, float* C, size_t n, size_t head,
                             size_t nvec, size_t ntiles)
{
    __shared__ __align__(16) unsigned char resp[16];
    __shared__ __align__(8) unsigned long long bar_storage;
    const uint32_t resp_s = smem_u32(resp), bar = smem_u32(&bar_storage);
    uint64_t pol = 0;
    if constexpr (LD == LD_NA_EF || ST == ST_NA_EF) pol = l2_evict_first_policy();

    if (threadIdx.x == 0) {
        mbar_init(bar, 1u);
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    __syncthreads();

    const float* a = A + head;
    const float* b = B + head;
    float* c = C + head;
    const size_t tile_vecs = static_cast<size_t>(blockDim.x) * UNROLL;
    pdl_launch_dependents();
    if constexpr (EARLY == 2) prefetch_first_tile<VW>(a, b, blockIdx.x, tile_vecs, nvec);
    bool waited = EARLY != 1;
    if constexpr (EARLY != 1) pdl_wait();

    uint32_t tile = blockIdx.x, ph = 0;
    while (true) {
        if (threadIdx.x == 0) {                           // ask for the tile after this one
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(bar, 16u);
            clc_try_cancel(resp_s, bar);
        }
        if (tile < ntiles) vec_tile<VW, UNROLL, LD, ST>(a, b, c, tile, tile_vecs, nvec, pol, waited);
        if (tile == 0) {                                  // edges travel with tile 0, whoever runs it
            if (!waited) { pdl_wait(); waited = true; }
            add_edges(A, B, C, n, head, head + nvec * VW, threadIdx.x);
        }
        mbar_wait(bar, ph);
        ph ^= 1u;
        uint32_t next = 0;
        const bool ok = clc_query(resp_s, next);
        __syncthreads();                                  // all reads of the slot precede its re-arm
        if (!ok) break;
        tile = next;
    }
    if (!waited) pdl_wait();
}

// ------------------------------------------------------------------------------ K2
// Shared memory: [stages][A tile | B tile] then full[stages], empty[stages] mbarriers.
// Warp 0 lane 0 = producer, warps 1.. = consumers (blockDim.x - 32 threads).
//   STORE_MODE 0: consumers ld.shared both tiles, add, st.global.v4 from registers;
//                 every consumer warp arrives on empty[s] when done reading.
//   STORE_MODE 1: consumers write the sum in place over the A tile, fence to the async
//                 proxy, and consumer thread 0 bulk-stores the tile (UBLKCP.G.S); the
//                 stage is released once that store has finished reading shared memory
//                 (cp.async.bulk.wait_group.read), one tile later.
template <int STORE_MODE, bool L2_HINT, int ST>
__global__ void __launch_bounds__(1024, 1)
vadd_tma(const float* A, const float* B, float* C, size_t n, size_t head, size_t nvec4,
         uint32_t tile_bytes, uint32_t stages)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + stages * 2u * tile_bytes;
    const uint32_t n_cons = blockDim.x - 32u;
    const uint32_t n_cons_warps = n_cons >> 5;

    auto full_bar = [&](uint32_t s) { return bar_base + s * 8u; };
    auto empty_bar = [&](uint32_t s) { return bar_base + (stages + s) * 8u; };
    auto tile_a = [&](uint32_t s) { return smem_base + s * 2u * tile_bytes; };
    auto tile_b = [&](uint32_t s) { return smem_base + s * 2u * tile_bytes + tile_bytes; };

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < stages; ++s) {
            mbar_init(full_bar(s), 1u);
            mbar_init(empty_bar(s), STORE_MODE == 0 ? n_cons_warps : 1u);
        }
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    __syncthreads();

    const unsigned char* a = reinterpret_cast<const unsigned char*>(A + head);
    const unsigned char* b = reinterpret_cast<const unsigned char*>(B + head);
    unsigned char* c = reinterpret_cast<unsigned char*>(C + head);
    const size_t body_bytes = nvec4 * 16u;
    const size_t ntiles = (body_bytes + tile_bytes - 1) / tile_bytes;

    uint64_t pol = 0;
    if constexpr (L2_HINT || ST == ST_NA_EF) pol = l2_evict_first_policy();
    pdl_launch_dependents();
    pdl_wait();   // barrier init and address math overlapped the previous launch's tail

    if (threadIdx.x < 32) {
        // ------------------------------------------------------------ producer
        if (threadIdx.x == 0) {
            uint32_t s = 0, ph = 0;
            for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                mbar_wait(empty_bar(s), ph ^ 1u);
                const size_t off = t * tile_bytes;
                const size_t left = body_bytes - off;
                const uint32_t bytes = left < tile_bytes ? static_cast<uint32_t>(left) : tile_bytes;
                mbar_arrive_expect_tx(full_bar(s), 2u * bytes);
                bulk_g2s<L2_HINT>(tile_a(s), a + off, bytes, full_bar(s), pol);
                bulk_g2s<L2_HINT>(tile_b(s), b + off, bytes, full_bar(s), pol);
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------ consumers
        const uint32_t ct = threadIdx.x - 32u;
        uint32_t s = 0, ph = 0, prev_s = 0;
        bool have_prev = false;
        for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            mbar_wait(full_bar(s), ph);
            const size_t off = t * tile_bytes;
            const size_t left = body_bytes - off;
            const uint32_t bytes = left < tile_bytes ? static_cast<uint32_t>(left) : tile_bytes;
            const uint32_t nv = bytes >> 4;
            const uint32_t sa = tile_a(s), sb = tile_b(s);

            uint32_t v = ct;
            for (; v + 3u * n_cons < nv; v += 4u * n_cons) {
                f32x4 x[4], y[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = lds128(sa + (v + j * n_cons) * 16u);
#pragma unroll
                for (int j = 0; j < 4; ++j) y[j] = lds128(sb + (v + j * n_cons) * 16u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const f32x4 r = add_vec(x[j], y[j]);
                    if constexpr (STORE_MODE == 0)
                        stg128<ST>(reinterpret_cast<float*>(c + off) + (v + j * n_cons) * 4u, r, pol);
                    else
                        sts128(sa + (v + j * n_cons) * 16u, r);
                }
            }
            for (; v < nv; v += n_cons) {
                const f32x4 r = add_vec(lds128(sa + v * 16u), lds128(sb + v * 16u));
                if constexpr (STORE_MODE == 0)
                    stg128<ST>(reinterpret_cast<float*>(c + off) + v * 4u, r, pol);
                else
                    sts128(sa + v * 16u, r);
            }

            if constexpr (STORE_MODE == 0) {
                __syncwarp();
                if ((threadIdx.x & 31u) == 0) mbar_arrive(empty_bar(s));
            } else {
                fence_proxy_async_smem();
                named_bar_sync(1, n_cons);
                if (ct == 0) {
                    bulk_s2g<L2_HINT>(c + off, sa, bytes, pol);
                    bulk_commit();
                    if (have_prev) {
                        bulk_wait_read<1>();
                        mbar_arrive(empty_bar(prev_s));
                    }
                }
                prev_s = s;
                have_prev = true;
            }
            if (++s == stages) { s = 0; ph ^= 1u; }
        }
        if constexpr (STORE_MODE == 1) {
            if (ct == 0) bulk_wait_all<0>();
        }
        if (blockIdx.x == 0) add_edges(A, B, C, n, head, head + nvec4 * 4u, ct);
    }
}

// ------------------------------------------------------------------------------ K2c
// The TMA ring with Blackwell's cluster-launch-control tile scheduler.  The grid has one
// CTA per tile, but only as many CTAs as fit are ever resident: each resident CTA's
// producer lane, besides streaming its own tile through the ring, keeps `stages`
// try_cancel requests in flight; every success hands it the index of a CTA that has not
// started yet, whose tile it then loads into the next free stage.  Work is therefore
// balanced dynamically across SMs (a static persistent split costs ~10 % here because the
// SMs do not all see the same memory latency) while the ring, the barriers and the CTA
// stay alive.  Consumers learn each stage's tile index from shared memory; a sentinel
// ends them.  Register stores only (STORE_MODE 0 of vadd_tma).
template <bool L2_HINT, int ST>
__global__ void __launch_bounds__(1024, 1)
vadd_tma_clc(const float* A, const float* B, float* C, size_t n, size_t head, size_t nvec4,
             uint32_t tile_bytes, uint32_t stages)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + stages * 2u * tile_bytes;       // full[], empty[], clc_bar[]
    const uint32_t clc_base = (bar_base + 3u * stages * 8u + 15u) & ~15u; // 16-byte aligned responses
    const uint32_t slot_base = clc_base + stages * 16u;                   // uint32 tile index per stage
    const uint32_t n_cons = blockDim.x - 32u;
    const uint32_t n_cons_warps = n_cons >> 5;
    constexpr uint32_t kDone = 0xffffffffu;

    auto full_bar = [&](uint32_t s) { return bar_base + s * 8u; };
    auto empty_bar = [&](uint32_t s) { return bar_base + (stages + s) * 8u; };
    auto clc_bar = [&](uint32_t s) { return bar_base + (2u * stages + s) * 8u; };
    auto clc_resp = [&](uint32_t s) { return clc_base + s * 16u; };
    auto tile_a = [&](uint32_t s) { return smem_base + s * 2u * tile_bytes; };
    auto tile_b = [&](uint32_t s) { return smem_base + s * 2u * tile_bytes + tile_bytes; };
    volatile uint32_t* tile_slot = reinterpret_cast<volatile uint32_t*>(smem + (slot_base - smem_base));

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < stages; ++s) {
            mbar_init(full_bar(s), 1u);
            mbar_init(empty_bar(s), n_cons_warps);
            mbar_init(clc_bar(s), 1u);
        }
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    __syncthreads();

    const unsigned char* a = reinterpret_cast<const unsigned char*>(A + head);
    const unsigned char* b = reinterpret_cast<const unsigned char*>(B + head);
    unsigned char* c = reinterpret_cast<unsigned char*>(C + head);
    const size_t body_bytes = nvec4 * 16u;

    uint64_t pol = 0;
    if constexpr (L2_HINT || ST == ST_NA_EF) pol = l2_evict_first_policy();
    pdl_launch_dependents();
    pdl_wait();

    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) {
            // ---------------------------------------------------- producer + scheduler
            for (uint32_t j = 0; j < stages; ++j) {               // `stages` requests in flight
                mbar_arrive_expect_tx(clc_bar(j), 16u);
                clc_try_cancel(clc_resp(j), clc_bar(j));
            }
            uint32_t outstanding = stages, consumed = 0;
            bool failed = false;
            uint32_t tile = blockIdx.x, s = 0, ph = 0;
            while (body_bytes != 0) {                             // (n < 4: only the scalar edges exist)
                mbar_wait(empty_bar(s), ph ^ 1u);
                const size_t off = static_cast<size_t>(tile) * tile_bytes;
                const size_t left = body_bytes - off;
                const uint32_t bytes = left < tile_bytes ? static_cast<uint32_t>(left) : tile_bytes;
                tile_slot[s] = tile;                              // published by the arrive below (release)
                mbar_arrive_expect_tx(full_bar(s), 2u * bytes);
                bulk_g2s<L2_HINT>(tile_a(s), a + off, bytes, full_bar(s), pol);
                bulk_g2s<L2_HINT>(tile_b(s), b + off, bytes, full_bar(s), pol);
                if (++s == stages) { s = 0; ph ^= 1u; }

                // next tile: the oldest outstanding response; after the first failure nothing
                // is re-armed, but every outstanding response is still drained before exit
                bool have_next = false;
                while (outstanding > 0 && !have_next) {
                    const uint32_t j = consumed % stages;
                    mbar_wait(clc_bar(j), (consumed / stages) & 1u);
                    uint32_t next = 0;
                    const bool ok = clc_query(clc_resp(j), next);
                    ++consumed;
                    --outstanding;
                    if (ok) {
                        tile = next;
                        have_next = true;
                        if (!failed) {
                            fence_proxy_async_smem();             // the generic-proxy read of the slot precedes its async re-write
                            mbar_arrive_expect_tx(clc_bar(j), 16u);
                            clc_try_cancel(clc_resp(j), clc_bar(j));
                            ++outstanding;
                        }
                    } else {
                        failed = true;
                    }
                }
                if (!have_next) break;
            }
            while (outstanding > 0) {                             // (only reached with requests left when n < 4)
                mbar_wait(clc_bar(consumed % stages), (consumed / stages) & 1u);
                ++consumed;
                --outstanding;
            }
            mbar_wait(empty_bar(s), ph ^ 1u);                     // sentinel: no more tiles
            tile_slot[s] = kDone;
            mbar_arrive(full_bar(s));
        }
    } else {
        // ------------------------------------------------------------ consumers
        const uint32_t ct = threadIdx.x - 32u;
        uint32_t s = 0, ph = 0;
        while (true) {
            mbar_wait(full_bar(s), ph);
            const uint32_t tile = tile_slot[s];
            if (tile == kDone) break;
            const size_t off = static_cast<size_t>(tile) * tile_bytes;
            const size_t left = body_bytes - off;
            const uint32_t bytes = left < tile_bytes ? static_cast<uint32_t>(left) : tile_bytes;
            const uint32_t nv = bytes >> 4;
            const uint32_t sa = tile_a(s), sb = tile_b(s);
            float* cg = reinterpret_cast<float*>(c + off);

            uint32_t v = ct;
            for (; v + 3u * n_cons < nv; v += 4u * n_cons) {
                f32x4 x[4], y[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = lds128(sa + (v + j * n_cons) * 16u);
#pragma unroll
                for (int j = 0; j < 4; ++j) y[j] = lds128(sb + (v + j * n_cons) * 16u);
#pragma unroll
                for (int j = 0; j < 4; ++j) stg128<ST>(cg + (v + j * n_cons) * 4u, add_vec(x[j], y[j]), pol);
            }
            for (; v < nv; v += n_cons)
                stg128<ST>(cg + v * 4u, add_vec(lds128(sa + v * 16u), lds128(sb + v * 16u)), pol);
            // the scalar head/tail travel with tile 0, whichever CTA ends up running it (CTA 0 may be
            // cancelled and its tile taken over under cluster launch control)
            if (tile == 0) add_edges(A, B, C, n, head, head + nvec4 * 4u, ct);

            __syncwarp();
            if ((threadIdx.x & 31u) == 0) mbar_arrive(empty_bar(s));
            if (++s == stages) { s = 0; ph ^= 1u; }
        }
        // no vector body at all (n < 4 after the head): the grid is the single CTA 0
        if (body_bytes == 0 && blockIdx.x == 0) add_edges(A, B, C, n, head, head + nvec4 * 4u, ct);
    }
}

// ---------------------------------------------------------------- ceiling probes
// The production geometry (512 threads x one 128-bit vector per array) with one side of the
// traffic removed: what the HBM gives a pure read stream, a pure write stream and a 1:1 copy,
// next to the add's 2:1 mix (b200va_probe_f32; profiles/r02/README.md section 1a).
__global__ void probe_read2(const float* A, const float* B, float* C, size_t nvec)
{
    pdl_launch_dependents();
    pdl_wait();
    const size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    float sum = 0.f;
    if (v < nvec) {
        const f32x4 a = ldg128<LD_PLAIN>(A + v * 4, 0), b = ldg128<LD_PLAIN>(B + v * 4, 0);   // asm volatile: never elided
        sum = a.x + b.x + a.y + b.y + a.z + b.z + a.w + b.w;
    }
    if (threadIdx.x == 0) C[blockIdx.x] = sum;          // 4 bytes per CTA: 0.02 % of the traffic
}

__global__ void probe_fill(float* C, size_t nvec, float value)
{
    pdl_launch_dependents();
    pdl_wait();
    const size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v < nvec) stg128<ST_NA>(C + v * 4, f32x4{value, value, value, value}, 0);
}

__global__ void probe_copy(const float* A, float* C, size_t nvec)
{
    pdl_launch_dependents();
    pdl_wait();
    const size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v < nvec) stg128<ST_NA>(C + v * 4, ldg128<LD_PLAIN>(A + v * 4, 0), 0);
}

// ---------------------------------------------------------------- support kernels
__device__ __forceinline__ uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// a2 for the large configs: x[i] = (float)(splitmix64(base + i) >> 40) * 2^-24.
// A 24-bit integer converts to binary32 exactly and the scale is a power of two, so the
// device values are bit-identical to the host generator by construction.
__global__ void fill_ctr(float* x, size_t n, uint64_t base)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        x[i] = __uint2float_rn(static_cast<uint32_t>(splitmix64(base + i) >> 40)) * 0x1.0p-24f;
}

// 128-bit form for 16-byte-aligned x: thread v generates elements 4v..4v+3 and stores them with one STG.128;
// the < 4-element tail is written by the first threads of CTA 0.  Same values as fill_ctr by construction.
__global__ void fill_ctr_vec(float* x, size_t n, uint64_t base)
{
    const size_t nvec = n / 4, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        f32x4 r;
        r.x = __uint2float_rn(static_cast<uint32_t>(splitmix64(base + 4 * v + 0) >> 40)) * 0x1.0p-24f;
        r.y = __uint2float_rn(static_cast<uint32_t>(splitmix64(base + 4 * v + 1) >> 40)) * 0x1.0p-24f;
        r.z = __uint2float_rn(static_cast<uint32_t>(splitmix64(base + 4 * v + 2) >> 40)) * 0x1.0p-24f;
        r.w = __uint2float_rn(static_cast<uint32_t>(splitmix64(base + 4 * v + 3) >> 40)) * 0x1.0p-24f;
        stg128<ST_NA>(x + 4 * v, r, 0);
    }
    if (blockIdx.x == 0 && 4 * nvec + threadIdx.x < n)
        x[4 * nvec + threadIdx.x] = __uint2float_rn(static_cast<uint32_t>(splitmix64(base + 4 * nvec + threadIdx.x) >> 40)) * 0x1.0p-24f;
}

__device__ __forceinline__ bool is_nan_bits(uint32_t u) { return (u & 0x7fffffffu) > 0x7f800000u; }

__device__ __forceinline__ bool bits_differ(float want, float got)
{
    const uint32_t w = __float_as_uint(want), g = __float_as_uint(got);
    return w != g && !(is_nan_bits(w) && is_nan_bits(g));
}

// a6 in HBM: result[0] += #mismatches, result[1] = min mismatching index.
__global__ void verify_bits(const float* A, const float* B, const float* C, size_t n,
                            unsigned long long* result)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    unsigned long long bad = 0, first = ~0ull;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t want = __float_as_uint(__fadd_rn(A[i], B[i]));
        const uint32_t got = __float_as_uint(C[i]);
        if (want != got && !(is_nan_bits(want) && is_nan_bits(got))) {
            ++bad;
            if (i < first) first = i;
        }
    }
    if (bad) {
        atomicAdd(&result[0], bad);
        atomicMin(&result[1], first);
    }
}

// a6 in HBM, 128-bit form (A, B, C 16-byte aligned): U = 2 vectors per array per thread in flight, so the
// check streams at the add's own rate instead of a third of it.  Same verdict as verify_bits.
__global__ void verify_bits_vec(const float* A, const float* B, const float* C, size_t n, unsigned long long* result)
{
    constexpr int U = 2;
    const size_t nvec = n / 4, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    unsigned long long bad = 0, first = ~0ull;
    auto check = [&](const f32x4& a, const f32x4& b, const f32x4& c, size_t v) {
        const float w[4] = {__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w)};
        const float g[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int k = 3; k >= 0; --k)
            if (bits_differ(w[k], g[k])) { ++bad; if (4 * v + k < first) first = 4 * v + k; }
    };
    size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; v + (U - 1) * stride < nvec; v += U * stride) {
        f32x4 a[U], b[U], c[U];
#pragma unroll
        for (int j = 0; j < U; ++j) a[j] = ldg128<LD_PLAIN>(A + 4 * (v + j * stride), 0);
#pragma unroll
        for (int j = 0; j < U; ++j) b[j] = ldg128<LD_PLAIN>(B + 4 * (v + j * stride), 0);
#pragma unroll
        for (int j = 0; j < U; ++j) c[j] = ldg128<LD_PLAIN>(C + 4 * (v + j * stride), 0);
#pragma unroll
        for (int j = 0; j < U; ++j) check(a[j], b[j], c[j], v + j * stride);
    }
    for (; v < nvec; v += stride) check(ldg128<LD_PLAIN>(A + 4 * v, 0), ldg128<LD_PLAIN>(B + 4 * v, 0), ldg128<LD_PLAIN>(C + 4 * v, 0), v);
    if (blockIdx.x == 0) {
        const size_t i = 4 * nvec + threadIdx.x;
        if (i < n && bits_differ(__fadd_rn(A[i], B[i]), C[i])) { ++bad; if (i < first) first = i; }
    }
    if (bad) {
        atomicAdd(&result[0], bad);
        atomicMin(&result[1], first);
    }
}

// 128-bit digest (X 16-byte aligned), 4 vectors per thread in flight.  Sum and xor are order-independent.
__global__ void digest_bits_vec(const float* X, size_t n, unsigned long long* out)
{
    constexpr int U = 4;
    const size_t nvec = n / 4, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    unsigned long long s = 0;
    uint32_t x = 0;
    auto take = [&](const f32x4& r) {
        const uint32_t u0 = __float_as_uint(r.x), u1 = __float_as_uint(r.y), u2 = __float_as_uint(r.z), u3 = __float_as_uint(r.w);
        s += static_cast<unsigned long long>(u0) + u1 + u2 + u3;
        x ^= u0 ^ u1 ^ u2 ^ u3;
    };
    size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; v + (U - 1) * stride < nvec; v += U * stride) {
        f32x4 r[U];
#pragma unroll
        for (int j = 0; j < U; ++j) r[j] = ldg128<LD_PLAIN>(X + 4 * (v + j * stride), 0);
#pragma unroll
        for (int j = 0; j < U; ++j) take(r[j]);
    }
    for (; v < nvec; v += stride) take(ldg128<LD_PLAIN>(X + 4 * v, 0));
    if (blockIdx.x == 0 && 4 * nvec + threadIdx.x < n) {
        const uint32_t u = __float_as_uint(X[4 * nvec + threadIdx.x]);
        s += u;
        x ^= u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        x ^= __shfl_xor_sync(0xffffffffu, x, o);
    }
    if ((threadIdx.x & 31u) == 0) {
        atomicAdd(&out[0], s);
        atomicXor(&out[1], static_cast<unsigned long long>(x));
    }
}

// out[0] += sum of uint32 patterns, out[1] ^= xor of patterns.
__global__ void digest_bits(const float* X, size_t n, unsigned long long* out)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    unsigned long long s = 0;
    uint32_t x = 0;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t u = __float_as_uint(X[i]);
        s += u;
        x ^= u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        x ^= __shfl_xor_sync(0xffffffffu, x, o);
    }
    if ((threadIdx.x & 31u) == 0) {
        atomicAdd(&out[0], s);
        atomicXor(&out[1], static_cast<unsigned long long>(x));
    }
}

__global__ void reset_verify(unsigned long long* result)
{
    result[0] = 0ull;
    result[1] = ~0ull;
}

__global__ void reset_digest(unsigned long long* out)
{
    out[0] = 0ull;
    out[1] = 0ull;
}

}  // namespace b200va
