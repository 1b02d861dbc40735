// b200va_ptx.cuh -- hand-written PTX primitives for the sm_100a vectorAdd kernels.
//
// Everything here is a thin wrapper over one PTX instruction so that the SASS the
// kernels produce is predictable (LDG.E.128 / LDG.E.256, UBLKCP, SYNCS.*): no CCCL,
// no CUTLASS.  Compiled only for -gencode arch=compute_100a,code=sm_100a.
#pragma once
#include <cstdint>

namespace b200va {

// ---------------------------------------------------------------- cache-hint enums
// Load hints (tune.ld_hint)
enum : int {
    LD_PLAIN = 0,   // ld.global
    LD_NA    = 1,   // ld.global.L1::no_allocate               (streaming: skip L1)
    LD_CS    = 2,   // ld.global.cs                             (evict-first streaming)
    LD_NA_EF = 3,   // L1::no_allocate + L2 evict_first         (policy operand / .L2::evict_first on 256-bit)
    LD_NC_NA = 4,   // ld.global.nc.L1::no_allocate             (read-only path; NOTE ptxas is then free to
                    //   sink the loads next to their use, which serialises the batch -- kept as a control)
    LD_NA_256 = 5,  // L1::no_allocate + L2::256B prefetch size (SASS LTC256B): L2 fetches 256-B granules
    LD_HINTS = 6
};
// Store hints (tune.st_hint)
enum : int {
    ST_PLAIN = 0,   // st.global
    ST_NA    = 1,   // st.global.L1::no_allocate
    ST_CS    = 2,   // st.global.cs
    ST_NA_EF = 3,   // L1::no_allocate + L2 evict_first
    ST_HINTS = 4
};

struct f32x4 { float x, y, z, w; };
struct f32x8 { float v[8]; };

__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// ---------------------------------------------------------------- 128-bit global
template <int HINT>
__device__ __forceinline__ f32x4 ldg128(const float* p, uint64_t pol)
{
    f32x4 r;
    if constexpr (HINT == LD_PLAIN)
        asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    else if constexpr (HINT == LD_NA)
        asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    else if constexpr (HINT == LD_NC_NA)
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    else if constexpr (HINT == LD_CS)
        asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    else if constexpr (HINT == LD_NA_256)
        asm volatile("ld.global.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    else
        asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol) : "memory");
    return r;
}

template <int HINT>
__device__ __forceinline__ void stg128(float* p, const f32x4& v, uint64_t pol)
{
    if constexpr (HINT == ST_PLAIN)
        asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};"
                     :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else if constexpr (HINT == ST_NA)
        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                     :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else if constexpr (HINT == ST_CS)
        asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"
                     :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else
        asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                     :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// ---------------------------------------------------------------- 256-bit global
// ld/st.global.v8.f32 are new with PTX 8.8 / sm_100 (SASS LDG.E.256 / STG.E.256):
// one instruction moves a full 32-byte sector per thread, 1 KiB per warp.
template <int HINT>
__device__ __forceinline__ f32x8 ldg256(const float* p, uint64_t pol)
{
    f32x8 r;
    (void)pol;
    if constexpr (HINT == LD_PLAIN)
        asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]),
                       "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p) : "memory");
    else if constexpr (HINT == LD_NA)
        asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]),
                       "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p) : "memory");
    else if constexpr (HINT == LD_NC_NA)
        asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]),
                       "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p) : "memory");
    else if constexpr (HINT == LD_CS)
        asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]),
                       "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p) : "memory");
    else if constexpr (HINT == LD_NA_256)
        asm volatile("ld.global.L1::no_allocate.L2::256B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]),
                       "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p) : "memory");
    else
        asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]),
                       "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p) : "memory");
    return r;
}

template <int HINT>
__device__ __forceinline__ void stg256(float* p, const f32x8& r, uint64_t pol)
{
    (void)pol;
    if constexpr (HINT == ST_PLAIN)
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
                        "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
    else if constexpr (HINT == ST_NA)
        asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
                        "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
    else if constexpr (HINT == ST_CS)
        asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
                        "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
    else
        asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
                        "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
}

// ---------------------------------------------------------------- shared memory
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ f32x4 lds128(uint32_t saddr)
{
    f32x4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr) : "memory");
    return r;
}

__device__ __forceinline__ void sts128(uint32_t saddr, const f32x4& v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(bar), "r"(bytes) : "memory");
}

// Blocks (hardware-suspended try_wait, re-armed in a loop) until the phase with the
// given parity has completed.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}

// ---------------------------------------------------------------- TMA 1-D bulk copies
// global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP.S.G).
// dst/src 16-B aligned, bytes a multiple of 16.
template <bool L2_HINT>
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar, uint64_t pol)
{
    if constexpr (L2_HINT)
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
            "[%0], [%1], %2, [%3], %4;"
            :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
    else
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// shared -> global, tracked by the issuing thread's bulk async-group (SASS: UBLKCP.G.S).
template <bool L2_HINT>
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes, uint64_t pol)
{
    if constexpr (L2_HINT)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                     :: "l"(dst), "r"(src_smem), "r"(bytes), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}

// Bulk prefetch of [src, src + bytes) into L2 (no destination, nothing to wait for).  src 16-B
// aligned, bytes a multiple of 16.  L2 is the point of coherence, so prefetching data that an
// earlier kernel is still writing is harmless.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_commit()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// Wait until at most N of this thread's bulk groups still have their *source reads*
// outstanding (the shared-memory stage may then be overwritten).
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}

// Wait until at most N bulk groups are still incomplete (writes performed).
template <int N>
__device__ __forceinline__ void bulk_wait_all()
{
    asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory");
}

// Generic-proxy shared-memory writes -> visible to the async proxy (TMA) reads.
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Programmatic dependent launch (sm_90+).  `pdl_launch_dependents` lets the NEXT kernel in
// the stream start scheduling its CTAs as soon as every CTA of this grid has issued it
// (or exited); `pdl_wait` blocks until the PREVIOUS grid has completed and its memory
// operations are visible.  Together they hide launch latency and CTA ramp-up behind the
// previous launch's tail without weakening stream order.  Both are no-ops for a launch
// without the programmatic-stream-serialization attribute.
__device__ __forceinline__ void pdl_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ void pdl_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---------------------------------------------------------------- cluster launch control
// Blackwell's hardware work-stealing for persistent kernels (PTX 8.6, sm_100): a running
// CTA asks to cancel the launch of a not-yet-started CTA of the same grid and, on success,
// receives that CTA's index and does its work.  The 16-byte response lands in shared
// memory through the async proxy and completes 16 bytes on an mbarrier.
__device__ __forceinline__ void clc_try_cancel(uint32_t response_smem, uint32_t bar)
{
    asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];"
                 :: "r"(response_smem), "r"(bar) : "memory");
}

// Decodes a response: returns true and the cancelled CTA's blockIdx.x if the request
// succeeded; false if there was nothing left to cancel (the index is then undefined).
__device__ __forceinline__ bool clc_query(uint32_t response_smem, uint32_t& ctaid_x)
{
    uint64_t lo, hi;
    uint32_t ok, x;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(response_smem) : "memory");
    asm volatile(
        "{\n"
        ".reg .b128 resp;\n"
        ".reg .pred p;\n"
        "mov.b128 resp, {%2, %3};\n"
        "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p, resp;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "mov.b32 %1, 0;\n"
        "@p clusterlaunchcontrol.query_cancel.get_first_ctaid::x.b32.b128 %1, resp;\n"
        "}\n"
        : "=r"(ok), "=r"(x) : "l"(lo), "l"(hi));
    ctaid_x = x;
    return ok != 0;
}

// Named barrier over a subset of the CTA's warps (id 1..15; 0 is __syncthreads).
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

}  // namespace b200va
