// surf_ptx.cuh -- the inline-PTX building blocks shared by the kernels that stage data with the sm_100
// asynchronous copy engines: mbarrier (init / expect_tx / bounded wait), TMA tensor-tile loads
// (cp.async.bulk.tensor through a CUtensorMap), TMA bulk 1-D copies and the generic<->async proxy fence.
// Every wait is bounded: a copy that never lands is reported (false) instead of hanging the GPU.
#pragma once

#include <cstdint>

#include <cuda.h>

namespace surfb200 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// makes the initialised barriers visible to the async proxy (call once after the mbar_init's, before the barrier)
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// one arrival (release: this thread's earlier writes are visible to whoever sees the phase complete)
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Spins on the barrier's phase `parity`; false if it has not completed after ~2^24 polls.
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spin = 0; !done && spin < (1u << 24); spin++)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    return done != 0;
}

// generic-proxy accesses of shared memory issued so far are ordered before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TMA tensor-tile load: box of the 3-D tensor `map` at (x, y, z) -> shared memory; x must be 16-byte aligned
// and dst 128-byte aligned.  Completion is counted in bytes on `bar` (arm it with mbar_expect_tx first).
__device__ __forceinline__ void tma_load_3d(const CUtensorMap *map, uint32_t dst, int x, int y, int z, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}

// TMA bulk copy (1-D, no tensor map): global -> shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

}  // namespace surfb200
