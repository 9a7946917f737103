// Blackwell / Hopper bulk-async ("TMA") plumbing shared by the image kernels: tensor-map encoding on the host (the driver
// entry point is fetched through the runtime, libcuda is not linked) and the mbarrier / cp.async.bulk PTX on the device.
// SASS evidence: cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS (profiles/r02_sass_grep.txt).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

// u8 tensor [n][h][w] with byte strides (1, row_stride, img_pitch); box (box_w, box_h, 1); out-of-bounds bytes read as 0.
// row_stride and img_pitch must be multiples of 16, base 16-byte aligned, box_w a multiple of 16.  Returns 0 on success.
int svs_i_tmap_u8_3d(CUtensorMap *m, const void *base, int w, int h, int n, size_t row_stride, size_t img_pitch, int box_w, int box_h);

#ifdef __CUDACC__
namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the barrier initialisation visible to the async proxy before the first bulk copy names it
__device__ __forceinline__ void fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order earlier generic-proxy accesses of shared memory before later async-proxy (bulk copy) writes to it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D;\n"
        "bra W;\n"
        "D:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one box of a 3-D tensor map -> shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void load_3d(void *smem, const CUtensorMap *map, uint64_t *bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
// `bytes` (multiple of 16) from 16-byte aligned global memory -> 16-byte aligned shared memory
__device__ __forceinline__ void load_1d(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
}  // namespace tma
#endif
