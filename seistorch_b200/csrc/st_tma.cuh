// TMA (cp.async.bulk.tensor) + mbarrier helpers for sm_100a and the host-side tensor-map
// encoder.  The stencil kernels stage field tiles (with their halos) in shared memory through
// a multi-stage ring of bulk tensor loads, so the bytes in flight per SM are set by the ring
// depth and not by resident warps x registers; out-of-range box elements are zero-filled by
// the hardware, which is exactly the zero-padding of the reference's conv2d (acoustic.py:60-71).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "st_common.cuh"

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void st_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async (TMA) proxy
__device__ __forceinline__ void st_mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void st_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = st_smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// one box of a rank-3 tensor (cols, rows, planes) -> shared memory; completion on `bar`
__device__ __forceinline__ void st_tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            st_smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(st_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// (programmatic dependent launch helpers st_pdl_*: st_common.cuh)
__device__ __forceinline__ void st_tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
#endif

// Host: describe `planes` row-major fp32 planes [rows][ld] (plane stride `plane_elems`) and a
// (box_cols x box_rows x 1) box.  Returns 0 on success.
int st_tma_encode_planes(CUtensorMap* out, const float* base, int cols, int rows, long long planes, int ld,
                         long long plane_elems, int box_cols, int box_rows);
