// TMA / mbarrier plumbing shared by the kernels that move tiles with cp.async.bulk.tensor
// (slice_cols_tma in gx_fused.cu, detector_affine_brick_kernel in gx_detector_affine.cu).
#pragma once
#include <cuda.h>
#include "gx_common.cuh"

typedef CUresult (*gx_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                       const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);


// cuTensorMapEncodeTiled through the runtime's driver entry point query: the library does not link libcuda
static inline gx_encode_tiled_fn gx_tensor_map_encoder()
{
    static gx_encode_tiled_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess) {
            gx_set_error("cuTensorMapEncodeTiled is not available from this driver");
            cudaGetLastError();
            return nullptr;
        }
        encode = reinterpret_cast<gx_encode_tiled_fn>(fn);
    }
    return encode;
}

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t gx_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gx_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gx_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gx_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(gx_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(gx_smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(gx_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_barrier(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(gx_smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(gx_smem_u32(bar)) : "memory");
}
#endif
