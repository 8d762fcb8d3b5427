// TMA (cp.async.bulk.tensor) + mbarrier plumbing for the tile-staged kernels.
//
// Host side: tensor maps are encoded with cuTensorMapEncodeTiled, resolved at run time through
// cudaGetDriverEntryPoint so the library does not link against libcuda.  Device side: raw PTX.
// Out-of-bounds parts of a box are zero-filled by the TMA unit, which is precisely grid_sample's
// padding_mode='zeros' (reference utils.py:555) -- the staged kernels need no bounds predicates.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ofb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// (W, H, P) tensor of `elem`-byte elements, dense rows; box (bw, bh, bp).  Returns false if the tensor does not
// meet the TMA constraints (16 B aligned base, row pitch multiple of 16 B) or the driver refuses.
inline bool make_map_3d(CUtensorMap *m, CUtensorMapDataType dt, int elem, const void *base, uint64_t W, uint64_t H,
                        uint64_t P, uint32_t bw, uint32_t bh, uint32_t bp)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || ((W * elem) & 15u) || ((bw * elem) & 15u)) return false;
    if (bw > 256 || bh > 256 || bp > 256) return false;
    const cuuint64_t dims[3] = {W, H, P};
    const cuuint64_t strides[2] = {W * elem, W * H * elem};
    if (strides[1] & 15u) return false;
    const cuuint32_t box[3] = {bw, bh, bp};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// global -> shared box copy; completion is signalled on `bar` (complete_tx::bytes)
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// make generic-proxy smem accesses (the previous chunk's reads) ordered before the next async-proxy write
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace ofb
