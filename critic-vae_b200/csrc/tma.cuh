// Tensor-map (TMA) helpers shared by the conv kernels: the driver's encoder is fetched through
// cudaGetDriverEntryPoint (no libcuda link), maps travel as __grid_constant__ kernel parameters.
#pragma once
#include <cuda.h>   // CUtensorMap types only

#include "common.cuh"
#include "umma.cuh"

namespace cvae {

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encoder();   // conv_wgrad.cu; nullptr when the driver does not export it

// bf16 tensor viewed as {channels, w, h, n} with element strides (in elements) for w, h, n; box {bc, bw, bh, bn};
// out-of-bounds elements read as zero (that is where the convolution padding comes from).
static inline bool encode_map_4d(CUtensorMap* m, const void* base, int channels, int W, int H, int B, long sw, long sh, long sn, int bc, int bw,
                                 int bh, int bn, CUtensorMapSwizzle swz) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t gdim[4] = {(cuuint64_t)channels, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sn * 2};
    cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// plain 2-D bf16 matrix [rows][cols] (row pitch = cols), box {bcols, brows}, no swizzle: lands as brows dense rows
static inline bool encode_map_2d(CUtensorMap* m, const void* base, long cols, long rows, int bcols, int brows) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)bcols, (cuuint32_t)brows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace cvae
