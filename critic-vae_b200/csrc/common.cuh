// Error plumbing and small device helpers shared by every translation unit of libcvae.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cvae.h"

namespace cvae {

void set_error(const char* fmt, ...);  // defined in api.cu; thread-local message buffer
void count_launch();                   // api.cu: one tick per kernel launched by this library

#define CVAE_REQUIRE(cond, code, ...)   \
    do {                                \
        if (!(cond)) {                  \
            cvae::set_error(__VA_ARGS__); \
            return (code);              \
        }                               \
    } while (0)

#define CVAE_CUDA(expr)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            cvae::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                            __LINE__);                                                   \
            return CVAE_ECUDA;                                                           \
        }                                                                                \
    } while (0)

#define CVAE_LAUNCH_CHECK()                                                              \
    do {                                                                                 \
        cvae::count_launch();                                                            \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            cvae::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                         \
            return CVAE_ECUDA;                                                           \
        }                                                                                \
    } while (0)

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ float bf16_round(float x) {
    return __bfloat162float(__float2bfloat16_rn(x));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

int opt_in_smem(const void* func, size_t bytes);  // api.cu: raise a kernel's dynamic shared-memory limit (process-wide, per device)
#define CVAE_OPT_IN_SMEM(kern, bytes)                                                    \
    do {                                                                                 \
        int rc__ = cvae::opt_in_smem(reinterpret_cast<const void*>(kern), (bytes));      \
        if (rc__ != CVAE_OK) return rc__;                                                \
    } while (0)
int sm_count();     // cached multiprocessor count of the current device (api.cu)
int* fault_flag();  // device address of the pipeline-fault flag (api.cu)

}  // namespace cvae
