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
// Programmatic dependent launch: every kernel of this library is launched with the stream-serialization attribute and
// starts with grid_dependency_sync(), so its CTAs become resident (and run their prologue: barrier init, TMEM allocation,
// tensor-map prefetch) while the previous kernel of the stream drains, and go on the moment that kernel's memory is
// visible.  In a captured graph the edge becomes a programmatic one.  CVAE_PDL=0 launches without the attribute (the
// device-side instructions are then no-ops).  RULE: nothing a predecessor writes may be read, and nothing it reads may
// be written, before grid_dependency_sync().
bool pdl_enabled();  // api.cu
__device__ __forceinline__ void grid_dependency_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);
}

int sm_count();     // cached multiprocessor count of the current device (api.cu)
int* fault_flag();  // device address of the pipeline-fault flag (api.cu)

}  // namespace cvae
