// TMA throughput probe for the halo-plane layout: how fast does cp.async.bulk.tensor move boxes whose inner
// dimension is ONE 16-byte channel group (8 bf16 channels) of an NHWC tensor into shared-memory planes
// [virtual pixel][8 ch], with the padding columns / rows produced by out-of-bounds zero fill?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu && tools/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>

#include "../critic-vae_b200/csrc/umma.cuh"
using namespace cvae;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ int g_fault = 0;

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

struct Args {
    int planes, box_bytes, iters, images, nb, ih, pad, verify, inner;
    unsigned long long* cycles;
    uint4* dump;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap map, Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    __syncthreads();
    const int stride = (a.box_bytes + 1023) & ~1023;
    if (warp == 0) {
        long long t0 = clock64();
        for (int it = 0; it < a.iters; ++it) {
            const int buf = it & 1;
            // previous use of this buffer: wait (so two chunks are in flight)
            if (it >= 2) mbar_wait(&bar[buf], ((it - 2) >> 1) & 1, &g_fault);
            const int n0 = ((blockIdx.x * a.iters + it) * a.nb) % (a.images - a.nb + 1);
            if (lane == 0) mbar_expect_tx(&bar[buf], (uint32_t)(a.planes * a.box_bytes));
            __syncwarp();
            for (int q = lane; q < a.planes; q += 32)
                tma_load_4d(smem + (size_t)buf * a.planes * stride + (size_t)q * stride, &map, q * a.inner, 0, -a.pad, n0, &bar[buf]);
        }
        for (int it = (a.iters >= 2 ? a.iters - 2 : 0); it < a.iters; ++it) mbar_wait(&bar[it & 1], (it >> 1) & 1, &g_fault);
        long long t1 = clock64();
        if (lane == 0) a.cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    __syncthreads();
    if (a.verify && blockIdx.x == 0) {
        const int buf = (a.iters - 1) & 1;
        const uint4* s = reinterpret_cast<const uint4*>(smem + (size_t)buf * a.planes * stride);
        for (int i = tid; i < a.planes * stride / 16; i += blockDim.x) a.dump[i] = s[i];
    }
}

int main() {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    struct Cfg { int H, C, R, NB, planes; };   // R rows per box (IH = whole image), NB images per box
    const int inner = (getenv("TMA_INNER") ? atoi(getenv("TMA_INNER")) : 8);   // channels per box row: 8 (16 B, no swizzle) or 64 (128 B, SWIZZLE_128B)
    const Cfg cfgs[] = {{8, 128, 10, 5, 16}, {8, 128, 10, 5, 32}, {4, 256, 6, 16, 32}, {16, 64, 18, 2, 8}, {32, 32, 17, 1, 4}, {32, 64, 17, 1, 8},
                        {64, 32, 12, 1, 4}};
    const int B = 256, pad = 2;
    for (const Cfg& c : cfgs) {
        const int H = c.H, W = c.H, C = c.C, PW = W + pad;
        const size_t elems = (size_t)B * H * W * C;
        std::vector<__nv_bfloat16> h(elems);
        for (size_t i = 0; i < elems; ++i) h[i] = __float2bfloat16((float)((i * 2654435761u >> 20) % 251) - 125.f);
        __nv_bfloat16* d;
        cudaMalloc(&d, elems * 2);
        cudaMemcpy(d, h.data(), elems * 2, cudaMemcpyHostToDevice);
        CUtensorMap map;
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)inner, (cuuint32_t)PW, (cuuint32_t)c.R, (cuuint32_t)c.NB};
        if (inner > C) { printf("skip C=%d\n", C); cudaFree(d); continue; }
        const int planes = inner == 8 ? c.planes : (C / inner);
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        Args a{};
        a.planes = planes; a.box_bytes = 2 * inner * PW * c.R * c.NB; a.inner = inner; a.iters = 40; a.images = B; a.nb = c.NB; a.ih = c.R; a.pad = pad; a.verify = 1;
        const int stride = (a.box_bytes + 1023) & ~1023;
        const size_t smem = (size_t)2 * a.planes * stride;
        if (smem > 220 * 1024) { printf("cfg too big\n"); continue; }
        cudaMalloc(&a.cycles, 148 * 8);
        cudaMalloc(&a.dump, smem);
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int grid : {1, 148}) {
            probe_kernel<<<grid, 128, smem>>>(map, a);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<unsigned long long> cyc(grid);
            cudaMemcpy(cyc.data(), a.cycles, grid * 8, cudaMemcpyDeviceToHost);
            double mean = 0;
            for (auto v : cyc) mean += (double)v / grid;
            const double bytes = (double)a.iters * a.planes * a.box_bytes;
            printf("H=%2d C=%3d box {8,%d,%d,%d} = %6d B x %2d planes, grid %3d: %8.0f cycles, %6.1f B/clk/SM (%5.2f TB/s chip at 1.9 GHz)\n", H, C, PW, c.R,
                   c.NB, a.box_bytes, a.planes, grid, mean, bytes / mean, bytes / mean * grid * 1.9e9 / 1e12);
        }
        // verify the last box of CTA 0 (grid 148 run): plane q slot (img, r, col)
        std::vector<uint16_t> dump(smem / 2);
        cudaMemcpy(dump.data(), a.dump, (size_t)a.planes * stride, cudaMemcpyDeviceToHost);
        const int it = a.iters - 1, n0 = ((0 * a.iters + it) * a.nb) % (B - a.nb + 1);
        long bad = 0;
        if (inner == 64) {   // swizzled check: row = slot, 16-byte chunk ch of the 128-byte row sits at chunk (ch ^ (row & 7)) (absolute: planes are 1024 B aligned)
            for (int q = 0; q < a.planes; ++q)
                for (int slot = 0; slot < c.NB * c.R * PW; ++slot)
                    for (int ch = 0; ch < 64; ++ch) {
                        const int img = slot / (c.R * PW), rr = (slot / PW) % c.R, col = slot % PW, hh = rr - pad, n = n0 + img;
                        uint16_t want = 0;
                        if (hh >= 0 && hh < H && col < W) {
                            __nv_bfloat16 v = h[(((size_t)n * H + hh) * W + col) * C + q * 64 + ch];
                            want = *reinterpret_cast<uint16_t*>(&v);
                        }
                        const size_t byte = (size_t)q * stride + (size_t)slot * 128 + (size_t)(((ch >> 3) ^ (slot & 7)) << 4) + (ch & 7) * 2;
                        if (dump[byte / 2] != want) ++bad;
                    }
        } else
        for (int q = 0; q < a.planes; ++q)
            for (int img = 0; img < c.NB; ++img)
                for (int rr = 0; rr < c.R; ++rr)
                    for (int col = 0; col < PW; ++col)
                        for (int e = 0; e < 8; ++e) {
                            const int hh = rr - pad, n = n0 + img;
                            uint16_t want = 0;
                            if (hh >= 0 && hh < H && col < W) {
                                __nv_bfloat16 v = h[(((size_t)n * H + hh) * W + col) * C + q * 8 + e];
                                want = *reinterpret_cast<uint16_t*>(&v);
                            }
                            uint16_t got;
                            if (inner == 8) got = dump[((size_t)q * stride + ((size_t)(img * c.R + rr) * PW + col) * 16) / 2 + e];
                            else got = 0;
                            if (got != want) ++bad;
                        }
        printf("   verify: %ld mismatches\n", bad);
        cudaFree(d); cudaFree(a.cycles); cudaFree(a.dump);
    }
    return 0;
}
