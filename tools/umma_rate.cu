// tcgen05.mma issue-rate microbenchmark: how many cycles does one 128 x N x 16 bf16 MMA take for a given
// shared-memory layout (swizzle mode, strides, start alignment) and number of independent accumulators?
// Operands are whatever bytes sit in shared memory (zeros); only timing is observed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu && tools/umma_rate
#include "../critic-vae_b200/csrc/umma.cuh"
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

using namespace cvae;

__device__ int g_rate_fault = 0;

struct RateArgs {
    uint32_t n, accs, reps;          // MMA N, independent accumulators, MMAs issued
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    uint32_t a_swz, b_swz;           // descriptor layout_type (0 none, 2 = 128B, 4 = 64B, 6 = 32B)
    uint32_t a_shift;                // bytes added to the A start address
    uint32_t a_step, b_step;         // start-address advance per MMA (wraps every 8)
    uint32_t a_major, b_major;
    unsigned long long* out;         // [grid] cycles
};

__device__ __forceinline__ uint64_t desc_full(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t swz) {
    uint64_t d = smem_desc(saddr, lbo, sbo);
    d |= (uint64_t)(swz & 7u) << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) rate_kernel(RateArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid * 16; i < 160 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_bf16(p.n, p.a_major, p.b_major);
        const uint32_t sa = smem_u32(smem) + p.a_shift, sb = smem_u32(smem) + 96 * 1024;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (uint32_t r = 0; r < p.reps; r += p.accs) {
                const uint64_t da = desc_full(sa + (r & 7u) * p.a_step, p.a_lbo, p.a_sbo, p.a_swz);
                const uint64_t db = desc_full(sb + (r & 7u) * p.b_step, p.b_lbo, p.b_sbo, p.b_swz);
                for (uint32_t t = 0; t < p.accs; ++t) umma_bf16(tmem_base + t * p.n, da + t * 128, db, idesc, r > 0);
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0, &g_rate_fault);
        t1 = clock64();
        if (t0 != 0) p.out[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, 512);
}

template <int ACCS>
__global__ void __launch_bounds__(128, 1) tight_kernel(RateArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid * 16; i < 160 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_bf16(p.n, p.a_major, p.b_major);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 96 * 1024;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            const uint64_t da = desc_full(sa, p.a_lbo, p.a_sbo, p.a_swz);
            const uint64_t db = desc_full(sb, p.b_lbo, p.b_sbo, p.b_swz);
            t0 = clock64();
            for (uint32_t r = 0; r < p.reps; r += 8 * ACCS) {
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int t = 0; t < ACCS; ++t) umma_bf16(tmem_base + t * p.n, da + (uint64_t)(t * 128 + u), db + (uint64_t)(u * 16), idesc, 1u);
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0, &g_rate_fault);
        t1 = clock64();
        if (t0 != 0) p.out[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, 512);
}

// Realistic pattern: A start cycles over 25 tap offsets of a halo-plane buffer, B walks a 64 KB ring;
// mode bit 0: warps 1-3 hammer shared memory with st.shared.v4; bit 1: warp 1 streams cp.async.bulk
// copies (global -> smem ring) as fast as they complete.
template <int N, int TM>
__global__ void __launch_bounds__(128, 1) real_kernel(RateArgs p, const uint8_t* gsrc, uint32_t mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, cbar[4];
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (uint32_t i = tid * 16; i < 200 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(&bar, 1);
        for (int i = 0; i < 4; ++i) mbar_init(&cbar[i], 1);
        mbar_fence_init();
        stop = 0;
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_bf16(N, kMajorK, kMajorK);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 96 * 1024;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            const uint32_t a_hi = (128u >> 4) | (1u << 14), b_hi = (256u >> 4) | (1u << 14);
            const uint32_t plane_stride = 4928, PW = 10;
            t0 = clock64();
            for (uint32_t r = 0; r < p.reps; r += 25 * TM) {
#pragma unroll 5
                for (int tap = 0; tap < 25; ++tap) {
                    const uint32_t off = (uint32_t)(22 + (tap / 5 - 2) * (int)PW + (tap % 5 - 2));
                    const uint32_t a_lo = ((sa >> 4) + off) | ((plane_stride >> 4) << 16);
                    const uint32_t b_lo = ((sb >> 4) + ((r / TM + tap) & 15u) * (N * 2)) | ((128u >> 4) << 16);
                    const uint64_t db = ((uint64_t)b_hi << 32) | b_lo;
#pragma unroll
                    for (int t = 0; t < TM; ++t)
                        umma_bf16(tmem_base + t * N, ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + t * 128), db, idesc, 1u);
                    if ((mode & 4u) && (tap % 5) == 4) umma_commit(&cbar[3]);   // a commit every 5*TM MMAs (never waited on)
                    if ((mode & 8u) && (tap % 5) == 4) tc_fence_after();
                    if ((mode & 16u) && (tap % 5) == 4) { mbar_try_wait(&cbar[2], 1); }
                }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0, &g_rate_fault);
        t1 = clock64();
        if (t0 != 0) p.out[blockIdx.x] = (unsigned long long)(t1 - t0);
        stop = 1;
    } else {
        if ((mode & 2u) && warp == 1) {
            uint32_t i = 0;
            while (!stop) {
                if (elect_one()) {
                    mbar_expect_tx(&cbar[i & 3], 8192);
                    bulk_g2s(smem + 160 * 1024 + (i & 3) * 8192, gsrc + ((i * 8192u) & 0xFFFFFu), 8192, &cbar[i & 3]);
                }
                __syncwarp();
                mbar_wait(&cbar[i & 3], (i >> 2) & 1, &g_rate_fault);
                ++i;
            }
        } else if (mode & 1u) {
            uint4* dst = reinterpret_cast<uint4*>(smem + 140 * 1024 + (warp - 1) * 4096);
            uint32_t k = 0;
            while (!stop) {
                dst[lane + 32 * (k & 7)] = make_uint4(k, k, k, k);
                ++k;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, 512);
}

template <int N, int TM>
static void run_real(int grid, uint32_t mode, const uint8_t* gsrc) {
    unsigned long long* d;
    cudaMalloc(&d, sizeof(unsigned long long) * grid);
    cudaMemset(d, 0, sizeof(unsigned long long) * grid);
    RateArgs p{};
    p.reps = 25 * TM * 40;
    p.out = d;
    const size_t smem = 210 * 1024;
    cudaFuncSetAttribute(real_kernel<N, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    real_kernel<N, TM><<<grid, 128, smem>>>(p, gsrc, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("real ERROR %s\n", cudaGetErrorString(e)); exit(1); }
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * grid);
    cudaMemcpy(h, d, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
    double sum = 0;
    for (int i = 0; i < grid; ++i) sum += (double)h[i];
    printf("REAL pattern N=%3d TM=%d grid=%3d mode=%u (4=commit, 8=fence, 16=trywait per 5 K steps): %7.1f cycles/MMA\n", N, TM, grid, mode,
           sum / grid / p.reps);
    free(h);
    cudaFree(d);
}

// Queue depth: how long does the issuing thread spend issuing K back-to-back MMAs (vs their execution)?
template <int K>
__global__ void __launch_bounds__(128, 1) queue_kernel(RateArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid * 16; i < 160 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_bf16(p.n, kMajorK, kMajorK);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 96 * 1024;
        if (elect_one()) {
            const uint64_t da = desc_full(sa, 4928, 128, 0), db = desc_full(sb, 128, 256, 0);
            const long long t0 = clock64();
#pragma unroll
            for (int u = 0; u < K; ++u) umma_bf16(tmem_base, da + (uint64_t)u, db + (uint64_t)(u * 16), idesc, 1u);
            const long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, 0, &g_rate_fault);
            const long long t2 = clock64();
            p.out[0] = (unsigned long long)(t1 - t0);
            p.out[1] = (unsigned long long)(t2 - t0);
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, 512);
}

template <int K>
static void run_queue(uint32_t n) {
    unsigned long long* d;
    cudaMalloc(&d, 16);
    RateArgs p{};
    p.n = n; p.out = d;
    cudaFuncSetAttribute(queue_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    queue_kernel<K><<<1, 128, 200 * 1024>>>(p);
    cudaDeviceSynchronize();
    unsigned long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("QUEUE N=%3u K=%2d MMAs: issue returns after %5llu cycles, all complete after %5llu cycles\n", n, K, h[0], h[1]);
    cudaFree(d);
}

// Cost of feeding descriptors from vector registers: per group of GRP MMAs, NB bases come from shared memory
// (LDS -> R2UR), the rest of the group uses immediates.
template <int GRP, int NB>
__global__ void __launch_bounds__(128, 1) r2ur_kernel(RateArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ uint32_t table[64];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid * 16; i < 160 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    if (tid < 64) table[tid] = tid & 7;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_bf16(p.n, kMajorK, kMajorK);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 96 * 1024;
        if (elect_one()) {
            const uint64_t da = desc_full(sa, 4928, 128, 0), db = desc_full(sb, 128, 256, 0);
            const volatile uint32_t* tb = table;
            const long long t0 = clock64();
            for (uint32_t r = 0; r < p.reps; r += GRP) {
                uint64_t a0 = da, b0 = db;
                uint32_t acc = tmem_base;
                if (NB >= 1) a0 += tb[(r / GRP) & 63];
                if (NB >= 2) b0 += tb[(r / GRP + 1) & 63] * 16;
                if (NB >= 3) acc += tb[(r / GRP + 2) & 63] & 1;
#pragma unroll
                for (int u = 0; u < GRP; ++u) umma_bf16(acc, a0 + (uint64_t)u, b0 + (uint64_t)(u * 16), idesc, 1u);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0, &g_rate_fault);
            p.out[0] = (unsigned long long)(clock64() - t0);
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, 512);
}

template <int GRP, int NB>
static void run_r2ur(uint32_t n) {
    unsigned long long* d;
    cudaMalloc(&d, 16);
    RateArgs p{};
    p.n = n; p.out = d; p.reps = GRP * 128;
    cudaFuncSetAttribute(r2ur_kernel<GRP, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    r2ur_kernel<GRP, NB><<<1, 128, 200 * 1024>>>(p);
    cudaDeviceSynchronize();
    unsigned long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("R2UR N=%3u group=%2d vector-fed bases=%d : %6.1f cycles/MMA\n", n, GRP, NB, (double)h[0] / p.reps);
    cudaFree(d);
}

static int g_tight = 0;

static void run(const char* name, RateArgs p, int grid) {
    unsigned long long* d;
    cudaMalloc(&d, sizeof(unsigned long long) * grid);
    cudaMemset(d, 0, sizeof(unsigned long long) * grid);
    p.out = d;
    const size_t smem = 200 * 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (g_tight) {
        cudaFuncSetAttribute(tight_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(tight_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(tight_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (p.accs == 1) tight_kernel<1><<<grid, 128, smem>>>(p);
        else if (p.accs == 2) tight_kernel<2><<<grid, 128, smem>>>(p);
        else tight_kernel<4><<<grid, 128, smem>>>(p);
    } else
    rate_kernel<<<grid, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * grid);
    cudaMemcpy(h, d, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
    double sum = 0;
    for (int i = 0; i < grid; ++i) sum += (double)h[i];
    printf("%-44s N=%3u accs=%u grid=%3d : %7.1f cycles/MMA (ideal %u)\n", name, p.n, p.accs, grid, sum / grid / p.reps, p.n / 2);
    free(h);
    cudaFree(d);
}


// Tight loop with both start addresses advancing 256 B per MMA (the weight-gradient K loop) and the majors /
// strides taken from the arguments; B sits at +96 KB.  `sbo` values are plane strides.
__global__ void __launch_bounds__(128, 1) mn_kernel(RateArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid * 16; i < 200 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_bf16(p.n, p.a_major, p.b_major);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 96 * 1024;
        if (elect_one()) {
            const uint64_t da = desc_full(sa, p.a_lbo, p.a_sbo, 0), db = desc_full(sb, p.b_lbo, p.b_sbo, 0);
            const long long t0 = clock64();
            for (uint32_t r = 0; r < p.reps; r += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) umma_bf16(tmem_base + (u & 3) * p.n * (p.accs > 1), da + (uint64_t)(u * 16), db + (uint64_t)(u * 16), idesc, 1u);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0, &g_rate_fault);
            p.out[blockIdx.x] = (unsigned long long)(clock64() - t0);
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, 512);
}
static void run_mn(const char* name, RateArgs p, int grid) {
    unsigned long long* d;
    cudaMalloc(&d, grid * 8);
    cudaMemset(d, 0, grid * 8);
    p.out = d;
    cudaFuncSetAttribute(mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    mn_kernel<<<grid, 128, 200 * 1024>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
    unsigned long long h[148];
    cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
    double sum = 0;
    for (int i = 0; i < grid; ++i) sum += (double)h[i];
    printf("%-48s N=%3u accs=%u grid=%3d : %7.1f cycles/MMA (N/2 = %u)\n", name, p.n, p.accs, grid, sum / grid / p.reps, p.n / 2);
    cudaFree(d);
}

// The weights-as-A pattern of conv_wa.cu: A = no-swizzle K-major 128 x 64 units (16 KB: four K = 16 steps 4 KB apart) in a
// ring, B = one SWIZZLE_128B K-major pixel tile read at 25 tap-shifted (not 8-row aligned) starts, 32 B per K step.
// mode bit 0: a second warp keeps streaming 16 KB bulk copies global -> the A ring (what the weight producer does).
__global__ void __launch_bounds__(128, 1) wa_kernel(RateArgs p, const uint8_t* gsrc, uint32_t mode, uint32_t pw) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, cbar[4], done_bar, sink_bar[8];
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid * 16; i < 200 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init(&done_bar, 1);
        for (int i = 0; i < 8; ++i) mbar_init(&sink_bar[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&cbar[i], 1);
        mbar_fence_init();
        mbar_arrive(&done_bar);
        stop = 0;
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(p.n, kMajorK, kMajorK);
            const uint32_t a16 = (smem_u32(smem) & 0x3FFFFu) >> 4, b16 = ((smem_u32(smem) + 128 * 1024) & 0x3FFFFu) >> 4;
            const uint32_t a_hi = (256u >> 4) | (1u << 14), a_lbo = (128u >> 4) << 16;
            const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29), b_lbo = 1u << 16;
            const long long t0 = clock64();
            uint32_t unit = 0;
            for (uint32_t r = 0; r < p.reps; r += 4, ++unit) {
                const uint32_t tap = unit % 25u, ty = tap / 5u, tx = tap - ty * 5u;
                const uint32_t a_lo = (a16 + (unit & 7u) * 1024u) | a_lbo;
                const uint32_t b_lo = (b16 + (8u + ty * pw + tx) * 8u) | b_lbo;
                if ((unit & 1u) == 0) {      // what conv_wa's MMA thread does once per weight stage (two units)
                    if (mode & 8u) mbar_wait(&done_bar, 0, &g_rate_fault);       // a barrier that completed long ago
                    if (mode & 4u) tc_fence_after();
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    umma_bf16(tmem_base, ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)(j * 256)),
                              ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)(j * 2)), idesc, 1u);
                if ((unit & 1u) && (mode & 2u)) umma_commit(&sink_bar[(unit >> 1) & 7u]);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0, &g_rate_fault);
            p.out[blockIdx.x] = (unsigned long long)(clock64() - t0);
            stop = 1;
        }
        __syncwarp();
    } else if (warp == 1 && (mode & 1u)) {
        if (elect_one()) {
            uint32_t i = 0;
            while (!stop) {
                const uint32_t s = i & 3u;
                if (i >= 4) mbar_wait(&cbar[s], ((i >> 2) - 1) & 1u, &g_rate_fault);
                mbar_expect_tx(&cbar[s], 16384);
                bulk_g2s(smem + (size_t)(i & 7u) * 16384, gsrc + (size_t)((i * 16384u) & ((1u << 20) - 1)), 16384, &cbar[s]);
                ++i;
            }
            for (uint32_t k = (i > 4 ? i - 4 : 0); k < i; ++k) mbar_wait(&cbar[k & 3u], (k >> 2) & 1u, &g_rate_fault);
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, 512);
}
static void run_wa(uint32_t n, int grid, uint32_t mode, uint32_t pw, const uint8_t* gsrc) {
    unsigned long long* d;
    cudaMalloc(&d, grid * 8);
    cudaMemset(d, 0, grid * 8);
    RateArgs p{};
    p.n = n; p.reps = 1600; p.out = d;
    cudaFuncSetAttribute(wa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    wa_kernel<<<grid, 128, 200 * 1024>>>(p, gsrc, mode, pw);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("wa ERROR %s\n", cudaGetErrorString(e)); exit(1); }
    unsigned long long h[148];
    cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
    double sum = 0;
    for (int i = 0; i < grid; ++i) sum += (double)h[i];
    printf("WA pattern N=%3u grid=%3d weights %-30s per stage:%s%s%s : %7.1f cycles/MMA (N/2 = %u)\n", n, grid,
           (mode & 1) ? "streaming (16 KB bulk copies)" : "resident", (mode & 8) ? " try_wait" : "", (mode & 4) ? " fence" : "",
           (mode & 2) ? " commit" : "", sum / grid / p.reps, n / 2);
    cudaFree(d);
}

int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "wa")) {
        uint8_t* gsrc;
        cudaMalloc(&gsrc, 2 << 20);
        cudaMemset(gsrc, 0, 2 << 20);
        for (int grid : {1, 148})
            for (uint32_t mode : {0u, 1u})
                for (uint32_t n : {144u, 176u, 208u, 256u}) run_wa(n, grid, mode, 18, gsrc);
        for (uint32_t mode : {2u, 4u, 8u, 14u, 15u})
            for (uint32_t n : {64u, 144u, 176u}) run_wa(n, 148, mode, 18, gsrc);
        return 0;
    }
    const uint32_t reps = 960;
    if (argc > 1 && !strcmp(argv[1], "real")) {
        // mode bit 0: three warps hammer shared memory with st.shared.v4; bit 1: one warp streams 8 KB bulk copies
        uint8_t* gsrc;
        cudaMalloc(&gsrc, 2 << 20);
        cudaMemset(gsrc, 0, 2 << 20);
        for (int grid : {1, 148})
            for (uint32_t mode : {0u, 1u, 2u, 3u}) {
                run_real<128, 1>(grid, mode, gsrc);
                run_real<128, 2>(grid, mode, gsrc);
                run_real<64, 2>(grid, mode, gsrc);
                run_real<64, 4>(grid, mode, gsrc);
                run_real<32, 4>(grid, mode, gsrc);
            }
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "mn")) {
        // weight-gradient layouts: operands MN-major (16 B = 8 M/N elements, pixel slots = K), LBO 128,
        // SBO = plane stride (plane = 256 pixel slots here); the start address advances 256 B (16 pixels) per MMA
        for (int grid : {1, 148})
            for (uint32_t n : {16u, 32u, 64u, 128u, 256u})
                for (uint32_t accs : {1u, 4u}) {
                    if (accs * n > 512) continue;
                    const uint32_t ps = 257 * 16;
                    RateArgs p{n, accs, reps, 128, ps, 128, ps, 0, 0, 0, 256, 256, kMajorMN, kMajorMN, nullptr};
                    if (n * ps <= 8 * 100 * 1024) run_mn("MN-major A and B", p, grid);
                    RateArgs q{n, accs, reps, 128, 16, 128, ps, 0, 0, 0, 256, 256, kMajorMN, kMajorMN, nullptr};
                    run_mn("MN-major, shift-trick A (SBO 16)", q, grid);
                    RateArgs r{n, accs, reps, ps, 128, 128, ps, 0, 0, 0, 256, 256, kMajorK, kMajorMN, nullptr};
                    run_mn("K-major A (SBO 128, LBO plane), MN-major B", r, grid);
                    RateArgs t{n, accs, reps, ps, 128, 128, 256, 0, 0, 0, 256, 256, kMajorK, kMajorK, nullptr};
                    run_mn("K-major A, K-major packed B", t, grid);
                }
        return 0;
    }
    for (uint32_t n : {64u, 128u}) {
        run_r2ur<8, 0>(n); run_r2ur<8, 1>(n); run_r2ur<8, 2>(n); run_r2ur<8, 3>(n);
        run_r2ur<4, 0>(n); run_r2ur<4, 1>(n); run_r2ur<4, 3>(n);
        run_r2ur<16, 3>(n);
    }
    return 0;
    for (uint32_t n : {64u, 128u, 256u}) {
        run_queue<1>(n); run_queue<2>(n); run_queue<4>(n); run_queue<8>(n); run_queue<16>(n); run_queue<32>(n);
    }
    return 0;
    uint8_t* gsrc;
    cudaMalloc(&gsrc, 2 << 20);
    cudaMemset(gsrc, 0, 2 << 20);
    for (int grid : {1, 148})
        for (uint32_t mode : {0u, 4u, 8u, 16u, 28u}) {
            run_real<128, 1>(grid, mode, gsrc);
            run_real<128, 2>(grid, mode, gsrc);
            run_real<64, 2>(grid, mode, gsrc);
            run_real<64, 4>(grid, mode, gsrc);
        }
    return 0;
    g_tight = 1;
    for (uint32_t n : {16u, 32u, 64u, 128u, 256u})
        for (uint32_t accs : {1u, 2u, 4u}) {
            if (accs * n > 512) continue;
            RateArgs p{n, accs, reps, 4928, 128, 128, 256, 0, 0, 0, 16, n * 32, kMajorK, kMajorK, nullptr};
            run("TIGHT unrolled, noswz halo", p, 1);
            RateArgs q{n, accs, reps, 16, 1024, 16, 1024, 2, 2, 0, 32, 32, kMajorK, kMajorK, nullptr};
            run("TIGHT unrolled, swizzle128", q, 1);
        }
    g_tight = 0;
    for (int grid : {1}) {
        for (uint32_t n : {64u, 128u, 256u}) {
            for (uint32_t accs : {1u, 2u, 4u}) {
                if (accs * n > 512) continue;
                // halo-plane layout: K-major, no swizzle, A: LBO = plane stride (4928), SBO 128; B: LBO 128, SBO 256
                RateArgs p{n, accs, reps, 4928, 128, 128, 256, 0, 0, 0, 16, n * 32, kMajorK, kMajorK, nullptr};
                run("noswz halo (A shifts 16 B per MMA)", p, grid);
            }
        }
        {
            RateArgs p{128, 1, reps, 4928, 128, 128, 256, 0, 0, 0, 0, 0, kMajorK, kMajorK, nullptr};
            run("noswz, same operands every MMA", p, grid);
            RateArgs q{128, 1, reps, 128, 256, 128, 256, 0, 0, 0, 0, 0, kMajorK, kMajorK, nullptr};
            run("noswz canonical A (LBO 128, SBO 256)", q, grid);
            RateArgs r{128, 1, reps, 4928, 128, 128, 256, 0, 0, 0, 128, 4096, kMajorK, kMajorK, nullptr};
            run("noswz halo, A shifts 128 B per MMA", r, grid);
            RateArgs s{128, 1, reps, 4096, 128, 128, 256, 0, 0, 0, 16, 4096, kMajorK, kMajorK, nullptr};
            run("noswz halo, plane stride 4096", s, grid);
            RateArgs s2{128, 1, reps, 4928 + 64, 128, 128, 256, 0, 0, 0, 16, 4096, kMajorK, kMajorK, nullptr};
            run("noswz halo, plane stride 4992 (=0 mod 128)", s2, grid);
        }
        for (uint32_t n : {64u, 128u, 256u}) {
            // canonical SWIZZLE_128B K-major: rows of 128 B, 8-row atoms of 1024 B; K advance 32 B inside the atom
            RateArgs p{n, 1, reps, 16, 1024, 16, 1024, 2, 2, 0, 32, 32, kMajorK, kMajorK, nullptr};
            run("swizzle128 canonical", p, grid);
            RateArgs q{n, 2, reps, 16, 1024, 16, 1024, 2, 2, 0, 32, 32, kMajorK, kMajorK, nullptr};
            if (2 * n <= 512) run("swizzle128 canonical", q, grid);
        }
    }
    return 0;
}
