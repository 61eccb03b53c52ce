// Bring-up probe for the tcgen05 building blocks the conv kernels rely on.  The host hands over raw
// shared-memory images for the A and B operands plus descriptor fields; the kernel copies the
// images into smem verbatim, issues `ksteps` UMMAs (M=128) and returns the fp32 accumulator.
// tools/umma_probe.py builds images for every layout hypothesis and checks them against numpy.
#include "../critic-vae_b200/csrc/umma.cuh"
#include <stdio.h>

__device__ int g_probe_fault = 0;

using namespace cvae;

struct ProbeArgs {
    const uint8_t* a_img;
    const uint8_t* b_img;
    uint32_t a_bytes, b_bytes;
    uint32_t a_off, b_off;        // byte offset of the descriptor start inside each region
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    uint32_t a_kstep, b_kstep;    // start-address advance per K=16 step
    uint32_t ksteps;
    uint32_t idesc;
    uint32_t n;
    uint32_t use_bulk;            // copy the B image with cp.async.bulk instead of st.shared
    float* out;                   // [128][n]
    uint32_t a_swz, b_swz;        // descriptor layout type (bits 61-63): 0 none, 2 = 128 B swizzle, 4 = 64 B, 6 = 32 B
    uint32_t a_boff, b_boff;      // descriptor base offset (bits 49-51)
};

__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_mma, bar_copy;
    __shared__ uint32_t tmem_slot;

    uint8_t* sa = smem;
    uint8_t* sb = smem + ((p.a_bytes + 1023u) & ~1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&bar_mma, 1);
        mbar_init(&bar_copy, 1);
        mbar_fence_init();
    }
    uint32_t ncols = 32;
    while (ncols < p.n) ncols <<= 1;
    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    for (uint32_t i = tid * 16; i < p.a_bytes; i += 128 * 16)
        *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(p.a_img + i);
    if (p.use_bulk) {
        if (tid == 0) {
            mbar_expect_tx(&bar_copy, p.b_bytes);
            bulk_g2s(sb, p.b_img, p.b_bytes, &bar_copy);
        }
    } else {
        for (uint32_t i = tid * 16; i < p.b_bytes; i += 128 * 16)
            *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(p.b_img + i);
    }
    fence_proxy_async();
    __syncthreads();

    if (tid == 0) {
        bool ok = true;
        if (p.use_bulk) ok = mbar_wait(&bar_copy, 0, &g_probe_fault);
        if (ok) {
            tc_fence_after();
            for (uint32_t k = 0; k < p.ksteps; ++k) {
                uint64_t da = smem_desc(smem_u32(sa) + p.a_off + k * p.a_kstep, p.a_lbo, p.a_sbo);
                uint64_t db = smem_desc(smem_u32(sb) + p.b_off + k * p.b_kstep, p.b_lbo, p.b_sbo);
                da |= ((uint64_t)(p.a_swz & 7u) << 61) | ((uint64_t)(p.a_boff & 7u) << 49);
                db |= ((uint64_t)(p.b_swz & 7u) << 61) | ((uint64_t)(p.b_boff & 7u) << 49);
                umma_bf16(tmem_base, da, db, p.idesc, k > 0);
            }
        }
        umma_commit(&bar_mma);
    }
    __syncwarp();
    mbar_wait(&bar_mma, 0, &g_probe_fault);
    tc_fence_after();

    for (uint32_t c = 0; c < p.n; c += 8) {
        uint32_t v[8];
        tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_wait_ld();
        float* o = p.out + (size_t)(warp * 32 + lane) * p.n + c;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, ncols);
}

extern "C" int probe_run_swz(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes,
                             uint32_t a_off, uint32_t b_off, uint32_t a_lbo, uint32_t a_sbo,
                             uint32_t b_lbo, uint32_t b_sbo, uint32_t a_kstep, uint32_t b_kstep,
                             uint32_t ksteps, uint32_t idesc, uint32_t n, uint32_t use_bulk, float* out,
                             uint32_t a_swz, uint32_t b_swz, uint32_t a_boff, uint32_t b_boff);

extern "C" int probe_run(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes,
                         uint32_t a_off, uint32_t b_off, uint32_t a_lbo, uint32_t a_sbo,
                         uint32_t b_lbo, uint32_t b_sbo, uint32_t a_kstep, uint32_t b_kstep,
                         uint32_t ksteps, uint32_t idesc, uint32_t n, uint32_t use_bulk, float* out) {
    return probe_run_swz(a_img, a_bytes, b_img, b_bytes, a_off, b_off, a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep, ksteps, idesc, n,
                         use_bulk, out, 0, 0, 0, 0);
}

extern "C" int probe_run_swz(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes,
                             uint32_t a_off, uint32_t b_off, uint32_t a_lbo, uint32_t a_sbo,
                             uint32_t b_lbo, uint32_t b_sbo, uint32_t a_kstep, uint32_t b_kstep,
                             uint32_t ksteps, uint32_t idesc, uint32_t n, uint32_t use_bulk, float* out,
                             uint32_t a_swz, uint32_t b_swz, uint32_t a_boff, uint32_t b_boff) {
    ProbeArgs p{(const uint8_t*)a_img, (const uint8_t*)b_img, a_bytes, b_bytes, a_off, b_off,
                a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep, ksteps, idesc, n, use_bulk, out, a_swz, b_swz, a_boff, b_boff};
    size_t smem = ((a_bytes + 1023u) & ~1023u) + ((b_bytes + 1023u) & ~1023u) + 1024;
    cudaError_t e = cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return -1;
    int zero = 0;
    cudaMemcpyToSymbol(g_probe_fault, &zero, sizeof(int));
    probe_kernel<<<1, 128, smem>>>(p);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        fprintf(stderr, "probe: %s\n", cudaGetErrorString(e));
        return -2;
    }
    int fault = 0;
    cudaMemcpyFromSymbol(&fault, g_probe_fault, sizeof(int));
    return fault ? -3 : 0;
}
