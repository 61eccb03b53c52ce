// PROTOTYPE (round-2 groundwork, written without a GPU at hand; validate with tools/conv_swap_proto.py before use).
//
// Convolution forward with the operand roles swapped relative to conv_pipe_kernel:
//   D[c_out][pixel] = sum_{tap, c_in} W[c_out][tap, c_in] * X[pixel + tap offset][c_in]
//   A = weights, K-major no-swizzle core matrices (the [kstep][128][16] blocks cvae_pack_weights already produces),
//       128 output channels per M block, streamed through a bulk-copy ring;
//   B = pixels, K-major SWIZZLE_128B tile [virtual pixel][64 channels] written by TMA, one box per virtual row
//       ({64 ch, W + pad, 1, 1}: padding columns, padding rows and rows outside the batch all come from
//       out-of-bounds zero fill, so every row of the tile is written by TMA); N = 256 pixels per tile;
//       a filter tap is the descriptor start moved by dy * PW + dx rows (tools/umma_probe_swz.py: the swizzle is
//       applied to absolute address bits), a K = 16 step is +32 B inside the 128-byte row.
// One M=128 x N=256 x K=16 MMA takes max(N/2, 32 + N/4) = 128 cycles: math bound, unlike the N <= 128 MMAs of the
// pixel-as-M orientation.  One CTA per (256-pixel tile, M block), not persistent, naive descriptor arithmetic:
// this file is about correctness of the scheme and a first timing, not the final kernel.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC -o tools/libconv_swap_proto.so tools/conv_swap_proto.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>

#include "../critic-vae_b200/csrc/umma.cuh"

using namespace cvae;

__device__ int g_proto_fault = 0;

struct ProtoArgs {
    int B, H, W, pad, KW, PW, IH;
    int cin, cout, cblocks;      // cin % 64 == 0, cout % 128 == 0, cblocks = cin / 64
    int ksteps, kpt;             // K = 16 steps in total, per tap (cin / 16)
    int halo;                    // pad * PW + pad
    int rows_per_tile;           // virtual rows a tile needs
    int block_bytes;             // one 64-channel block of the pixel tile (multiple of 1024)
    int ksps, nstages;           // weight ring: K steps per stage, stages
    int relu;
    const __nv_bfloat16* wpack;  // [cout/128][ksteps][128][16]
    const float* bias;           // [cout] or NULL
    __nv_bfloat16* out;          // NHWC [B][H][W][cout]
};

static constexpr int kTileN = 256;
static constexpr int kProtoThreads = 192;   // warp 0: TMA + weight producer, warp 1: MMA issuer, warps 2-5: epilogue

__device__ __forceinline__ void tma_row_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

__global__ void __launch_bounds__(kProtoThreads, 1) conv_swap_kernel(const ProtoArgs a, const __grid_constant__ CUtensorMap mapX) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t act_full, acc_full, w_full[8], w_empty[8];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, mb = blockIdx.y;
    const uint32_t stage_bytes = (uint32_t)a.ksps * 4096u;
    uint8_t* tile_smem = smem;                                            // [cblocks][rows_per_tile * PW][128 B]
    uint8_t* wring = smem + (size_t)a.cblocks * a.block_bytes;           // [nstages][ksps][4096]

    if (tid == 0) {
        mbar_init(&act_full, 1);
        mbar_init(&acc_full, 1);
        for (int s = 0; s < a.nstages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, kTileN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    // virtual pixel geometry of this tile
    const int v0 = a.pad * a.PW + tile * kTileN;                          // first output pixel
    const int vr0 = (v0 - a.halo + 8 * a.PW) / a.PW - 8;                  // first virtual row needed (may be -1)
    const int slot0 = v0 - vr0 * a.PW;                                    // tile slot of pixel v0
    const int nstage_total = a.ksteps / a.ksps;

    if (warp == 0) {
        if (elect_one()) {
            // ---- activations: one box per (virtual row, 64-channel block) ----
            mbar_expect_tx(&act_full, (uint32_t)(a.rows_per_tile * a.cblocks * a.PW * 128));
            for (int j = 0; j < a.rows_per_tile; ++j) {
                const int vr = vr0 + j;
                int n, h;
                if (vr < 0) { n = -1; h = 0; }                            // before the first image: any out-of-bounds coordinate zero-fills
                else { n = vr / a.IH; h = vr - n * a.IH - a.pad; }
                for (int q = 0; q < a.cblocks; ++q)
                    tma_row_4d(smem_u32(tile_smem) + (uint32_t)q * a.block_bytes + (uint32_t)(j * a.PW) * 128u, &mapX, q * 64, 0, h, n,
                               &act_full);
            }
            // ---- weights: ring of bulk copies ----
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wpack) + (size_t)mb * a.ksteps * 4096;
            bool alive = true;
            for (int s = 0; s < nstage_total && alive; ++s) {
                const int slot = s % a.nstages;
                alive = mbar_wait(&w_empty[slot], ((s / a.nstages) & 1) ^ 1, &g_proto_fault);
                mbar_expect_tx(&w_full[slot], stage_bytes);
                bulk_g2s(wring + (size_t)slot * stage_bytes, wsrc + (size_t)s * stage_bytes, stage_bytes, &w_full[slot]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(kTileN, kMajorK, kMajorK);
            const uint32_t a_hi = (256u >> 4) | (1u << 14);                       // weights: SBO 256, no swizzle
            const uint32_t a_lbo = (128u >> 4) << 16;
            const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);          // pixels: SBO 1024, SWIZZLE_128B
            const uint32_t b_lbo = 1u << 16;                                        // ignored for swizzled K-major
            bool alive = mbar_wait(&act_full, 0, &g_proto_fault);
            tc_fence_after();
            const uint32_t tile16 = (smem_u32(tile_smem) & 0x3FFFFu) >> 4;
            const uint32_t wring16 = (smem_u32(wring) & 0x3FFFFu) >> 4;
            for (int s = 0; s < nstage_total && alive; ++s) {
                const int slot = s % a.nstages;
                alive = mbar_wait(&w_full[slot], (s / a.nstages) & 1, &g_proto_fault);
                tc_fence_after();
                for (int ks = 0; ks < a.ksps; ++ks) {
                    const int k = s * a.ksps + ks;
                    const int tap = k / a.kpt, cp = k - tap * a.kpt;
                    const int dy = tap / a.KW - a.pad, dx = tap % a.KW - a.pad;
                    const int blk = cp >> 2, sub = cp & 3;
                    const uint32_t a_lo = (wring16 + (uint32_t)(slot * (int)stage_bytes + ks * 4096) / 16u) | a_lbo;
                    const uint32_t b_lo = (tile16 + (uint32_t)(blk * a.block_bytes) / 16u + (uint32_t)(slot0 + dy * a.PW + dx) * 8u +
                                           (uint32_t)sub * 2u) | b_lbo;
                    umma_bf16(tmem_base, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc, k > 0 ? 1u : 0u);
                }
                umma_commit(&w_empty[slot]);
            }
            umma_commit(&acc_full);
        }
        __syncwarp();
    } else {
        // ---- epilogue: lane = output channel, column = pixel ----
        mbar_wait(&acc_full, 0, &g_proto_fault);
        tc_fence_after();
        const int quarter = warp & 3;
        const int co = mb * 128 + quarter * 32 + lane;
        const float bias = a.bias ? a.bias[co] : 0.f;
        for (int c0 = 0; c0 < kTileN; c0 += 16) {
            uint32_t raw[16];
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, raw);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int v = v0 + c0 + i;
                const int vrow = v / a.PW, vcol = v - vrow * a.PW;
                const int n = vrow / a.IH, r = vrow - n * a.IH;
                if (vcol < a.W && r >= a.pad && n < a.B) {
                    float f = __uint_as_float(raw[i]) + bias;
                    if (a.relu) f = fmaxf(f, 0.f);
                    a.out[(((size_t)n * a.H + (r - a.pad)) * a.W + vcol) * a.cout + co] = __float2bfloat16_rn(f);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, kTileN);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// x: bf16 NHWC [B][H][W][cin]; wpack: cvae_pack_weights(CVAE_PACK_FWD5) output; out: bf16 NHWC [B][H][W][cout].
// Returns 0, or a negative code (-1 bad shape, -2 CUDA error, -3 device-side bounded wait expired, -4 no TMA encoder).
extern "C" int conv_swap_run(int B, int H, int W, int ksize, int cin, int cout, const void* x, const void* wpack, const float* bias,
                             int relu, void* out, float* elapsed_ms, int iters) {
    if (cin % 64 != 0 || cout % 128 != 0 || (ksize != 5 && ksize != 3) || B <= 0) return -1;
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) return -4;
    ProtoArgs a{};
    a.B = B; a.H = H; a.W = W; a.KW = ksize; a.pad = ksize / 2; a.PW = W + a.pad; a.IH = H + a.pad;
    a.cin = cin; a.cout = cout; a.cblocks = cin / 64;
    a.kpt = cin / 16; a.ksteps = ksize * ksize * a.kpt;
    a.halo = a.pad * a.PW + a.pad;
    a.rows_per_tile = (kTileN + 2 * a.halo + a.PW - 1) / a.PW + 2;
    a.block_bytes = (a.rows_per_tile * a.PW * 128 + 1023) & ~1023;
    a.ksps = 4;
    while (a.ksteps % a.ksps != 0) a.ksps /= 2;
    a.nstages = 4;
    a.relu = relu; a.wpack = (const __nv_bfloat16*)wpack; a.bias = bias; a.out = (__nv_bfloat16*)out;
    const size_t smem = (size_t)a.cblocks * a.block_bytes + (size_t)a.nstages * a.ksps * 4096;
    if (smem > 220 * 1024) return -1;
    CUtensorMap map;
    cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)cin * 2, (cuuint64_t)W * cin * 2, (cuuint64_t)H * W * cin * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)a.PW, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -4;
    if (cudaFuncSetAttribute(conv_swap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
    const long total_v = (long)B * a.IH * a.PW - (long)a.pad * a.PW;
    const int tiles = (int)((total_v + kTileN - 1) / kTileN);
    dim3 grid(tiles, cout / 128);
    int zero = 0;
    cudaMemcpyToSymbol(g_proto_fault, &zero, sizeof(int));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    conv_swap_kernel<<<grid, kProtoThreads, smem>>>(a, map);   // warm-up / correctness run
    if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "conv_swap: %s\n", cudaGetErrorString(cudaGetLastError())); return -2; }
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) conv_swap_kernel<<<grid, kProtoThreads, smem>>>(a, map);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) return -2;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (elapsed_ms) *elapsed_ms = iters > 0 ? ms / iters : 0.f;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    int fault = 0;
    cudaMemcpyFromSymbol(&fault, g_proto_fault, sizeof(int));
    return fault ? -3 : 0;
}
