// Implicit-GEMM convolution with the WEIGHTS as the A operand ("weights-as-A"), for layers with >= 128 output
// channels and source channels in multiples of 64:
//
//     D[c_out][pixel] = sum_{block q, tap t, k} Wt[c_out][q, t, k] * X[pixel + offset(t)][64 q + k]
//
//   A = weights: K-major no-swizzle core matrices, 128 output channels per M block, one 16 KB "unit" per
//       (64-channel block, tap) = four K = 16 steps; units stream through an mbarrier ring of bulk copies which,
//       in a thread-block cluster, every CTA fetches a 1/CL slice of and MULTICASTS to its peers (the CTAs of a
//       cluster work on different pixel tiles of the same M block in lockstep, so L2 weight traffic drops by CL);
//   B = pixels: K-major SWIZZLE_128B tiles [virtual pixel][64 channels] (128-byte rows) written by TMA, one
//       box {64 ch, W + pad, 1, 1} per virtual row: padding columns, padding rows and rows outside the batch are
//       out-of-bounds zero fill.  The tile is loaded ONCE per 64-channel block and serves all 25 (9) taps: a
//       tap is the descriptor start moved by dy * PW + dx rows (tcgen05.mma applies the 128-byte swizzle to
//       absolute shared-memory address bits: tools/umma_probe_swz.py sections 5-6, profiles/r02_umma_probe_swz.log);
//   N = up to 256 pixels per MMA: M = 128 x N = 256 x K = 16 takes max(N/2, 32 + N/4) = 128 cycles, i.e. the MMA
//       is math bound and reads shared memory at 96 of 128 B/clk (the pixels-as-M orientation of conv_gemm.cu
//       with N <= 128 needs all 128 B/clk and is slowed down by every other shared-memory access).
//
// A tile is a range of virtual rows (see planes.cuh for the virtual pixel space): its outputs are the pixels of
// those rows (N = rows * PW, rounded up to 16), its loads are those rows plus `pad` rows on either side.  The
// virtual rows of the batch are dealt evenly to the CTAs, each CTA cuts its share into `nt` tiles.
//
// Roles (7 warps): warps 0-3 epilogue (TMEM lane = output channel, column = pixel: BatchNorm statistics are
// per-thread sums), warp 4 pixel-tile TMA producer, warp 5 weight producer, warp 6 MMA issuer (one elected
// thread).  Two 256-column accumulators: the epilogue of item i overlaps the MMAs of item i + 1.
//
// Reference ops replaced: the same as conv_gemm.cu (nn.Conv2d at vae_nets.py:79,84,117,121,125 forward, and the
// data gradients of :84,117,121), selected by the caller with CVAE_KTAB_BLOCK64 (block-major packed weights).
#include "common.cuh"
#include "umma.cuh"
#include "tma.cuh"

namespace cvae {

static constexpr int kWaThreads = 224;
static constexpr int kWaMaxStages = 8, kWaMaxSlots = 6;
static constexpr int kWaMarginLo = 8;       // zero rows in front of the loaded rows (left padding of the first row)
static constexpr int kWaUnitBytes = 16384;  // 128 rows x 64 K x bf16

struct WaArgs {
    int B, H, W, pad, KW, PW, IH;
    int taps, nblk, units;        // units = nblk * taps per (tile, M block)
    int m_blocks, c_total;
    int T0, T;                    // first virtual row with outputs, number of virtual rows to cover
    int nt;                       // tiles per CTA
    int slot_bytes, nslots;       // pixel-tile ring (one 64-channel block per slot)
    int ups, nstages, stage_bytes;  // weight ring: units per stage
    int phase_src;                // 1: block q comes from tensor map q (space-to-depth phases), channel 0
    int epilogue;
    const __nv_bfloat16* wpack;   // [m_blocks][units][4][128 x 16]
    void* out;
    const float* bias;
    const __nv_bfloat16* act;
    double* stats;
    int* fault;
    unsigned long long* dbg;
};

struct WaBars {
    uint64_t w_full[kWaMaxStages], w_empty[kWaMaxStages];
    uint64_t b_full[kWaMaxSlots], b_empty[kWaMaxSlots];
    uint64_t acc_full[2], acc_empty[2];
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 1-D bulk copy global -> the same shared-memory offset of every CTA in `mask`; each destination CTA's mbarrier (same offset) gets the bytes
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
// arrive on the barrier at the same offset in every CTA of `mask` once all previously issued MMAs have retired
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                     "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

// Tile t of this CTA: virtual rows [ra, ra + nrows).  The T rows are dealt evenly to the gridDim.x CTAs, every CTA
// cuts its share into nt near-equal tiles (a tile may be empty when there are fewer rows than tiles).
__device__ __forceinline__ void wa_tile(const WaArgs& a, int t, int& ra, int& nrows) {
    const long G = gridDim.x, g = blockIdx.x;
    const int lo = (int)((long)a.T * g / G), hi = (int)((long)a.T * (g + 1) / G);
    const int n = hi - lo;
    const int b0 = (int)((long)n * t / a.nt), b1 = (int)((long)n * (t + 1) / a.nt);
    ra = a.T0 + lo + b0;
    nrows = b1 - b0;
}
__device__ __forceinline__ int wa_n(const WaArgs& a, int nrows) {
    int n = (nrows * a.PW + 15) & ~15;
    return n < 16 ? 16 : n;
}

// The 8 lanes that share (lane >> 3) each hold one channel (lane & 7) of 8 consecutive pixels, R[m] = bf16 pair of
// pixels (2m, 2m + 1).  On return lane j holds pixel j: R[k] = bf16 pair of channels (2k, 2k + 1), i.e. the 16 bytes
// an NHWC store wants.  Three butterfly stages, 8 shuffles (simulated in numpy before it ever ran on a GPU).
__device__ __forceinline__ void transpose8x8_bf16(uint32_t (&R)[4], int lane) {
    const bool hi4 = (lane & 4) != 0, hi2 = (lane & 2) != 0;
    uint32_t s0 = hi4 ? R[0] : R[2], s1 = hi4 ? R[1] : R[3];
    uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 4), r1 = __shfl_xor_sync(0xffffffffu, s1, 4);
    const uint32_t X0 = hi4 ? r0 : R[0], X1 = hi4 ? r1 : R[1], X2 = hi4 ? R[2] : r0, X3 = hi4 ? R[3] : r1;
    s0 = hi2 ? X0 : X1;
    s1 = hi2 ? X2 : X3;
    r0 = __shfl_xor_sync(0xffffffffu, s0, 2);
    r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
    const uint32_t Y0 = hi2 ? r0 : X0, Y1 = hi2 ? X1 : r0, Y2 = hi2 ? r1 : X2, Y3 = hi2 ? X3 : r1;
    const uint32_t sel = (lane & 1) ? 0x3276u : 0x5410u;
    R[0] = __byte_perm(Y0, __shfl_xor_sync(0xffffffffu, Y0, 1), sel);
    R[1] = __byte_perm(Y1, __shfl_xor_sync(0xffffffffu, Y1, 1), sel);
    R[2] = __byte_perm(Y2, __shfl_xor_sync(0xffffffffu, Y2, 1), sel);
    R[3] = __byte_perm(Y3, __shfl_xor_sync(0xffffffffu, Y3, 1), sel);
}

// Epilogue of 8 consecutive output pixels (pix0 .. pix0 + 7, the first `nvalid` real) for the warp's 32 channels
// [cbase, cbase + 32): v[i] is this lane's channel at pixel pix0 + i (zero for i >= nvalid).
template <int EPI>
__device__ __forceinline__ void wa_emit8(const WaArgs& a, float (&v)[8], int nvalid, int pix0, int cbase, int lane, float bias,
                                         float& t1, float& t2) {
    if constexpr (EPI == CVAE_EPI_BIAS_RELU || EPI == CVAE_EPI_PHASE_BIAS_RELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i] + bias, 0.f);
    }
    uint32_t R[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) R[m] = pack_bf16x2(v[2 * m], v[2 * m + 1]);
    if constexpr (EPI == CVAE_EPI_STATS) {   // BatchNorm statistics of the stored (bf16-rounded) values; lane = channel
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const float lo = bf16_lo(R[m]), hi = bf16_hi(R[m]);
            t1 += lo + hi;
            t2 = fmaf(lo, lo, fmaf(hi, hi, t2));
        }
    }
    transpose8x8_bf16(R, lane);
    const int j = lane & 7, cg = cbase + (lane >> 3) * 8;
    if (j < nvalid) {
        const int pix = pix0 + j;
        size_t off;
        if constexpr (EPI == CVAE_EPI_PHASE_BIAS_RELU) {
            const int cout4 = a.c_total >> 2, ph = cg / cout4, oc = cg - ph * cout4;
            const int n = pix / (a.H * a.W), rem = pix - n * a.H * a.W, h = rem / a.W, w = rem - h * a.W;
            off = ((size_t)(n * 2 * a.H + 2 * h + (ph >> 1)) * (2 * a.W) + 2 * w + (ph & 1)) * cout4 + oc;
        } else {
            off = (size_t)pix * a.c_total + cg;
        }
        if constexpr (EPI == CVAE_EPI_MASK) {   // ReLU backward: keep the gradient where the saved activation is positive
            const uint4 m = __ldg(reinterpret_cast<const uint4*>(a.act + off));
            const uint32_t mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!(bf16_lo(mm[k]) > 0.f)) R[k] &= 0xFFFF0000u;
                if (!(bf16_hi(mm[k]) > 0.f)) R[k] &= 0x0000FFFFu;
            }
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + off) = make_uint4(R[0], R[1], R[2], R[3]);
    }
}

template <int EPI, int CL>
__global__ void __launch_bounds__(kWaThreads, 1)
conv_wa_kernel(const WaArgs a, const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
               const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ WaBars bars;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* const bring = smem;                                       // [nslots][slot_bytes]
    uint8_t* const wring = smem + (size_t)a.nslots * a.slot_bytes;     // [nstages][stage_bytes]
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

    if (tid == 0) {
        for (int s = 0; s < a.nstages; ++s) { mbar_init(&bars.w_full[s], 1); mbar_init(&bars.w_empty[s], CL); }
        for (int s = 0; s < a.nslots; ++s) { mbar_init(&bars.b_full[s], 1); mbar_init(&bars.b_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bars.acc_full[b], 1); mbar_init(&bars.acc_empty[b], 4); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    // the margins around the TMA-written rows must read as zero: clear the whole pixel ring once
    {
        const int n16 = (a.nslots * a.slot_bytes) >> 4;
        for (int i = tid; i < n16; i += kWaThreads) reinterpret_cast<uint4*>(bring)[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
    tc_fence_before();
    if constexpr (CL > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int items = a.nt * a.m_blocks;

    if (warp < 4) {
        // ================================ epilogue ================================================
        // TMEM lane = output channel (this thread's), column = pixel of the tile.  Only the W real pixels of the real
        // rows are read back; they are consecutive in the NHWC output, so the tile leaves as 8-pixel groups whose
        // 8 x 8 (pixel x channel) blocks are transposed with shuffles into 16-byte stores.
        float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
        for (int it = 0; it < items; ++it) {
            const int t = it / a.m_blocks, mb = it - t * a.m_blocks;
            int ra, nrows;
            wa_tile(a, t, ra, nrows);
            const uint32_t ab = (uint32_t)it & 1u;
            const int cbase = mb * 128 + warp * 32;
            float bias = 0.f;
            if constexpr (EPI == CVAE_EPI_BIAS_RELU) bias = __ldg(a.bias + cbase + lane);
            if constexpr (EPI == CVAE_EPI_PHASE_BIAS_RELU) bias = __ldg(a.bias + (cbase + lane) % (a.c_total >> 2));
            mbar_wait(&bars.acc_full[ab], ((uint32_t)it >> 1) & 1u, a.fault);
            tc_fence_after();
            const uint32_t tbase = tmem_base + ab * 256u + ((uint32_t)(warp * 32) << 16);
            float t1 = 0.f, t2 = 0.f;
            int n = ra / a.IH, r = ra - n * a.IH;
            float pend[4];           // W == 4: the first half of an 8-pixel group waits for the next real row
            int pend_pix = -1;
            for (int tr = 0; tr < nrows; ++tr) {
                if (r >= a.pad && n < a.B) {
                    const int pixbase = (n * a.H + r - a.pad) * a.W;
                    const uint32_t col = tbase + (uint32_t)(tr * a.PW);
                    if (a.W >= 16) {
                        for (int c = 0; c < a.W; c += 16) {
                            uint32_t raw[16];
                            tmem_ld16(col + (uint32_t)c, raw);
                            tmem_wait_ld();
                            float v[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[i]);
                            wa_emit8<EPI>(a, v, 8, pixbase + c, cbase, lane, bias, t1, t2);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[8 + i]);
                            wa_emit8<EPI>(a, v, 8, pixbase + c + 8, cbase, lane, bias, t1, t2);
                        }
                    } else if (a.W == 8) {
                        uint32_t raw[8];
                        tmem_ld8(col, raw);
                        tmem_wait_ld();
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[i]);
                        wa_emit8<EPI>(a, v, 8, pixbase, cbase, lane, bias, t1, t2);
                    } else {   // W == 4
                        uint32_t raw[4];
                        tmem_ld4(col, raw);
                        tmem_wait_ld();
                        if (pend_pix < 0) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) pend[i] = __uint_as_float(raw[i]);
                            pend_pix = pixbase;
                        } else {   // real rows are consecutive in the output: pend_pix + 4 == pixbase
                            float v[8];
#pragma unroll
                            for (int i = 0; i < 4; ++i) { v[i] = pend[i]; v[4 + i] = __uint_as_float(raw[i]); }
                            wa_emit8<EPI>(a, v, 8, pend_pix, cbase, lane, bias, t1, t2);
                            pend_pix = -1;
                        }
                    }
                }
                if (++r == a.IH) { r = 0; ++n; }
            }
            if (pend_pix >= 0) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) { v[i] = pend[i]; v[4 + i] = 0.f; }
                if constexpr (EPI == CVAE_EPI_BIAS_RELU || EPI == CVAE_EPI_PHASE_BIAS_RELU) {
#pragma unroll
                    for (int i = 4; i < 8; ++i) v[i] = -bias;   // stays zero through bias + ReLU (never stored anyway)
                }
                wa_emit8<EPI>(a, v, 4, pend_pix, cbase, lane, bias, t1, t2);
            }
            s1[mb & 1] += t1;
            s2[mb & 1] += t2;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.acc_empty[ab]);
        }
        if constexpr (EPI == CVAE_EPI_STATS) {
            for (int mb = 0; mb < a.m_blocks; ++mb) {
                const int co = mb * 128 + warp * 32 + lane;
                atomicAdd(a.stats + co, (double)s1[mb & 1]);
                atomicAdd(a.stats + a.c_total + co, (double)s2[mb & 1]);
            }
        }
    } else if (warp == 4) {
        // ================================ pixel-tile producer (TMA) ================================
        if (elect_one()) {
            tma_prefetch_desc(&map0);
            uint32_t seq = 0;
            bool alive = true;
            for (int it = 0; it < items && alive; ++it) {
                const int t = it / a.m_blocks;
                int ra, nrows;
                wa_tile(a, t, ra, nrows);
                const int nload = nrows + 2 * a.pad;
                for (int q = 0; q < a.nblk && alive; ++q, ++seq) {
                    const uint32_t slot = seq % (uint32_t)a.nslots;
                    alive = mbar_wait(&bars.b_empty[slot], ((seq / (uint32_t)a.nslots) & 1u) ^ 1u, a.fault);
                    mbar_expect_tx(&bars.b_full[slot], (uint32_t)(nload * a.PW) * 128u);
                    const CUtensorMap* mp = (!a.phase_src || q == 0) ? &map0 : (q == 1 ? &map1 : (q == 2 ? &map2 : &map3));
                    const int c0 = a.phase_src ? 0 : q * 64;
                    uint32_t dst = smem_u32(bring) + slot * (uint32_t)a.slot_bytes + (uint32_t)kWaMarginLo * 128u;
                    const int vr = ra - a.pad;             // >= 0: outputs start at virtual row `pad`
                    int n = vr / a.IH, r = vr - n * a.IH;
                    for (int j = 0; j < nload; ++j) {
                        // a padding row (r < pad) or a row past the batch (n >= B) is entirely out of bounds: zero fill
                        tma_load_4d(dst, mp, c0, 0, r - a.pad, n, &bars.b_full[slot]);
                        dst += (uint32_t)a.PW * 128u;
                        if (++r == a.IH) { r = 0; ++n; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ================================ weight producer ==========================================
        if (elect_one()) {
            const uint8_t* wbase = reinterpret_cast<const uint8_t*>(a.wpack);
            const int spi = (a.units + a.ups - 1) / a.ups;   // stages per item
            uint32_t s = 0;
            bool alive = true;
            for (int it = 0; it < items && alive; ++it) {
                const int mb = it % a.m_blocks;
                const uint8_t* src = wbase + (size_t)mb * a.units * kWaUnitBytes;
                for (int st = 0; st < spi && alive; ++st, ++s) {
                    const uint32_t slot = s % (uint32_t)a.nstages;
                    alive = mbar_wait(&bars.w_empty[slot], ((s / (uint32_t)a.nstages) & 1u) ^ 1u, a.fault);
                    const int nu = min(a.ups, a.units - st * a.ups);
                    const uint32_t bytes = (uint32_t)nu * kWaUnitBytes;
                    mbar_expect_tx(&bars.w_full[slot], bytes);
                    uint8_t* dst = wring + (size_t)slot * a.stage_bytes;
                    const uint8_t* from = src + (size_t)st * a.ups * kWaUnitBytes;
                    if constexpr (CL == 1) {
                        bulk_g2s(dst, from, bytes, &bars.w_full[slot]);
                    } else {
                        const uint32_t slice = bytes / CL;
                        bulk_g2s_multicast(dst + rank * slice, from + rank * slice, slice, &bars.w_full[slot], kMask);
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ================================ MMA issuer ===============================================
        if (elect_one()) {
            const uint32_t a_hi = (256u >> 4) | (1u << 14);                 // weights: SBO 256 B, no swizzle
            const uint32_t a_lbo = (128u >> 4) << 16;
            const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);    // pixels: SBO 1024 B, SWIZZLE_128B
            const uint32_t b_lbo = 1u << 16;                                  // ignored for swizzled K-major
            const uint32_t bring16 = (smem_u32(bring) & 0x3FFFFu) >> 4, wring16 = (smem_u32(wring) & 0x3FFFFu) >> 4;
            const int out_row0 = kWaMarginLo + a.pad * a.PW;                  // ring row of the tile's first output pixel
            uint32_t seq = 0, s = 0;
            bool alive = true;
            long long t_acc = 0, t_b = 0, t_w = 0, tq = 0;
            const bool prof = a.dbg != nullptr;
            const long long t0 = prof ? clock64() : 0;
            unsigned long long g0 = 0;
            if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
            for (int it = 0; it < items && alive; ++it) {
                const int t = it / a.m_blocks;
                int ra, nrows;
                wa_tile(a, t, ra, nrows);
                const uint32_t idesc = umma_idesc_bf16((uint32_t)wa_n(a, nrows), kMajorK, kMajorK);
                const uint32_t ab = (uint32_t)it & 1u;
                if (prof) tq = clock64();
                alive = mbar_wait(&bars.acc_empty[ab], (((uint32_t)it >> 1) & 1u) ^ 1u, a.fault);
                if (prof) t_acc += clock64() - tq;
                tc_fence_after();
                const uint32_t acc = tmem_base + ab * 256u;
                uint32_t accumulate = 0;
                int u = 0, uin = 0;      // unit inside the item / inside the weight stage
                uint32_t wslot = 0;
                for (int q = 0; q < a.nblk && alive; ++q, ++seq) {
                    const uint32_t slot = seq % (uint32_t)a.nslots;
                    if (prof) tq = clock64();
                    alive = mbar_wait(&bars.b_full[slot], (seq / (uint32_t)a.nslots) & 1u, a.fault);
                    if (prof) t_b += clock64() - tq;
                    tc_fence_after();
                    const int b_row0 = (int)(slot * (uint32_t)(a.slot_bytes >> 7)) + out_row0;
                    int ty = 0, tx = 0;
                    for (int tap = 0; tap < a.taps && alive; ++tap, ++u) {
                        if (uin == 0) {
                            wslot = s % (uint32_t)a.nstages;
                            if (prof) tq = clock64();
                            alive = mbar_wait(&bars.w_full[wslot], (s / (uint32_t)a.nstages) & 1u, a.fault);
                            if (prof) t_w += clock64() - tq;
                            tc_fence_after();
                        }
                        const uint32_t a_lo = (wring16 + ((wslot * (uint32_t)a.stage_bytes + (uint32_t)uin * kWaUnitBytes) >> 4)) | a_lbo;
                        const int row = b_row0 + (ty - a.pad) * a.PW + (tx - a.pad);
                        const uint32_t b_lo = (bring16 + (uint32_t)row * 8u) | b_lbo;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            umma_bf16(acc, ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)(j * 256)),
                                      ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)(j * 2)), idesc, j == 0 ? accumulate : 1u);
                        accumulate = 1u;
                        if (++tx == a.KW) { tx = 0; ++ty; }
                        if (++uin == a.ups || u + 1 == a.units) {
                            if constexpr (CL == 1) umma_commit(&bars.w_empty[wslot]);
                            else umma_commit_multicast(&bars.w_empty[wslot], kMask);
                            uin = 0;
                            ++s;
                        }
                    }
                    umma_commit(&bars.b_empty[slot]);
                }
                umma_commit(&bars.acc_full[ab]);
            }
            if (prof) {
                unsigned long long* o = a.dbg + (size_t)blockIdx.x * 8;
                unsigned long long g1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
                o[0] = (unsigned long long)(clock64() - t0); o[1] = t_acc; o[2] = t_b; o[3] = t_w; o[4] = items; o[5] = g1 - g0;
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    if constexpr (CL > 1) cluster_sync_all(); else __syncthreads();
    if (warp == 0) tmem_free(tmem_base, 512);
}

static const bool g_wa_debug = getenv("CVAE_DEBUG") != nullptr;
// tuning overrides (0 = automatic), set through cvae_conv_wa_tune (tests sweep them, tools/conv_bench.py explores them)
static int g_wa_cluster = 0, g_wa_grid = 0, g_wa_ups = 0, g_wa_nt = 0;
unsigned long long* g_wa_dbg = nullptr;

template <int EPI, int CL>
static int launch_wa(const WaArgs& a, const CUtensorMap* maps, int grid, size_t smem, cudaStream_t stream) {
    CVAE_OPT_IN_SMEM((conv_wa_kernel<EPI, CL>), smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kWaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    CVAE_CUDA(cudaLaunchKernelEx(&cfg, conv_wa_kernel<EPI, CL>, a, maps[0], maps[1], maps[2], maps[3]));
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

// Called by cvae_conv_gemm for descriptors with ktab == CVAE_KTAB_BLOCK64.
int conv_wa_dispatch(const cvae_conv_desc* d, cudaStream_t stream) {
    CVAE_REQUIRE(d->src_channels % 64 == 0 && d->src_channels <= 256, CVAE_EINVAL, "conv_wa: src_channels %d", d->src_channels);
    CVAE_REQUIRE(d->n_total % 128 == 0 && d->n_total <= 256, CVAE_EINVAL, "conv_wa: n_total %d", d->n_total);
    CVAE_REQUIRE(d->loader == CVAE_LOAD_NHWC || (d->loader == CVAE_LOAD_S2D && d->src_channels == 256), CVAE_EINVAL,
                 "conv_wa: loader %d with %d channels", d->loader, d->src_channels);
    CVAE_REQUIRE(d->epilogue == CVAE_EPI_STATS || d->epilogue == CVAE_EPI_BIAS_RELU || d->epilogue == CVAE_EPI_PHASE_BIAS_RELU ||
                     d->epilogue == CVAE_EPI_MASK || d->epilogue == CVAE_EPI_PLAIN, CVAE_EINVAL, "conv_wa: epilogue %d", d->epilogue);
    CVAE_REQUIRE(d->epilogue != CVAE_EPI_STATS || d->stats, CVAE_EINVAL, "conv_wa: statistics buffer missing");
    CVAE_REQUIRE((d->epilogue != CVAE_EPI_BIAS_RELU && d->epilogue != CVAE_EPI_PHASE_BIAS_RELU) || d->bias, CVAE_EINVAL, "conv_wa: bias missing");
    CVAE_REQUIRE(d->epilogue != CVAE_EPI_MASK || d->act, CVAE_EINVAL, "conv_wa: activation (mask) tensor missing");
    CVAE_REQUIRE(d->width == 4 || d->width == 8 || d->width % 16 == 0, CVAE_EINVAL, "conv_wa: width %d (4, 8 or a multiple of 16)", d->width);
    CVAE_REQUIRE(d->epilogue != CVAE_EPI_PHASE_BIAS_RELU || (d->n_total / 4) % 8 == 0, CVAE_EINVAL, "conv_wa: phase epilogue needs n_total / 4 in multiples of 8");
    CVAE_REQUIRE(tensor_map_encoder() != nullptr, CVAE_ECUDA, "conv_wa: cuTensorMapEncodeTiled unavailable");

    WaArgs a{};
    a.B = d->batch; a.H = d->height; a.W = d->width; a.KW = d->ksize; a.pad = d->ksize / 2;
    a.PW = a.W + a.pad; a.IH = a.H + a.pad;
    CVAE_REQUIRE(a.PW * 1 <= 256 && (long)a.B * a.IH * a.PW < (1L << 30), CVAE_EINVAL, "conv_wa: map too large");
    a.taps = a.KW * a.KW;
    a.nblk = d->src_channels / 64;
    a.units = a.nblk * a.taps;
    a.m_blocks = d->n_total / 128;
    a.c_total = d->n_total;
    a.T0 = a.pad;
    a.T = a.B * a.IH - a.pad;
    a.phase_src = d->loader == CVAE_LOAD_S2D ? 1 : 0;
    a.epilogue = d->epilogue;
    a.wpack = (const __nv_bfloat16*)d->wpack; a.out = d->out; a.bias = d->bias; a.act = (const __nv_bfloat16*)d->act; a.stats = d->stats;
    a.fault = fault_flag();
    a.dbg = g_wa_dbg;
    CVAE_REQUIRE(a.fault != nullptr, CVAE_ECUDA, "conv_wa: fault flag unavailable");

    // ---- work decomposition -----------------------------------------------------------------------------------
    // Choose the grid size and the cluster size from a small cost model: per CTA, MMA time = MMAs x max(N/2, 32 + N/4)
    // cycles against the time to pull this layer's weight stream out of L2 (every CTA streams the whole weight
    // matrix of its M blocks once per tile; a cluster shares one stream), plus a per-tile pipeline cost.
    const int sms = sm_count();
    const int rows_max = 256 / a.PW;                          // rows per tile so that N <= 256
    int grid = 1, cl = 1;
    {
        double best = 1e30;
        const double l2_bytes_per_clk = 2800.0;                // chip-wide, conservative (B300_MICROARCH.md: ~6300 B/clk peak)
        for (int c : {1, 2, 4}) {
            if (g_wa_cluster > 0 && c != g_wa_cluster) continue;
            const int usable = c == 4 ? sms - sms % 4 - 16 : sms - sms % c;   // clusters of 4 strand ~16 SMs (GPC shapes)
            for (int g = c; g <= usable; g += c) {
                if (g_wa_grid > 0 && g != g_wa_grid - g_wa_grid % c) continue;
                const int rows = (a.T + g - 1) / g;
                if (rows * a.PW < 48 && g > c && g_wa_grid == 0) continue;   // keep >= ~48 pixels per CTA
                int nt = (rows + rows_max - 1) / rows_max;
                if (g_wa_nt > nt) nt = g_wa_nt;
                const int tr = (rows + nt - 1) / nt;
                const double n = (double)((tr * a.PW + 15) & ~15);
                const double mma = (n / 2 > 32 + n / 4 ? n / 2 : 32 + n / 4) * 1.1;
                const double t_mma = (double)nt * a.m_blocks * a.units * 4.0 * mma;
                const double t_l2 = (double)g / c * nt * a.m_blocks * a.units * kWaUnitBytes / l2_bytes_per_clk;
                const double est = (t_mma > t_l2 ? t_mma : t_l2) + 4000.0 + 1500.0 * nt * a.m_blocks;
                if (est < best) { best = est; grid = g; cl = c; }
            }
        }
        CVAE_REQUIRE(best < 1e30, CVAE_EINVAL, "conv_wa: no launch configuration (cluster %d, grid %d)", g_wa_cluster, g_wa_grid);
    }
    const int rows_cta = (a.T + grid - 1) / grid;              // largest share
    a.nt = (rows_cta + rows_max - 1) / rows_max;
    if (g_wa_nt > a.nt) a.nt = g_wa_nt;
    const int tile_rows = (rows_cta + a.nt - 1) / a.nt;        // largest tile
    // pixel ring: margin, loaded rows, tail margin (N rounding + pad + slack), 1024-byte aligned slots
    const int slot_rows = kWaMarginLo + (tile_rows + 2 * a.pad) * a.PW + 16 + a.pad + 8;
    a.slot_bytes = (slot_rows * 128 + 1023) & ~1023;
    const size_t cap = 214 * 1024;
    a.nslots = a.nblk >= 3 ? 3 : 2;
    a.ups = g_wa_ups > 0 ? g_wa_ups : 2;
    for (;;) {
        a.stage_bytes = a.ups * kWaUnitBytes;
        const long left = (long)cap - (long)a.nslots * a.slot_bytes;
        a.nstages = (int)(left / a.stage_bytes);
        if (a.nstages >= 3 || a.ups == 1) break;
        a.ups = 1;
    }
    if (a.nstages > kWaMaxStages) a.nstages = kWaMaxStages;
    CVAE_REQUIRE(a.nstages >= 2, CVAE_EINVAL, "conv_wa: shape does not fit shared memory");
    // spend what is left on more pixel slots (deeper prefetch across tiles)
    while (a.nslots < kWaMaxSlots && a.nslots < 2 * a.nblk &&
           (size_t)(a.nslots + 1) * a.slot_bytes + (size_t)a.nstages * a.stage_bytes <= cap) ++a.nslots;
    const size_t smem = (size_t)a.nslots * a.slot_bytes + (size_t)a.nstages * a.stage_bytes;
    CVAE_REQUIRE((size_t)a.nslots * a.slot_bytes + (size_t)a.nstages * a.stage_bytes < (1u << 18), CVAE_EINVAL, "conv_wa: descriptor range");

    // ---- tensor maps: one row box {64 ch, PW, 1, 1} per call ----------------------------------------------------
    CUtensorMap maps[4];
    bool ok = true;
    if (!a.phase_src) {
        const long C = d->src_channels;
        ok = encode_map_4d(&maps[0], d->src, (int)C, a.W, a.H, a.B, C, (long)a.W * C, (long)a.H * a.W * C, 64, a.PW, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B);
        maps[1] = maps[2] = maps[3] = maps[0];
    } else {   // source [B][2H][2W][64]: one strided view per phase (a, b) = block q
        const long C = 64;
        for (int ab = 0; ab < 4 && ok; ++ab) {
            const __nv_bfloat16* b0 = (const __nv_bfloat16*)d->src + ((long)(ab >> 1) * 2 * a.W + (ab & 1)) * C;
            ok = encode_map_4d(&maps[ab], b0, (int)C, a.W, a.H, a.B, 2 * C, 2L * 2 * a.W * C, 4L * a.H * a.W * C, 64, a.PW, 1, 1,
                               CU_TENSOR_MAP_SWIZZLE_128B);
        }
    }
    CVAE_REQUIRE(ok, CVAE_ECUDA, "conv_wa: cuTensorMapEncodeTiled failed");
    if (g_wa_debug)
        fprintf(stderr, "conv_wa E%d B=%d %dx%d k%d C%d->N%d: grid=%d cluster=%d nt=%d tile_rows=%d (N<=%d) units=%d ups=%d stages=%d slots=%d x %d B smem=%zu\n",
                d->epilogue, a.B, a.H, a.W, a.KW, d->src_channels, d->n_total, grid, cl, a.nt, tile_rows, (tile_rows * a.PW + 15) & ~15, a.units,
                a.ups, a.nstages, a.nslots, a.slot_bytes, smem);
#define CVAE_WA_CASE(E_)                                                             \
    if (d->epilogue == (E_)) {                                                       \
        if (cl == 4) return launch_wa<E_, 4>(a, maps, grid, smem, stream);           \
        if (cl == 2) return launch_wa<E_, 2>(a, maps, grid, smem, stream);           \
        return launch_wa<E_, 1>(a, maps, grid, smem, stream);                        \
    }
    CVAE_WA_CASE(CVAE_EPI_STATS)
    CVAE_WA_CASE(CVAE_EPI_BIAS_RELU)
    CVAE_WA_CASE(CVAE_EPI_PHASE_BIAS_RELU)
    CVAE_WA_CASE(CVAE_EPI_MASK)
    CVAE_WA_CASE(CVAE_EPI_PLAIN)
#undef CVAE_WA_CASE
    return CVAE_EINVAL;
}

}  // namespace cvae

// Profiling aid: per-CTA cycle counters of the MMA thread (total, wait accumulator / pixels / weights, items).
extern "C" void cvae_conv_wa_debug_counters(void* device_buf) { cvae::g_wa_dbg = (unsigned long long*)device_buf; }
// Tuning / test hook: force the cluster size (1, 2, 4), the grid size, the (block, tap) units per weight stage and a
// minimum number of tiles per CTA of the weights-as-A kernel; 0 restores the automatic choice.  Process-wide.
extern "C" void cvae_conv_wa_tune(int cluster, int grid, int units_per_stage, int tiles_per_cta) {
    cvae::g_wa_cluster = cluster; cvae::g_wa_grid = grid; cvae::g_wa_ups = units_per_stage; cvae::g_wa_nt = tiles_per_cta;
}
