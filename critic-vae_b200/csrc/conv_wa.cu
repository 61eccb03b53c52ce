// Implicit-GEMM convolution with the WEIGHTS as the A operand ("weights-as-A"):
//
//     D[row][pixel] = sum_{block q, tap group g, k} A[row][q, g, k] * X[pixel + goff(g)][kb q + k]
//
//   A = weights: K-major no-swizzle core matrices, 128 GEMM rows per M block, one "unit" per (channel block, tap
//       group) = kb / 16 K = 16 steps (16 KB for kb = 64); units stream through an mbarrier ring of bulk copies
//       (optionally multicast inside a thread-block cluster whose CTAs walk the same units on different pixels);
//   B = pixels: K-major swizzled tiles [virtual pixel][kb channels] (128-byte rows / SWIZZLE_128B for kb = 64,
//       64-byte rows / SWIZZLE_64B for kb = 32) written by TMA, one box {kb ch, W + pad, 1, 1} per virtual row:
//       padding columns, padding rows and rows outside the batch are out-of-bounds zero fill.  The tile is loaded
//       ONCE per channel block and serves every tap: a tap is the descriptor start moved by dy * PW + dx rows
//       (tcgen05.mma applies the swizzle to absolute shared-memory address bits: tools/umma_probe_swz.py sections
//       5-7, profiles/r02_umma_probe_swz.log);
//   N = up to 256 pixels per MMA: M = 128 x N = 256 x K = 16 is math bound (the pixels-as-M orientation of
//       conv_gemm.cu with N <= 128 output channels is bound by shared-memory reads of its 4 KB A tile).
//
// Tap stacking (J = 2, 4) for layers with 64 / 32 output channels: the 128 GEMM rows of a unit are (output channel,
// j) pairs, row (co, j) holding the weights of tap dx = s - j of a group of J horizontally adjacent taps whose B
// operand starts at shift s.  Accumulator row (co, j) at column c then belongs to output pixel c + j, so the
// epilogue adds the J rows of a channel (adjacent TMEM lanes = adjacent threads: two shuffles) with their columns
// shifted by j.  A 5 x 5 filter costs 10 (J = 4) or 15 (J = 2) units per channel block instead of 25 M = 128 MMAs
// that would leave 3/4 (1/2) of the tensor core's rows empty.
//
// Split-K (ksplit = 2, 4) for the 4 x 4 maps, where there are fewer 256-pixel tiles than SMs and every tile needs
// the whole weight matrix: CTAs of one tile take different channel blocks, write fp32 partials to a workspace, and
// the last CTA to arrive (one atomic counter per tile) adds them in a fixed order and runs the epilogue.
//
// A tile is a range of virtual rows (planes.cuh describes the virtual pixel space): its outputs are the pixels of
// those rows (N = rows * PW + J - 1, rounded up to 16), its loads are those rows plus `pad` rows on either side.
// The virtual rows of the batch are dealt evenly to the CTAs, each CTA cuts its share into `nt` tiles.
//
// Roles (11 warps): warps 0-7 epilogue (TMEM lane = GEMM row, column = pixel: BatchNorm statistics are per-thread
// sums; 8 x 8 pixel x channel blocks are transposed with shuffles into 16-byte NHWC stores; two warps per TMEM lane
// quarter, several pixel groups in flight per warp), warp 8 pixel-tile TMA producer, warp 9 weight producer, warp 10
// MMA issuer (one elected thread).  Two 256-column accumulators: the epilogue of item i overlaps the MMAs of item i + 1.
//
// Reference ops replaced: nn.Conv2d at vae_nets.py:74,79,84,117,121,125,129 (forward) and the data gradients of
// :74,79,84,117,121,125,129, selected by the caller with CVAE_KTAB_BLOCK64 and block-major packed weights
// (cvae_pack_weights with CVAE_PACK_KORDER_BLOCK64 | CVAE_PACK_STACKx).
#include "common.cuh"
#include "umma.cuh"
#include "tma.cuh"
#include "wa_groups.cuh"

namespace cvae {

static constexpr int kWaThreads = 352;      // 8 epilogue warps, pixel producer, weight producer, MMA issuer
static constexpr int kWaMaxStages = 8, kWaMaxSlots = 6;
static constexpr int kWaMarginLo = 8;        // zero rows in front of the loaded rows (left padding of the first row, J - 1 lead columns)
static constexpr int kWaCounters = 4096;     // split-K arrival counters at the head of the workspace
static constexpr int kWaPartialFloats = 128 * 256;

struct WaArgs {
    int B, H, W, pad, PW, IH;
    int nblk, kb;                 // source channel blocks, channels per block (64 or 32)
    int upb, units;               // tap groups per block, units of one K slice (= nblk / ksplit * upb)
    int ksplit, J;
    int m_blocks, c_total;        // GEMM M blocks (128 rows = 128 / J output channels each), channels of the output tensor
    int T0, T;                    // first virtual row with outputs, number of virtual rows to cover
    int nt;                       // tiles per CTA
    int slot_bytes, nslots;       // pixel-tile ring (one channel block per slot)
    int unit_bytes, ups, nstages, stage_bytes;  // weight ring: units per stage
    int phase_src;                // 1: block q comes from tensor map q (space-to-depth phases), channel 0
    int w_tma;                    // weight stages through a 2-D tensor map (one box per unit) instead of 1-D bulk copies
    int dbg_flags;                // experiments (CVAE_WA_DBG): 1 skip TMEM loads, 2 skip the transpose, 4 skip the stores
    int epilogue;
    int goff[kWaMaxGroups];       // B operand start of every tap group, in 16-byte descriptor units from the slot base:
                                  // (margin + pad * PW - (J - 1) + dy * PW + s) * row_bytes / 16
    const __nv_bfloat16* wpack;   // [m_blocks][nblk * upb units][kb / 16][128 x 16]
    void* out;
    const float* bias;
    const __nv_bfloat16* act;
    double* stats;
    float* ws;                    // split-K partials [item][ksplit][column 256][row 128]
    int* counters;
    int* fault;
    unsigned long long* dbg;
};

struct WaBars {
    uint64_t w_full[kWaMaxStages], w_empty[kWaMaxStages];
    uint64_t b_full[kWaMaxSlots], b_empty[kWaMaxSlots];
    uint64_t acc_full[2], acc_empty[2];
};

__device__ __forceinline__ unsigned long long wa_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
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

// Work of this CTA: tile group g of G = gridDim.x / ksplit, K slice ks.  Tile t of the group covers virtual rows
// [ra, ra + nrows): the T rows are dealt evenly to the G groups, every group cuts its share into nt near-equal tiles
// (a tile may be empty when there are fewer rows than tiles).
__device__ __forceinline__ void wa_tile(const WaArgs& a, int t, int& ra, int& nrows) {
    const long G = gridDim.x / a.ksplit, g = blockIdx.x % G;
    const int lo = (int)((long)a.T * g / G), hi = (int)((long)a.T * (g + 1) / G);
    const int n = hi - lo;
    const int b0 = (int)((long)n * t / a.nt), b1 = (int)((long)n * (t + 1) / a.nt);
    ra = a.T0 + lo + b0;
    nrows = b1 - b0;
}
__device__ __forceinline__ int wa_n(const WaArgs& a, int nrows) {
    int n = (nrows * a.PW + a.J - 1 + 15) & ~15;
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

// Tap-stacked accumulators: lane (channel = lane / J, j = lane % J) holds its row's columns [c, c + 16) in raw[];
// output pixel i of the group is the sum over j of column i + J - 1 - j.  Afterwards lane L < 32 / J holds channel L
// of the warp (lanes above repeat them and are masked out by the caller).
template <int J>
__device__ __forceinline__ void wa_unstack(const uint32_t (&raw)[16], float (&v)[8], int lane) {
    static_assert(J == 2 || J == 4, "tap stacking factor");
    const int jj = lane & (J - 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float x;
        if constexpr (J == 2) {
            x = jj == 0 ? __uint_as_float(raw[i + 1]) : __uint_as_float(raw[i]);
        } else {
            const float hi = (jj & 1) ? __uint_as_float(raw[i + 2]) : __uint_as_float(raw[i + 3]);   // jj = 1 : jj = 0
            const float lo = (jj & 1) ? __uint_as_float(raw[i]) : __uint_as_float(raw[i + 1]);       // jj = 3 : jj = 2
            x = (jj & 2) ? lo : hi;
        }
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        if constexpr (J == 4) x += __shfl_xor_sync(0xffffffffu, x, 2);
        v[i] = __shfl_sync(0xffffffffu, x, (lane & (32 / J - 1)) * J);
    }
}

// Epilogue of 8 consecutive output pixels (pix0 .. pix0 + 7, the first `nvalid` real) for the warp's channels
// [cbase, cbase + 32 / J): v[i] is this lane's channel at pixel pix0 + i (zero for i >= nvalid).
template <int EPI, int J>
__device__ __forceinline__ void wa_emit8(const WaArgs& a, float (&v)[8], int nvalid, int pix0, int cbase, int lane, float bias,
                                         float& t1, float& t2) {
    const bool lane_live = J == 1 || lane < 32 / J;
    if constexpr (EPI == CVAE_EPI_BIAS_RELU || EPI == CVAE_EPI_PHASE_BIAS_RELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i] + bias, 0.f);
    }
    uint32_t R[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) R[m] = pack_bf16x2(v[2 * m], v[2 * m + 1]);
    if constexpr (EPI == CVAE_EPI_STATS) {   // BatchNorm statistics of the stored (bf16-rounded) values; lane = channel
        if (lane_live) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const float lo = bf16_lo(R[m]), hi = bf16_hi(R[m]);
                t1 += lo + hi;
                t2 = fmaf(lo, lo, fmaf(hi, hi, t2));
            }
        }
    }
    if (!(a.dbg_flags & 2)) transpose8x8_bf16(R, lane);
    const int j = lane & 7, cg = cbase + (lane >> 3) * 8;
    if (j < nvalid && lane_live && !(a.dbg_flags & 4)) {
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

// One unit of epilogue work: 8 consecutive output pixels pix0 .. pix0 + 7 (the first `nvalid` real) whose accumulator
// columns start at `col` (W == 4: two half rows, the second at `col2`).
struct WaGroup {
    int col, col2, pix0, nvalid;
};

// Up to NB groups at once: all their accumulator loads are issued before the first one is consumed, and the NB
// independent shuffle transposes interleave -- one warp per scheduler is otherwise bound by the latency of
// tcgen05.ld, of the split-K partial loads and of the dependent shuffle stages, not by issue slots.
template <int EPI, int J, bool FROM_WS, int NB>
__device__ __forceinline__ void wa_batch(const WaArgs& a, const WaGroup (&g)[NB], int count, int cbase, uint32_t tbase, const float* part, int row,
                                         int lane, float bias, float& t1, float& t2) {
    constexpr int NC = J == 1 ? 8 : 16;        // tap-stacked rows need the J - 1 columns after the group as well
    uint32_t raw[NB][NC];
    if constexpr (!FROM_WS) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (a.dbg_flags & 1) {
#pragma unroll
                for (int i = 0; i < NC; ++i) raw[b][i] = 0x3f800000u;
            } else if (b < count) {
                if constexpr (J > 1) {
                    tmem_ld16(tbase + (uint32_t)g[b].col, raw[b]);
                } else if (g[b].col2 == -2) {          // W >= 8: eight columns of one row
                    tmem_ld8(tbase + (uint32_t)g[b].col, raw[b]);
                } else {                               // W == 4: two half rows (the second may be missing)
                    uint32_t lo[4], hi[4] = {0u, 0u, 0u, 0u};
                    tmem_ld4(tbase + (uint32_t)g[b].col, lo);
                    if (g[b].col2 >= 0) tmem_ld4(tbase + (uint32_t)g[b].col2, hi);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { raw[b][i] = lo[i]; raw[b][4 + i] = hi[i]; }
                }
            }
        }
        tmem_wait_ld();
    } else {
        // split-K, last CTA of the tile: the ksplit fp32 partials are added in slice order; every load of the batch
        // is in flight before the first add
        float p[NB][NC][4];
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                int c = g[b].col + i;
                if (J == 1 && g[b].col2 != -2 && i >= 4) c = g[b].col2 >= 0 ? g[b].col2 + i - 4 : -1;
                const bool live = b < count && c >= 0 && c < 256;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    p[b][i][k] = (live && k < a.ksplit) ? __ldcg(part + (size_t)k * kWaPartialFloats + (size_t)c * 128 + row) : 0.f;
            }
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < NC; ++i) raw[b][i] = __float_as_uint(((p[b][i][0] + p[b][i][1]) + p[b][i][2]) + p[b][i][3]);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (b < count) {
            float v[8];
            if constexpr (J == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[b][i]);
                if constexpr (EPI == CVAE_EPI_BIAS_RELU || EPI == CVAE_EPI_PHASE_BIAS_RELU) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (i >= g[b].nvalid) v[i] = -bias;   // stays zero through bias + ReLU (never stored anyway)
                }
            } else {
                wa_unstack<J>(raw[b], v, lane);
            }
            wa_emit8<EPI, J>(a, v, g[b].nvalid, g[b].pix0, cbase, lane, bias, t1, t2);
        }
    }
}

// Epilogue of one (tile, M block) by the 8 epilogue warps: walk the real rows, 8 output pixels per group, NB groups
// per batch; batches alternate between the two warps that share a TMEM lane quarter (half = warp / 4).
template <int EPI, int J, bool FROM_WS>
__device__ __forceinline__ void wa_epilogue_tile(const WaArgs& a, int ra, int nrows, int mb, uint32_t tbase, const float* part, int warp, int lane,
                                                 float& t1, float& t2) {
    constexpr int NB = 2;
    const int quarter = warp & 3, half = warp >> 2;
    const int cbase = mb * (128 / J) + quarter * (32 / J);
    const int row = quarter * 32 + lane;
    float bias = 0.f;
    if constexpr (EPI == CVAE_EPI_BIAS_RELU) bias = __ldg(a.bias + min(cbase + lane, a.c_total - 1));
    if constexpr (EPI == CVAE_EPI_PHASE_BIAS_RELU) bias = __ldg(a.bias + min(cbase + lane, a.c_total - 1) % (a.c_total >> 2));
    WaGroup g[NB];
    int count = 0, batch = 0;
    auto flush = [&]() {
        if (count > 0 && (batch & 1) == half) wa_batch<EPI, J, FROM_WS, NB>(a, g, count, cbase, tbase, part, row, lane, bias, t1, t2);
        if (count > 0) ++batch;
        count = 0;
    };
    int n = ra / a.IH, r = ra - n * a.IH;
    int pend_col = -1, pend_pix = 0;     // W == 4: the first half of a group waits for the next real row
    for (int tr = 0; tr < nrows; ++tr) {
        if (r >= a.pad && n < a.B) {
            const int pixbase = (n * a.H + r - a.pad) * a.W;
            const int col = tr * a.PW;           // J > 1: output pixel i of the row <- columns col + i + (J - 1 - j)
            if (a.W >= 8) {
                for (int c = 0; c < a.W; c += 8) {
                    g[count++] = WaGroup{col + c, -2, pixbase + c, 8};
                    if (count == NB) flush();
                }
            } else if (pend_col < 0) {
                pend_col = col;
                pend_pix = pixbase;
            } else {   // real rows are consecutive in the output: pend_pix + 4 == pixbase
                g[count++] = WaGroup{pend_col, col, pend_pix, 8};
                pend_col = -1;
                if (count == NB) flush();
            }
        }
        if (++r == a.IH) { r = 0; ++n; }
    }
    if (pend_col >= 0) g[count++] = WaGroup{pend_col, -1, pend_pix, 4};
    flush();
}

template <int EPI, int CL, int J>
__global__ void __launch_bounds__(kWaThreads, 1)
conv_wa_kernel(const WaArgs a, const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
               const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3, const __grid_constant__ CUtensorMap mapW) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ WaBars bars;
    __shared__ uint32_t tmem_slot;
    __shared__ int last_flag;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (a.dbg && tid == 0) a.dbg[(size_t)blockIdx.x * 16 + 8] = wa_now_ns();     // phase timestamps (profiling aid): kernel entry
    uint8_t* const bring = smem;                                       // [nslots][slot_bytes]
    uint8_t* const wring = smem + (size_t)a.nslots * a.slot_bytes;     // [nstages][stage_bytes]
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

    if (tid == 0) {
        for (int s = 0; s < a.nstages; ++s) { mbar_init(&bars.w_full[s], 1); mbar_init(&bars.w_empty[s], CL); }
        for (int s = 0; s < a.nslots; ++s) { mbar_init(&bars.b_full[s], 1); mbar_init(&bars.b_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bars.acc_full[b], 1); mbar_init(&bars.acc_empty[b], 8); }
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
    grid_dependency_sync();     // everything above touched only this CTA's shared memory and TMEM
    if (a.dbg && tid == 0) a.dbg[(size_t)blockIdx.x * 16 + 9] = wa_now_ns();     // prologue done
    const uint32_t tmem_base = tmem_slot;
    const int items = a.nt * a.m_blocks;
    const int G = gridDim.x / a.ksplit, ks = blockIdx.x / G;           // tile group count, this CTA's K slice
    const int nblk_s = a.nblk / a.ksplit, q0 = ks * nblk_s;            // channel blocks [q0, q0 + nblk_s)
    const int row_bytes = a.kb * 2;                                     // one pixel of one channel block

    if (warp < 8) {
        // ================================ epilogue ================================================
        // TMEM lane = GEMM row (this thread's), column = pixel of the tile.  Only the W real pixels of the real rows
        // are read back; they are consecutive in the NHWC output, so the tile leaves as 8-pixel groups.
        float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
        long long t_epi = 0, t_red = 0;
        for (int it = 0; it < items; ++it) {
            const int t = it / a.m_blocks, mb = it - t * a.m_blocks;
            int ra, nrows;
            wa_tile(a, t, ra, nrows);
            const uint32_t ab = (uint32_t)it & 1u;
            mbar_wait_relaxed(&bars.acc_full[ab], ((uint32_t)it >> 1) & 1u, a.fault);
            tc_fence_after();
            const long long te = clock64();
            const uint32_t tbase = tmem_base + ab * 256u + ((uint32_t)((warp & 3) * 32) << 16);
            float t1 = 0.f, t2 = 0.f;
            if (J > 1 || a.ksplit == 1) {      // (tap-stacked layers never split K: wa_plan)
                wa_epilogue_tile<EPI, J, false>(a, ra, nrows, mb, tbase, nullptr, warp, lane, t1, t2);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.acc_empty[ab]);
            } else if constexpr (J == 1) {
                // split-K: park the fp32 partial, count arrivals, the last CTA of the tile reduces and stores
                const int item_id = ((int)(blockIdx.x % G) * a.nt + t) * a.m_blocks + mb;
                float* part = a.ws + (size_t)item_id * a.ksplit * kWaPartialFloats;
                float* mine = part + (size_t)ks * kWaPartialFloats + (warp & 3) * 32 + lane;
                const int N = wa_n(a, nrows);
                for (int c0 = (warp >> 2) * 16; c0 < N; c0 += 32) {
                    uint32_t raw[16];
                    tmem_ld16(tbase + (uint32_t)c0, raw);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) __stcg(mine + (size_t)(c0 + i) * 128, __uint_as_float(raw[i]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.acc_empty[ab]);
                __threadfence();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (tid == 0) {
                    const int prev = atomicAdd(a.counters + item_id, 1);
                    last_flag = prev == a.ksplit - 1;
                    if (last_flag) a.counters[item_id] = 0;      // every slice has arrived: ready for the next launch
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (last_flag) {
                    __threadfence();
                    const long long tr0 = clock64();
                    wa_epilogue_tile<EPI, J, true>(a, ra, nrows, mb, 0u, part, warp, lane, t1, t2);
                    t_red += clock64() - tr0;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");      // last_flag is rewritten by the next item
            }
            s1[mb & 1] += t1;
            s2[mb & 1] += t2;
            t_epi += clock64() - te;
        }
        if (a.dbg && tid == 0) { a.dbg[(size_t)blockIdx.x * 16 + 6] = (unsigned long long)t_epi; a.dbg[(size_t)blockIdx.x * 16 + 7] = (unsigned long long)t_red; }
        if (a.dbg && lane == 0) atomicMax(a.dbg + (size_t)blockIdx.x * 16 + 12, wa_now_ns());    // last epilogue warp done
        if constexpr (EPI == CVAE_EPI_STATS) {
            // warps w and w + 4 hold the sums of the same channels (the two column halves): combine them in shared memory, one
            // set of double atomics per CTA (same-address atomics serialise in L2 and every CTA flushes at the same moment)
            // (the pixel ring at the start of the dynamic shared memory is dead by now: every tile of this CTA has been
            // through its MMAs and its epilogue; static shared memory has no room left beside the largest stage plans)
            float (*stat_comb)[2][2][32] = reinterpret_cast<float (*)[2][2][32]>(smem);      // [warp & 3][m block parity][sum | sum of squares][lane]
            asm volatile("bar.sync 2, 256;" ::: "memory");      // all eight epilogue warps are past their last tile
            if (warp >= 4) {
#pragma unroll
                for (int p = 0; p < 2; ++p) { stat_comb[warp & 3][p][0][lane] = s1[p]; stat_comb[warp & 3][p][1][lane] = s2[p]; }
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (warp < 4 && (J == 1 || lane < 32 / J)) {
                for (int mb = 0; mb < a.m_blocks; ++mb) {
                    const int co = mb * (128 / J) + (warp & 3) * (32 / J) + lane;
                    atomicAdd(a.stats + co, (double)(s1[mb & 1] + stat_comb[warp][mb & 1][0][lane]));
                    atomicAdd(a.stats + a.c_total + co, (double)(s2[mb & 1] + stat_comb[warp][mb & 1][1][lane]));
                }
            }
        }
    } else if (warp == 8) {
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
                for (int q = q0; q < q0 + nblk_s && alive; ++q, ++seq) {
                    const uint32_t slot = seq % (uint32_t)a.nslots;
                    alive = mbar_wait_relaxed(&bars.b_empty[slot], ((seq / (uint32_t)a.nslots) & 1u) ^ 1u, a.fault);
                    mbar_expect_tx(&bars.b_full[slot], (uint32_t)(nload * a.PW * row_bytes));
                    const CUtensorMap* mp = (!a.phase_src || q == 0) ? &map0 : (q == 1 ? &map1 : (q == 2 ? &map2 : &map3));
                    const int c0 = a.phase_src ? 0 : q * a.kb;
                    uint32_t dst = smem_u32(bring) + slot * (uint32_t)a.slot_bytes + (uint32_t)(kWaMarginLo * row_bytes);
                    const int vr = ra - a.pad;             // >= 0: outputs start at virtual row `pad`
                    int n = vr / a.IH, r = vr - n * a.IH;
                    for (int j = 0; j < nload; ++j) {
                        // a padding row (r < pad) or a row past the batch (n >= B) is entirely out of bounds: zero fill
                        tma_load_4d(dst, mp, c0, 0, r - a.pad, n, &bars.b_full[slot]);
                        dst += (uint32_t)(a.PW * row_bytes);
                        if (++r == a.IH) { r = 0; ++n; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ================================ weight producer ==========================================
        if (elect_one()) {
            const uint8_t* wbase = reinterpret_cast<const uint8_t*>(a.wpack);
            const int spi = (a.units + a.ups - 1) / a.ups;   // stages per item
            uint32_t s = 0;
            bool alive = true;
            for (int it = 0; it < items && alive; ++it) {
                const int mb = it % a.m_blocks;
                const uint8_t* src = wbase + ((size_t)mb * a.ksplit + ks) * a.units * a.unit_bytes;
                for (int st = 0; st < spi && alive; ++st, ++s) {
                    const uint32_t slot = s % (uint32_t)a.nstages;
                    alive = mbar_wait_relaxed(&bars.w_empty[slot], ((s / (uint32_t)a.nstages) & 1u) ^ 1u, a.fault);
                    const int nu = min(a.ups, a.units - st * a.ups);
                    const uint32_t bytes = (uint32_t)(nu * a.unit_bytes);
                    mbar_expect_tx(&bars.w_full[slot], bytes);
                    uint8_t* dst = wring + (size_t)slot * a.stage_bytes;
                    const uint8_t* from = src + (size_t)st * a.ups * a.unit_bytes;
                    if constexpr (CL == 1) {
                        if (a.w_tma) {   // the packed weights as a [rows][64] matrix of 128-byte rows: one dense box per unit
                            const int rpu = a.unit_bytes >> 7;
                            const int row0 = (int)((from - wbase) >> 7);
                            for (int k = 0; k < nu; ++k)
                                tma_load_2d(smem_u32(dst) + (uint32_t)(k * a.unit_bytes), &mapW, 0, row0 + k * rpu, &bars.w_full[slot]);
                        } else {
                            bulk_g2s(dst, from, bytes, &bars.w_full[slot]);
                        }
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
            // pixels: 8-row atoms (SBO = 8 rows), SWIZZLE_128B (layout type 2) for 128-byte rows, SWIZZLE_64B (4) for 64-byte rows
            const uint32_t b_hi = ((uint32_t)(8 * row_bytes) >> 4) | (1u << 14) | ((a.kb == 64 ? 2u : 4u) << 29);
            const uint32_t b_lbo = 1u << 16;                                  // ignored for swizzled K-major
            const uint32_t bring16 = (smem_u32(bring) & 0x3FFFFu) >> 4, wring16 = (smem_u32(wring) & 0x3FFFFu) >> 4;
            const uint32_t unit16 = (uint32_t)a.unit_bytes >> 4, stage16 = (uint32_t)a.stage_bytes >> 4;
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
                for (int q = 0; q < nblk_s && alive; ++q, ++seq) {
                    const uint32_t slot = seq % (uint32_t)a.nslots;
                    if (prof) tq = clock64();
                    alive = mbar_wait(&bars.b_full[slot], (seq / (uint32_t)a.nslots) & 1u, a.fault);
                    if (prof) t_b += clock64() - tq;
                    tc_fence_after();
                    const uint32_t b_slot16 = bring16 + ((slot * (uint32_t)a.slot_bytes) >> 4);
                    int goff_next = a.goff[0];
                    for (int g = 0; g < a.upb && alive; ++g, ++u) {
                        const int goff = goff_next;                          // the table lives in constant memory: fetch the next
                        goff_next = a.goff[g + 1 < a.upb ? g + 1 : 0];       // group's offset behind this group's MMAs
                        if (uin == 0) {
                            wslot = s % (uint32_t)a.nstages;
                            if (prof) tq = clock64();
                            alive = mbar_wait(&bars.w_full[wslot], (s / (uint32_t)a.nstages) & 1u, a.fault);
                            if (prof) t_w += clock64() - tq;
                            tc_fence_after();
                        }
                        const uint32_t a_lo = (wring16 + wslot * stage16 + (uint32_t)uin * unit16) | a_lbo;
                        const uint32_t b_lo = (b_slot16 + (uint32_t)goff) | b_lbo;
                        if (prof && it == 0 && u == 0) a.dbg[(size_t)blockIdx.x * 16 + 10] = wa_now_ns();   // first MMA
                        if (a.kb == 64) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                umma_bf16(acc, ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)(j * 256)),
                                          ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)(j * 2)), idesc, j == 0 ? accumulate : 1u);
                        } else {
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                umma_bf16(acc, ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)(j * 256)),
                                          ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)(j * 2)), idesc, j == 0 ? accumulate : 1u);
                        }
                        accumulate = 1u;
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
                unsigned long long* o = a.dbg + (size_t)blockIdx.x * 16;
                unsigned long long g1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
                o[0] = (unsigned long long)(clock64() - t0); o[1] = t_acc; o[2] = t_b; o[3] = t_w; o[4] = items; o[5] = g1 - g0;
                o[11] = g1;                                                                         // last MMA issued
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    if constexpr (CL > 1) cluster_sync_all(); else __syncthreads();
    if (warp == 0) tmem_free(tmem_base, 512);
    if (a.dbg && tid == 0) a.dbg[(size_t)blockIdx.x * 16 + 13] = wa_now_ns();    // kernel exit
}

static const bool g_wa_debug = getenv("CVAE_DEBUG") != nullptr;
// tuning overrides (0 = automatic), set through cvae_conv_wa_tune (tests sweep them, tools/conv_bench.py explores them)
static int g_wa_cluster = 0, g_wa_grid = 0, g_wa_ups = 0, g_wa_nt = 0, g_wa_ksplit = 0, g_wa_wload = 0;
unsigned long long* g_wa_dbg = nullptr;

template <int EPI, int CL, int J>
static int launch_wa(const WaArgs& a, const CUtensorMap* maps, int grid, size_t smem, cudaStream_t stream) {
    CVAE_OPT_IN_SMEM((conv_wa_kernel<EPI, CL, J>), smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kWaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CL > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CL; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {      // programmatic dependent launch (common.cuh: launch / grid_dependency_sync)
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    CVAE_CUDA(cudaLaunchKernelEx(&cfg, conv_wa_kernel<EPI, CL, J>, a, maps[0], maps[1], maps[2], maps[3], maps[4]));
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

// Everything the launch needs that follows from the descriptor alone (also what the workspace query reports).
struct WaPlan {
    WaArgs a;
    int grid, cl;
    size_t smem;
    int64_t ws_bytes;
};

static int wa_plan(const cvae_conv_desc* d, WaPlan& p) {
    const int J = d->stack > 0 ? d->stack : 1;
    CVAE_REQUIRE(J == 1 || J == 2 || J == 4, CVAE_EINVAL, "conv_wa: stack %d", d->stack);
    CVAE_REQUIRE(d->src_channels % 32 == 0 && d->src_channels <= 256, CVAE_EINVAL, "conv_wa: src_channels %d", d->src_channels);
    CVAE_REQUIRE((d->n_total * J) % 128 == 0 && d->n_total * J <= 256, CVAE_EINVAL, "conv_wa: n_total %d with stack %d", d->n_total, J);
    CVAE_REQUIRE(d->loader == CVAE_LOAD_NHWC || d->loader == CVAE_LOAD_S2D, CVAE_EINVAL, "conv_wa: loader %d", d->loader);
    CVAE_REQUIRE(d->epilogue == CVAE_EPI_STATS || d->epilogue == CVAE_EPI_BIAS_RELU || d->epilogue == CVAE_EPI_PHASE_BIAS_RELU ||
                     d->epilogue == CVAE_EPI_MASK || d->epilogue == CVAE_EPI_PLAIN, CVAE_EINVAL, "conv_wa: epilogue %d", d->epilogue);
    CVAE_REQUIRE(d->width == 8 || d->width % 16 == 0 || (d->width == 4 && J == 1), CVAE_EINVAL, "conv_wa: width %d (4, 8 or a multiple of 16)", d->width);
    CVAE_REQUIRE(d->epilogue != CVAE_EPI_PHASE_BIAS_RELU || ((d->n_total / 4) % 8 == 0 && J == 1), CVAE_EINVAL,
                 "conv_wa: phase epilogue needs n_total / 4 in multiples of 8 and no tap stacking");

    WaArgs& a = p.a;
    a = WaArgs{};
    a.B = d->batch; a.H = d->height; a.W = d->width; a.pad = d->ksize / 2;
    a.PW = a.W + a.pad; a.IH = a.H + a.pad;
    // TMA writes need 128-byte aligned shared-memory destinations: with 64-byte pixel rows (32-channel blocks) the
    // virtual row pitch must be even -- one more out-of-bounds (zero) column per row where W + pad is odd
    const bool narrow = d->loader == CVAE_LOAD_S2D ? d->src_channels / 4 == 32 : d->src_channels % 64 != 0;
    if (narrow && (a.PW & 1)) a.PW += 1;
    CVAE_REQUIRE(a.PW <= 120 && (long)a.B * a.IH * a.PW < (1L << 30), CVAE_EINVAL, "conv_wa: map too large");
    a.J = J;
    // channel blocks: 64 channels (128-byte rows) where the source allows, else 32; a space-to-depth source has one block per phase
    a.phase_src = d->loader == CVAE_LOAD_S2D ? 1 : 0;
    if (a.phase_src) {
        a.kb = d->src_channels / 4;
        CVAE_REQUIRE(a.kb == 64 || a.kb == 32, CVAE_EINVAL, "conv_wa: space-to-depth source with %d channels per phase", a.kb);
        a.nblk = 4;
    } else {
        a.kb = d->src_channels % 64 == 0 ? 64 : 32;
        a.nblk = d->src_channels / a.kb;
    }
    a.unit_bytes = 128 * a.kb * 2;
    a.upb = wa_group_count(d->ksize, J);
    CVAE_REQUIRE(a.upb <= kWaMaxGroups, CVAE_EINVAL, "conv_wa: too many tap groups");
    for (int g = 0; g < a.upb; ++g) {
        int dy, s, lo;
        wa_group(d->ksize, J, g, dy, s, lo);
        a.goff[g] = (kWaMarginLo + a.pad * a.PW - (J - 1) + dy * a.PW + s) * (a.kb * 2 / 16);
    }
    a.m_blocks = d->n_total * J / 128;
    a.c_total = d->n_total;
    a.T0 = a.pad;
    a.T = a.B * a.IH - a.pad;
    a.epilogue = d->epilogue;
    a.wpack = (const __nv_bfloat16*)d->wpack; a.out = d->out; a.bias = d->bias; a.act = (const __nv_bfloat16*)d->act; a.stats = d->stats;
    a.dbg = g_wa_dbg;

    // ---- work decomposition -----------------------------------------------------------------------------------
    // Grid size, K split and cluster size from a small cost model: per CTA, MMA time = MMAs x max(N/2, 32 + N/4)
    // cycles against the time to pull the weight stream out of L2 (every tile needs the whole weight matrix of its
    // M blocks once, whichever CTAs share it), plus per-tile pipeline costs and the split-K round trip.
    const int sms = sm_count();
    const int rows_max = (256 - (J > 1 ? 8 + J - 1 : 0)) / a.PW;   // rows per tile so that N (and the epilogue's 16-column windows) stay <= 256
    CVAE_REQUIRE(rows_max >= 1, CVAE_EINVAL, "conv_wa: map too wide");
    // Full grids win (tools/conv_bench.py --wa-sweep, profiles/r02_wa_sweep.log): one CTA per SM, each CTA's share of the
    // virtual rows cut into tiles of <= 256 pixels.  Where a full grid would leave a CTA fewer than ~96 pixels (the 4 x 4
    // maps: fewer 256-pixel tiles than SMs, and every tile streams the whole weight matrix) the channel blocks are split
    // over ksplit CTAs per tile instead.  Clusters with multicast weight stages never paid on this chip (measured, and
    // B300_MICROARCH.md: multicast saves L2 traffic only from cluster size 8 up): opt-in through cvae_conv_wa_tune.
    int grid = 0, cl = g_wa_cluster > 0 ? g_wa_cluster : 1, ksplit = 1;
    {
        const long px = (long)a.T * a.PW;
        if (J == 1 && px / sms < 96) {
            for (int ksp : {4, 2})
                if (a.nblk % ksp == 0 && px / (sms / ksp) <= 256 && ksplit == 1) ksplit = ksp;
        }
        if (J == 1 && g_wa_ksplit > 0) {                        // a forced split that does not divide the blocks is halved until it does
            ksplit = g_wa_ksplit;
            while (ksplit > 1 && a.nblk % ksplit) ksplit >>= 1;
        }
        if (ksplit > 1 || J > 1) cl = 1;
        if (cl != 1 && cl != 2 && cl != 4) cl = 1;
        int usable = cl == 4 ? sms - sms % 4 - 16 : sms;         // clusters of 4 strand ~16 SMs (GPC shapes)
        if (g_wa_grid > 0 && g_wa_grid < usable) usable = g_wa_grid;
        int G = usable / ksplit;
        const long want = (px + 47) / 48;                        // tiny batches: keep >= ~48 pixels per CTA
        if (g_wa_grid == 0 && G > want) G = (int)(want < 1 ? 1 : want);
        G -= G % cl;
        if (G < cl) { cl = 1; if (G < 1) G = 1; }
        grid = G * ksplit;
    }
    a.ksplit = ksplit;
    a.units = a.nblk / ksplit * a.upb;
    const int G = grid / ksplit;
    const int rows_cta = (a.T + G - 1) / G;              // largest share
    a.nt = (rows_cta + rows_max - 1) / rows_max;
    if (g_wa_nt > a.nt) a.nt = g_wa_nt;
    const int tile_rows = (rows_cta + a.nt - 1) / a.nt;        // largest tile
    // pixel ring: margin, loaded rows, tail margin (N rounding + pad + J - 1 + slack), 1024-byte aligned slots
    const int row_bytes = a.kb * 2;
    const int slot_rows = kWaMarginLo + (tile_rows + 2 * a.pad) * a.PW + 16 + a.pad + J + 8;
    a.slot_bytes = (slot_rows * row_bytes + 1023) & ~1023;
    const size_t cap = 224 * 1024;                       // 227 KB per CTA minus the static shared memory (barriers)
    const int nblk_s = a.nblk / ksplit;
    a.nslots = 2;
    // Weight stages as large as still leaves three of them in flight: the MMA thread pays ~900 cycles per stage for its
    // barrier wait, fence and commit whatever the stage holds (tools/umma_rate.cu wa: 8 MMAs per stage run at
    // max(900, MMA time + 220) cycles), so an N = 176 tile needs >= 12 MMAs per stage to stay bound by the tensor pipe.
    const int per16k = 16384 / a.unit_bytes;              // units per 16 KB (1 for 64-channel blocks, 2 for 32-channel blocks)
    for (int u16 : {4, 3, 2, 1}) {
        a.ups = g_wa_ups > 0 ? g_wa_ups : u16 * per16k;
        a.stage_bytes = a.ups * a.unit_bytes;
        const long left = (long)cap - (long)a.nslots * a.slot_bytes;
        a.nstages = left > 0 ? (int)(left / a.stage_bytes) : 0;
        if (a.nstages >= 3 || g_wa_ups > 0) break;
    }
    if (a.nstages > kWaMaxStages) a.nstages = kWaMaxStages;
    CVAE_REQUIRE(a.nstages >= 2, CVAE_EINVAL, "conv_wa: shape does not fit shared memory");
    // spend what is left on more pixel slots (deeper prefetch across blocks and tiles)
    while (a.nslots < kWaMaxSlots && a.nslots < 2 * nblk_s &&
           (size_t)(a.nslots + 1) * a.slot_bytes + (size_t)a.nstages * a.stage_bytes <= cap) ++a.nslots;
    p.smem = (size_t)a.nslots * a.slot_bytes + (size_t)a.nstages * a.stage_bytes;
    CVAE_REQUIRE(p.smem < (1u << 18), CVAE_EINVAL, "conv_wa: descriptor range");
    p.grid = grid;
    p.cl = cl;
    const int64_t items_total = (int64_t)G * a.nt * a.m_blocks;
    CVAE_REQUIRE(ksplit == 1 || items_total <= kWaCounters, CVAE_EINVAL, "conv_wa: too many split-K tiles");
    p.ws_bytes = ksplit == 1 ? 0 : (int64_t)kWaCounters * 4 + items_total * ksplit * kWaPartialFloats * 4;
    return CVAE_OK;
}

int64_t conv_wa_workspace_bytes(const cvae_conv_desc* d) {
    WaPlan p;
    if (wa_plan(d, p) != CVAE_OK) return -1;
    return p.ws_bytes;
}

// Called by cvae_conv_gemm for descriptors with ktab == CVAE_KTAB_BLOCK64.
int conv_wa_dispatch(const cvae_conv_desc* d, cudaStream_t stream) {
    CVAE_REQUIRE(d->epilogue != CVAE_EPI_STATS || d->stats, CVAE_EINVAL, "conv_wa: statistics buffer missing");
    CVAE_REQUIRE((d->epilogue != CVAE_EPI_BIAS_RELU && d->epilogue != CVAE_EPI_PHASE_BIAS_RELU) || d->bias, CVAE_EINVAL, "conv_wa: bias missing");
    CVAE_REQUIRE(d->epilogue != CVAE_EPI_MASK || d->act, CVAE_EINVAL, "conv_wa: activation (mask) tensor missing");
    CVAE_REQUIRE(tensor_map_encoder() != nullptr, CVAE_ECUDA, "conv_wa: cuTensorMapEncodeTiled unavailable");
    WaPlan p;
    const int rc = wa_plan(d, p);
    if (rc != CVAE_OK) return rc;
    WaArgs& a = p.a;
    a.fault = fault_flag();
    CVAE_REQUIRE(a.fault != nullptr, CVAE_ECUDA, "conv_wa: fault flag unavailable");
    if (a.ksplit > 1) {
        CVAE_REQUIRE(d->workspace && d->workspace_bytes >= p.ws_bytes, CVAE_EINVAL,
                     "conv_wa: split-K needs a zero-initialised workspace of %lld bytes (cvae_conv_gemm_workspace_bytes)", (long long)p.ws_bytes);
        a.counters = (int*)d->workspace;
        a.ws = (float*)((char*)d->workspace + (size_t)kWaCounters * 4);
    }

    // ---- tensor maps: one row box {kb ch, PW, 1, 1} per call ----------------------------------------------------
    CUtensorMap maps[5];
    bool ok = true;
    const CUtensorMapSwizzle swz = a.kb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    if (!a.phase_src) {
        const long C = d->src_channels;
        ok = encode_map_4d(&maps[0], d->src, (int)C, a.W, a.H, a.B, C, (long)a.W * C, (long)a.H * a.W * C, a.kb, a.PW, 1, 1, swz);
        maps[1] = maps[2] = maps[3] = maps[0];
    } else {   // source [B][2H][2W][kb]: one strided view per phase (a, b) = block q
        const long C = a.kb;
        for (int ab = 0; ab < 4 && ok; ++ab) {
            const __nv_bfloat16* b0 = (const __nv_bfloat16*)d->src + ((long)(ab >> 1) * 2 * a.W + (ab & 1)) * C;
            ok = encode_map_4d(&maps[ab], b0, (int)C, a.W, a.H, a.B, 2 * C, 2L * 2 * a.W * C, 4L * a.H * a.W * C, a.kb, a.PW, 1, 1, swz);
        }
    }
    a.w_tma = g_wa_wload == 2 ? 1 : 0;
    static const int dbg_flags = getenv("CVAE_WA_DBG") ? atoi(getenv("CVAE_WA_DBG")) : 0;
    a.dbg_flags = dbg_flags;
    ok = ok && encode_map_2d(&maps[4], d->wpack, 64, (long)a.m_blocks * a.nblk * a.upb * (a.unit_bytes >> 7), 64, a.unit_bytes >> 7);
    CVAE_REQUIRE(ok, CVAE_ECUDA, "conv_wa: cuTensorMapEncodeTiled failed");
    if (g_wa_debug)
        fprintf(stderr, "conv_wa E%d B=%d %dx%d k%d C%d->N%d J=%d kb=%d: grid=%d cluster=%d ksplit=%d nt=%d units=%d ups=%d stages=%d slots=%d x %d B smem=%zu ws=%lld\n",
                d->epilogue, a.B, a.H, a.W, d->ksize, d->src_channels, d->n_total, a.J, a.kb, p.grid, p.cl, a.ksplit, a.nt, a.units, a.ups,
                a.nstages, a.nslots, a.slot_bytes, p.smem, (long long)p.ws_bytes);
#define CVAE_WA_CASE(E_)                                                                    \
    if (d->epilogue == (E_)) {                                                              \
        if (a.J == 4) return launch_wa<E_, 1, 4>(a, maps, p.grid, p.smem, stream);          \
        if (a.J == 2) return launch_wa<E_, 1, 2>(a, maps, p.grid, p.smem, stream);          \
        if (p.cl == 4) return launch_wa<E_, 4, 1>(a, maps, p.grid, p.smem, stream);         \
        if (p.cl == 2) return launch_wa<E_, 2, 1>(a, maps, p.grid, p.smem, stream);         \
        return launch_wa<E_, 1, 1>(a, maps, p.grid, p.smem, stream);                        \
    }
    CVAE_WA_CASE(CVAE_EPI_STATS)
    CVAE_WA_CASE(CVAE_EPI_MASK)
    CVAE_WA_CASE(CVAE_EPI_PLAIN)
#undef CVAE_WA_CASE
#define CVAE_WA_CASE1(E_)                                                                   \
    if (d->epilogue == (E_)) {                                                              \
        CVAE_REQUIRE(a.J == 1, CVAE_EINVAL, "conv_wa: epilogue %d has no tap-stacked variant", (E_)); \
        if (p.cl == 4) return launch_wa<E_, 4, 1>(a, maps, p.grid, p.smem, stream);         \
        if (p.cl == 2) return launch_wa<E_, 2, 1>(a, maps, p.grid, p.smem, stream);         \
        return launch_wa<E_, 1, 1>(a, maps, p.grid, p.smem, stream);                        \
    }
    CVAE_WA_CASE1(CVAE_EPI_BIAS_RELU)
    CVAE_WA_CASE1(CVAE_EPI_PHASE_BIAS_RELU)
#undef CVAE_WA_CASE1
    return CVAE_EINVAL;
}

}  // namespace cvae

// Profiling aid: per-CTA cycle counters of the MMA thread (total, wait accumulator / pixels / weights, items, ns).
extern "C" void cvae_conv_wa_debug_counters(void* device_buf) { cvae::g_wa_dbg = (unsigned long long*)device_buf; }
// Tuning / test hook: force the cluster size (1, 2, 4), the grid size, the units per weight stage, a minimum number
// of tiles per CTA, the K split and how weight stages are fetched (1: 1-D bulk copies, 2: 2-D tensor-map boxes) of the
// weights-as-A kernel; 0 restores the automatic choice.  Process-wide.
extern "C" void cvae_conv_wa_tune(int cluster, int grid, int units_per_stage, int tiles_per_cta, int ksplit, int weight_load) {
    cvae::g_wa_cluster = cluster; cvae::g_wa_grid = grid; cvae::g_wa_ups = units_per_stage; cvae::g_wa_nt = tiles_per_cta;
    cvae::g_wa_ksplit = ksplit; cvae::g_wa_wload = weight_load;
}
