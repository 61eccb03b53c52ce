// Implicit-GEMM 5x5 / 3x3 convolution on tcgen05 for 64x64-and-smaller feature maps.
//
// Layout idea ("halo planes"): the CTA's slice of the input is staged in shared memory as planes
// [channel/8][virtual pixel][8 channels], i.e. in the no-swizzle UMMA core-matrix layout with
// pixels as rows.  Images are laid out in a linear "virtual pixel" space with `pad` shared zero
// columns per row and `pad` shared zero rows per image, so the A operand of every filter tap is the
// SAME shared-memory tile read through a descriptor whose start address is shifted by
// (dy*PW + dx) pixels.  No im2col, no re-fetch per tap: 25x (9x) reuse out of shared memory.
//
// Pipeline (conv_pipe_kernel, one persistent CTA per SM, 10 warps):
//   warps 4-7  plane producers: cp.async 16-byte copies global -> plane buffer (zero-fill for the
//              padding), two buffers, one (pixel chunk, channel group) each; the two fp32 NCHW sources
//              (frames for encoder conv 0, d_recon * tanh' for decoder conv 4's data gradient) are
//              converted to bf16 by the producer threads themselves
//   warp  8    weight producer: cp.async.bulk of packed K-step blocks into a deep mbarrier ring
//              (stages as large as fit: the MMA thread pays a fixed cost per stage)
//   warp  9    one elected thread issues tcgen05.mma: `tm` 128-pixel tiles share every weight stage,
//              their accumulators sit side by side in TMEM; two accumulator sets alternate
//   warps 0-3  epilogue: tcgen05.ld -> bias / activation / BatchNorm statistics / ReLU mask ->
//              global, overlapped with the MMAs of the next work item
// Work item = (pixel chunk of tm*128 virtual pixels, N block); channel groups of <= 128 channels
// are accumulated into the same TMEM tile so the plane buffers stay <= ~80 KB each.
// Encoder conv 0 (3 -> 8 padded channels) uses a paired-tap K order: 13 K = 16 steps cover the 25 taps.
//
// Reference ops replaced: nn.Conv2d(5,1,2) at vae_nets.py:69,74,79,84,117,121,125,129,133, the
// nn.Upsample(2) at :119,123,127,131 (folded: conv5x5(up2(x)) == depth_to_space(conv3x3_4C(x)))
// and their autograd data-gradients.
#include "common.cuh"
#include "umma.cuh"
#include "planes.cuh"

namespace cvae {

int conv_wa_dispatch(const cvae_conv_desc* d, cudaStream_t stream);   // conv_wa.cu
int64_t conv_wa_workspace_bytes(const cvae_conv_desc* d);

struct ConvArgs {
    int B, H, W, pad, KW;
    int PW, IH;          // W + pad, H + pad
    int planes;          // 8-channel planes of one channel group (= all planes in the sequential kernel)
    int ncg;             // channel groups
    int n_blocks;        // N blocks of the kernel (work items per chunk)
    int c_total;         // n_total
    int nb_pack;         // N block of the packed weights (min(n_total, 128))
    int ksteps;          // K=16 steps over all channel groups
    int kpg;             // K steps per channel group
    int ksps;            // K steps per weight stage
    int tm;              // tiles per pass
    int num_chunks;
    int halo;            // pad*PW + pad
    int L;               // pixels per plane (tm*128 + 2*halo + 8)
    int plane_stride;    // bytes
    int buf_bytes;       // bytes of one plane buffer (pipelined kernel)
    int ktab_mode;
    int nstages;
    int resident;        // whole weight matrix lives in the ring (loaded once)
    int kgroup;          // K steps issued per unrolled group (divides planes/2 and ksps)
    int rotate;          // CTAs start the K loop at different weight stages (spreads the L2 reads of the shared weights)
    FastDiv dPW, dIH;
    PlaneSrc ps;         // where the A-operand planes come from
    const __nv_bfloat16* wpack;
    void* out;
    const float* bias;
    const __nv_bfloat16* act;
    double* stats;
    int* fault;
    int dbg_flags;            // experiments: 1 = skip epilogue work, 2 = fill planes only once per buffer
    unsigned long long* dbg;  // optional per-CTA cycle counters [grid][8] (cvae_conv_debug_counters)
};

// --------------------------------------------------------------------------------------------
// epilogue of one 128-pixel tile: warp w owns TMEM lanes [32w, 32w+32) = rows of the tile
// --------------------------------------------------------------------------------------------
template <int EPI, int N>
__device__ __forceinline__ void epilogue_tile(const ConvArgs& a, uint32_t tmem_tile, int v_tile, int nb, int warp, int lane,
                                              float* s1, float* s2, float* stat_scratch) {
    const int v = v_tile + warp * 32 + lane;
    int vrow = v / a.PW;
    int vcol = v - vrow * a.PW;
    int n = vrow / a.IH;
    int r = vrow - n * a.IH;
    const bool valid = (vcol < a.W) && (r >= a.pad) && (n < a.B);
    const int h = r - a.pad, w = vcol;
    const size_t pix = ((size_t)n * a.H + h) * a.W + w;
    if constexpr (EPI == CVAE_EPI_STATS && N == 32) {
        // 32-channel layers (encoder conv 0): one thread owns one pixel's whole channel row, so the BatchNorm sums stay
        // thread-private (s1 / s2 hold one entry per CHANNEL here) until flush_stats reduces them across the lanes once
        // per kernel -- no shared-memory transpose per tile.  The sums are those of the bf16 values actually stored.
        uint32_t r0[16], r1[16];
        tmem_ld16(tmem_tile + ((uint32_t)(warp * 32) << 16), r0);
        tmem_ld16(tmem_tile + ((uint32_t)(warp * 32) << 16) + 16u, r1);
        tmem_wait_ld();
        if (valid) {
            uint32_t p[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                p[i] = i < 8 ? pack_bf16x2(__uint_as_float(r0[2 * i]), __uint_as_float(r0[2 * i + 1]))
                             : pack_bf16x2(__uint_as_float(r1[2 * i - 16]), __uint_as_float(r1[2 * i - 15]));
                const float lo = bf16_lo(p[i]), hi = bf16_hi(p[i]);
                s1[2 * i] += lo;
                s1[2 * i + 1] += hi;
                s2[2 * i] = fmaf(lo, lo, s2[2 * i]);
                s2[2 * i + 1] = fmaf(hi, hi, s2[2 * i + 1]);
            }
            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + pix * a.c_total + nb * N);
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
        }
        return;
    }
#pragma unroll
    for (int g = 0; g < N / 16; ++g) {
        uint32_t raw[16];
        tmem_ld16(tmem_tile + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * 16), raw);
        tmem_wait_ld();
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(raw[i]);
        const int c0 = nb * N + g * 16;  // first global output column of this group

        if constexpr (EPI == CVAE_EPI_STATS) {
            uint32_t p[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
            float* sc = stat_scratch + warp * 32 * 17;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                sc[lane * 17 + 2 * i] = valid ? bf16_lo(p[i]) : 0.f;
                sc[lane * 17 + 2 * i + 1] = valid ? bf16_hi(p[i]) : 0.f;
            }
            __syncwarp();
            const int col = lane & 15, half = lane >> 4;
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) {
                const float x = sc[(half * 16 + rr) * 17 + col];
                sa += x;
                sb += x * x;
            }
            sa += __shfl_xor_sync(0xffffffffu, sa, 16);
            sb += __shfl_xor_sync(0xffffffffu, sb, 16);
            s1[g] += sa;
            s2[g] += sb;
            __syncwarp();
            if (valid) {
                uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + pix * a.c_total + c0);
                o[0] = make_uint4(p[0], p[1], p[2], p[3]);
                o[1] = make_uint4(p[4], p[5], p[6], p[7]);
            }
        } else if constexpr (EPI == CVAE_EPI_BIAS_RELU || EPI == CVAE_EPI_PLAIN || EPI == CVAE_EPI_MASK) {
            if (valid) {
                if constexpr (EPI == CVAE_EPI_BIAS_RELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i] + __ldg(a.bias + c0 + i), 0.f);
                }
                if constexpr (EPI == CVAE_EPI_MASK) {
                    const uint4* m = reinterpret_cast<const uint4*>(a.act + pix * a.c_total + c0);
                    const uint4 m0 = __ldg(m), m1 = __ldg(m + 1);
                    const uint32_t mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (!(bf16_lo(mm[i]) > 0.f)) f[2 * i] = 0.f;
                        if (!(bf16_hi(mm[i]) > 0.f)) f[2 * i + 1] = 0.f;
                    }
                }
                uint32_t p[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
                uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + pix * a.c_total + c0);
                o[0] = make_uint4(p[0], p[1], p[2], p[3]);
                o[1] = make_uint4(p[4], p[5], p[6], p[7]);
            }
        } else if constexpr (EPI == CVAE_EPI_PHASE_BIAS_RELU) {
            if (valid) {
                const int cout = a.c_total >> 2;
                const int ab = c0 / cout, co = c0 - ab * cout;
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i] + __ldg(a.bias + co + i), 0.f);
                uint32_t p[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
                const size_t opix = ((size_t)n * 2 * a.H + 2 * h + (ab >> 1)) * (2 * a.W) + 2 * w + (ab & 1);
                uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + opix * cout + co);
                o[0] = make_uint4(p[0], p[1], p[2], p[3]);
                o[1] = make_uint4(p[4], p[5], p[6], p[7]);
            }
        } else {  // CVAE_EPI_PHASE_BIAS_TANH: 12 of the 16 columns are (a,b,c)
            if (valid) {
                float* o = reinterpret_cast<float*>(a.out);
                const int H2 = 2 * a.H, W2 = 2 * a.W;
#pragma unroll
                for (int ab = 0; ab < 4; ++ab)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const size_t idx = (((size_t)n * 3 + c) * H2 + 2 * h + (ab >> 1)) * W2 + 2 * w + (ab & 1);
                        o[idx] = tanhf(f[ab * 3 + c] + __ldg(a.bias + c));
                    }
            }
        }
    }
}

// The four epilogue warps combine their sums in shared memory (`comb`: the per-warp transposition scratch, [4 warps][32 * 17]
// floats, free between tiles; a warp parks 2 N <= 256 sums in its own region) and warp 0 alone adds them to the
// global statistics: same-address double atomics serialise in L2 (~20 cycles each), and all CTAs flush at the same moment at
// the end of the kernel -- 148 per address instead of 592.  Every epilogue warp calls this at the same items (named barrier 2).
template <int EPI, int N>
__device__ __forceinline__ void flush_stats(const ConvArgs& a, int nb, int warp, int lane, float* s1, float* s2, float* comb) {
    if constexpr (EPI == CVAE_EPI_STATS) {
        constexpr int kStride = 32 * 17;
        static_assert(2 * N <= kStride, "statistics do not fit the per-warp scratch");
        float* mine = comb + warp * kStride;
        if constexpr (N == 32) {      // thread-private per-channel sums (see epilogue_tile)
            float m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float t1 = warp_sum(s1[c]), t2 = warp_sum(s2[c]);
                if (lane == c) { m1 = t1; m2 = t2; }
                s1[c] = s2[c] = 0.f;
            }
            mine[lane] = m1;
            mine[N + lane] = m2;
        } else {
            if (lane < 16) {
#pragma unroll
                for (int g = 0; g < N / 16; ++g) {
                    mine[g * 16 + lane] = s1[g];
                    mine[N + g * 16 + lane] = s2[g];
                }
            }
#pragma unroll
            for (int g = 0; g < N / 16; ++g) s1[g] = s2[g] = 0.f;
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (warp == 0) {
            for (int c = lane; c < 2 * N; c += 32) {
                const float t = (comb[c] + comb[kStride + c]) + (comb[2 * kStride + c] + comb[3 * kStride + c]);
                const int which = c / N, ch = c - which * N;
                atomicAdd(a.stats + which * a.c_total + nb * N + ch, (double)t);
            }
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");      // comb may be rewritten by the next flush
    }
}

// Barriers of the pipelined kernel (static shared memory).
static constexpr int kMaxStages = 32;
struct PipeBars {
    uint64_t w_full[kMaxStages], w_empty[kMaxStages];
    uint64_t p_full[2], p_empty[2], acc_full[2], acc_empty[2];
};

// KS consecutive K steps (same tap, consecutive channel pairs, same ring stage) x TM tiles, fully unrolled:
// every descriptor is one of three bases plus an immediate (or plus j * cp16, a kernel parameter), so the
// compiler needs one register -> uniform-register move per base per call instead of several per MMA.
template <int N, int TM, int KS>
__device__ __forceinline__ void issue_fixed(uint32_t acc, uint32_t a_lo0, uint32_t b_lo0, uint32_t cp16, uint32_t idesc,
                                            uint32_t accumulate_first) {
    const uint32_t a_hi = (128u >> 4) | (1u << 14);  // SBO 128 B, descriptor version 1
    const uint32_t b_hi = (256u >> 4) | (1u << 14);  // SBO 256 B
#pragma unroll
    for (int j = 0; j < KS; ++j) {
        const uint64_t db = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo0 + (uint32_t)(j * N * 2));
        const uint32_t a_lo = a_lo0 + (uint32_t)j * cp16;
#pragma unroll
        for (int t = 0; t < TM; ++t)
            umma_bf16(acc + (uint32_t)(t * N), ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)(t * 128)), db, idesc,
                      j == 0 ? accumulate_first : 1u);
    }
}

// MMA role of the pipelined kernel, run by ONE elected thread.  The tensor pipe needs max(N/2, 32 + N/4)
// cycles per 128 x N x 16 MMA (tools/umma_rate.cu) and is starved by anything slower than a handful of
// uniform-datapath instructions per MMA; every register -> uniform-register move in front of a
// tcgen05.mma costs tens of cycles.  K steps are therefore issued in fully unrolled groups of KS
// (KS divides both the channel pairs per tap and the K steps per ring stage).
template <int N, int KW, int TM, int KS>
__device__ __forceinline__ void mma_role(const ConvArgs& a, PipeBars& bars, uint32_t tmem_base, uint32_t pbuf16, uint32_t wring_addr,
                                         int n_items, uint32_t stage_bytes) {
    const uint32_t idesc = umma_idesc_bf16(N, kMajorK, kMajorK);
    const uint32_t b_lbo = (128u >> 4) << 16;
    const uint32_t a_lbo = ((uint32_t)a.plane_stride >> 4) << 16;
    const uint32_t tap00 = (uint32_t)(a.halo - a.pad * a.PW - a.pad);  // pixel offset of tap (0,0)
    const uint32_t cp16 = (uint32_t)(2 * a.plane_stride) >> 4;
    const int G = a.planes >> 1;
    int logG = 0;
    while ((1 << logG) < G) ++logG;
    const int S = a.kpg / a.ksps, groups_per_stage = a.ksps / KS;
    const int s0 = a.rotate ? (int)(((long)blockIdx.x * S) / gridDim.x) : 0;
    uint32_t it = 0, p = 0, ws = 0, wphase = 0;
    bool first_item = true;
    long long t_acc = 0, t_pl = 0, t_w = 0, t0 = 0, tq = 0;
    const bool prof = a.dbg != nullptr;
    if (prof) t0 = clock64();
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t ab = it & 1u;
        if (prof) tq = clock64();
        mbar_wait(&bars.acc_empty[ab], ((it >> 1) & 1u) ^ 1u, a.fault);
        if (prof) t_acc += clock64() - tq;
        tc_fence_after();
        const uint32_t acc = tmem_base + ab * (uint32_t)(TM * N);
        uint32_t accumulate = 0;
        for (int cg = 0; cg < a.ncg; ++cg, ++p) {
            const uint32_t buf = p & 1u;
            if (prof) tq = clock64();
            mbar_wait(&bars.p_full[buf], (p >> 1) & 1u, a.fault);
            if (prof) t_pl += clock64() - tq;
            tc_fence_after();
            const uint32_t a_buf = pbuf16 + ((buf * (uint32_t)a.buf_bytes) >> 4) + tap00;
            for (int i = 0; i < S; ++i) {
                int st = s0 + i;
                if (st >= S) st -= S;
                if (prof) tq = clock64();
                if (!a.resident || first_item) mbar_wait(&bars.w_full[ws], wphase, a.fault);
                if (prof) t_w += clock64() - tq;
                tc_fence_after();
                uint32_t b_lo = (((wring_addr + ws * stage_bytes) & 0x3FFFFu) >> 4) | b_lbo;
                int klin = st * a.ksps;
                for (int gi = 0; gi < groups_per_stage; ++gi, klin += KS) {
                    const int tap = klin >> logG, cp0 = klin & (G - 1);
                    const int ty = tap / KW, tx = tap - ty * KW;
                    const uint32_t a_lo = (a_buf + (uint32_t)(ty * a.PW + tx) + (uint32_t)cp0 * cp16) | a_lbo;
                    issue_fixed<N, TM, KS>(acc, a_lo, b_lo, cp16, idesc, accumulate);
                    accumulate = 1u;
                    b_lo += (uint32_t)(KS * N * 2);
                }
                if (!a.resident) umma_commit(&bars.w_empty[ws]);
                if (++ws == (uint32_t)a.nstages) { ws = 0; wphase ^= 1u; }
            }
            umma_commit(&bars.p_empty[buf]);
        }
        umma_commit(&bars.acc_full[ab]);
        if (a.resident) { ws = 0; wphase = 0; }
        first_item = false;
    }
    if (prof) {
        unsigned long long* o = a.dbg + (size_t)blockIdx.x * 8;
        o[0] = (unsigned long long)(clock64() - t0); o[1] = t_acc; o[2] = t_pl; o[3] = t_w; o[4] = it;
    }
}

// MMA role for the 8-channel source of encoder conv 0 (CVAE_KTAB_PAIR8): 13 K steps cover the 25 taps two at a
// time -- the two 8-channel halves of a K = 16 step are the same plane read one pixel (LBO 16 B) or one row
// (LBO PW * 16 B) apart.  The 13 KB weight block is loaded once and stays resident.
template <int N, int TM>
__device__ __forceinline__ void mma_role_pair8(const ConvArgs& a, PipeBars& bars, uint32_t tmem_base, uint32_t pbuf16, uint32_t wring_addr,
                                               int n_items) {
    const uint32_t idesc = umma_idesc_bf16(N, kMajorK, kMajorK);
    const uint32_t a_hi = (128u >> 4) | (1u << 14), b_hi = (256u >> 4) | (1u << 14);
    const uint32_t b_lo0 = ((wring_addr & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
    uint32_t it = 0;
    const bool prof = a.dbg != nullptr;
    long long t0 = 0, tq = 0, t_acc = 0, t_pl = 0;
    if (prof) t0 = clock64();
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t ab = it & 1u, buf = it & 1u;
        if (prof) tq = clock64();
        mbar_wait(&bars.acc_empty[ab], ((it >> 1) & 1u) ^ 1u, a.fault);
        if (prof) { t_acc += clock64() - tq; tq = clock64(); }
        mbar_wait(&bars.p_full[buf], (it >> 1) & 1u, a.fault);
        if (prof) t_pl += clock64() - tq;
        if (it == 0) mbar_wait(&bars.w_full[0], 0, a.fault);
        tc_fence_after();
        const uint32_t acc = tmem_base + ab * (uint32_t)(TM * N);
        const uint32_t a_buf = pbuf16 + ((buf * (uint32_t)a.buf_bytes) >> 4) + (uint32_t)a.halo;
#pragma unroll
        for (int i = 0; i < 13; ++i) {
            int off;          // first tap of the pair, in pixels relative to the output pixel
            uint32_t lbo16;   // distance to the second tap, in 16-byte units
            if (i < 10) { off = ((i >> 1) - 2) * a.PW + ((i & 1) * 2 - 2); lbo16 = 1u; }
            else if (i < 12) { off = ((i - 10) * 2 - 2) * a.PW + 2; lbo16 = (uint32_t)a.PW; }
            else { off = 2 * a.PW + 2; lbo16 = 1u; }
            const uint32_t a_lo = (uint32_t)((int)a_buf + off) | (lbo16 << 16);
            const uint64_t db = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo0 + (uint32_t)(i * N * 2));
#pragma unroll
            for (int t = 0; t < TM; ++t)
                umma_bf16(acc + (uint32_t)(t * N), ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)(t * 128)), db, idesc, i > 0 ? 1u : 0u);
        }
        umma_commit(&bars.p_empty[buf]);
        umma_commit(&bars.acc_full[ab]);
    }
    if (prof) {
        unsigned long long* o = a.dbg + (size_t)blockIdx.x * 8;
        o[0] = (unsigned long long)(clock64() - t0); o[1] = t_acc; o[2] = t_pl; o[3] = 0; o[4] = it;
    }
}

template <int N, int KW, int TM>
__device__ __forceinline__ void mma_role_ks(const ConvArgs& a, PipeBars& bars, uint32_t tmem_base, uint32_t pbuf16,
                                            uint32_t wring_addr, int n_items, uint32_t stage_bytes) {
    if (a.kgroup == 4) mma_role<N, KW, TM, 4>(a, bars, tmem_base, pbuf16, wring_addr, n_items, stage_bytes);
    else if (a.kgroup == 2) mma_role<N, KW, TM, 2>(a, bars, tmem_base, pbuf16, wring_addr, n_items, stage_bytes);
    else mma_role<N, KW, TM, 1>(a, bars, tmem_base, pbuf16, wring_addr, n_items, stage_bytes);
}

// --------------------------------------------------------------------------------------------
// pipelined kernel
// --------------------------------------------------------------------------------------------
// Warps 0-3 epilogue, then the plane producers, then the weight producer and the MMA issuer (the last two warps).
// Encoder conv 0's producers convert fp32 frames themselves and were its limiter: that loader gets eight producer warps.
template <int LOADER> constexpr int pipe_threads() { return LOADER == CVAE_LOAD_NCHW3 ? 448 : 320; }
template <int LOADER> constexpr int pipe_producers() { return pipe_threads<LOADER>() - 192; }
static size_t kStageBytesMax = getenv("CVAE_STAGE_KB") ? (size_t)atoi(getenv("CVAE_STAGE_KB")) * 1024 : 0;   // 0: automatic
static int kMaxGroupPlanes = getenv("CVAE_GROUP_PLANES") ? atoi(getenv("CVAE_GROUP_PLANES")) : 0;         // 0: automatic
static constexpr size_t kDynSmemMax = 216 * 1024;  // 227 KB per CTA minus static shared memory (barriers, statistics scratch)

template <int LOADER, int EPI, int N, int KW>
__global__ void __launch_bounds__(pipe_threads<LOADER>(), 1) conv_pipe_kernel(const ConvArgs a) {
    constexpr int kPipeThreads = pipe_threads<LOADER>(), kProducerThreads = pipe_producers<LOADER>();
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ PipeBars bars;
    uint64_t* const w_full = bars.w_full; uint64_t* const w_empty = bars.w_empty;
    uint64_t* const p_full = bars.p_full; uint64_t* const p_empty = bars.p_empty;
    uint64_t* const acc_full = bars.acc_full; uint64_t* const acc_empty = bars.acc_empty;
    __shared__ uint32_t tmem_slot;
    __shared__ float stat_scratch[(EPI == CVAE_EPI_STATS) ? 4 * 32 * 17 : 1];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t kstep_bytes = (uint32_t)N * 32u;
    const uint32_t stage_bytes = (uint32_t)a.ksps * kstep_bytes;

    uint8_t* pbuf = smem;
    uint8_t* wring = smem + 2 * (size_t)a.buf_bytes;

    if (tid == 0) {
        for (int s = 0; s < a.nstages; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&p_full[b], kProducerThreads);
            mbar_init(&p_empty[b], 1);
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);
        }
        mbar_fence_init();
    }
    uint32_t ncols = 32;
    while (ncols < (uint32_t)(2 * a.tm * N)) ncols <<= 1;
    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    grid_dependency_sync();     // everything above touched only this CTA's shared memory and TMEM
    const uint32_t tmem_base = tmem_slot;
    const int n_items = a.num_chunks * a.n_blocks;
    const int spg = a.kpg / a.ksps;  // weight stages per channel group

    if (warp < 4) {
        // ================================ epilogue ================================================
        constexpr int kStatRegs = (EPI != CVAE_EPI_STATS) ? 1 : (N == 32 ? 32 : N / 16);
        float s1[kStatRegs], s2[kStatRegs];
#pragma unroll
        for (int g = 0; g < kStatRegs; ++g) s1[g] = s2[g] = 0.f;
        int cur_nb = -1;
        uint32_t it = 0;
        long long t_epi = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int nb = item / a.num_chunks, chunk = item - nb * a.num_chunks;
            if (nb != cur_nb) {
                if (cur_nb >= 0) flush_stats<EPI, N>(a, cur_nb, warp, lane, s1, s2, stat_scratch);
                cur_nb = nb;
            }
            const uint32_t ab = it & 1u;
            mbar_wait(&acc_full[ab], (it >> 1) & 1u, a.fault);
            tc_fence_after();
            const long long te = clock64();
            const int v0 = a.pad * a.PW + chunk * a.tm * 128;
            for (int t = 0; t < a.tm && !(a.dbg_flags & 1); ++t)
                epilogue_tile<EPI, N>(a, tmem_base + ab * (uint32_t)(a.tm * N) + (uint32_t)(t * N), v0 + t * 128, nb, warp,
                                      lane, s1, s2, stat_scratch);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
            t_epi += clock64() - te;
        }
        if (a.dbg && tid == 0) a.dbg[(size_t)blockIdx.x * 8 + 7] = t_epi;
        if (cur_nb >= 0) flush_stats<EPI, N>(a, cur_nb, warp, lane, s1, s2, stat_scratch);
    } else if (warp == kPipeThreads / 32 - 1) {
        // ================================ MMA issuer (highest warp id: the scheduler favours it) ==============================================
        if (elect_one()) {
            const uint32_t pbuf16 = (smem_u32(pbuf) & 0x3FFFFu) >> 4, wring_addr = smem_u32(wring);
            if constexpr (LOADER == CVAE_LOAD_NCHW3) {
                switch (a.tm) {
                    case 2: mma_role_pair8<N, 2>(a, bars, tmem_base, pbuf16, wring_addr, n_items); break;
                    case 4: mma_role_pair8<N, 4>(a, bars, tmem_base, pbuf16, wring_addr, n_items); break;
                    case 8: mma_role_pair8<N, 8>(a, bars, tmem_base, pbuf16, wring_addr, n_items); break;
                    default: mma_role_pair8<N, 1>(a, bars, tmem_base, pbuf16, wring_addr, n_items); break;
                }
            } else
            switch (a.tm) {
                case 1: mma_role_ks<N, KW, 1>(a, bars, tmem_base, pbuf16, wring_addr, n_items, stage_bytes); break;
                case 2: mma_role_ks<N, KW, 2>(a, bars, tmem_base, pbuf16, wring_addr, n_items, stage_bytes); break;
                case 3: mma_role_ks<N, KW, 3>(a, bars, tmem_base, pbuf16, wring_addr, n_items, stage_bytes); break;
                case 4: mma_role_ks<N, KW, 4>(a, bars, tmem_base, pbuf16, wring_addr, n_items, stage_bytes); break;
                case 6: mma_role_ks<N, KW, 6>(a, bars, tmem_base, pbuf16, wring_addr, n_items, stage_bytes); break;
                default: mma_role_ks<N, KW, 8>(a, bars, tmem_base, pbuf16, wring_addr, n_items, stage_bytes); break;
            }
        }
        __syncwarp();
    } else if (warp == kPipeThreads / 32 - 2) {
        // ================================ weight producer =========================================
        const uint8_t* wbase = reinterpret_cast<const uint8_t*>(a.wpack);
        const int cpairs_total = a.ksteps / (a.KW * a.KW);  // GENERIC only
        const int G = a.planes >> 1;
        const int s0w = a.rotate ? (int)(((long)blockIdx.x * spg) / gridDim.x) : 0;
        uint32_t ws = 0, wphase = 0;
        bool first_item = true;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            if (a.resident && !first_item) break;
            const int nb = item / a.num_chunks;
            const int col0 = nb * N;
            const int pb = col0 / a.nb_pack, sub = col0 - pb * a.nb_pack;
            for (int cg = 0; cg < a.ncg; ++cg) {
                for (int i = 0; i < spg; ++i) {
                    int st = s0w + i;
                    if (st >= spg) st -= spg;
                    if (!a.resident) mbar_wait(&w_empty[ws], wphase ^ 1u, a.fault);
                    int kglob;  // first global K step of this stage
                    if (a.ncg == 1) kglob = st * a.ksps;
                    else {
                        const int kl = st * a.ksps, tap = kl / G, j0 = kl - tap * G;
                        kglob = tap * cpairs_total + cg * G + j0;
                    }
                    if (elect_one()) {
                        mbar_expect_tx(&w_full[ws], stage_bytes);
                        uint8_t* dst = wring + (size_t)ws * stage_bytes;
                        if (N == a.nb_pack) {
                            bulk_g2s(dst, wbase + ((size_t)pb * a.ksteps + kglob) * a.nb_pack * 32, stage_bytes, &w_full[ws]);
                        } else {
                            for (int ks = 0; ks < a.ksps; ++ks)
                                bulk_g2s(dst + (size_t)ks * kstep_bytes,
                                         wbase + (((size_t)pb * a.ksteps + kglob + ks) * a.nb_pack + sub) * 32, kstep_bytes,
                                         &w_full[ws]);
                        }
                    }
                    __syncwarp();
                    if (++ws == (uint32_t)a.nstages) { ws = 0; wphase ^= 1u; }
                }
            }
            first_item = false;
        }
    } else {
        // ================================ plane producers =========================================
        const int ptid = tid - 128;  // warps 4-7
        uint32_t p = 0;
        bool alive = true;
        long long t_fill = 0, t_pe = 0, tq;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int nb = item / a.num_chunks, chunk = item - nb * a.num_chunks;
            const int v0 = a.pad * a.PW + chunk * a.tm * 128;
            for (int cg = 0; cg < a.ncg; ++cg, ++p) {
                const uint32_t buf = p & 1u;
                tq = clock64();
                if (alive) alive = mbar_wait(&p_empty[buf], ((p >> 1) & 1u) ^ 1u, a.fault);
                t_pe += clock64() - tq;
                tq = clock64();
                if (!(a.dbg_flags & 2) || p < 2) {
                    if constexpr (LOADER == CVAE_LOAD_NHWC || LOADER == CVAE_LOAD_S2D)
                        fill_planes_async<LOADER>(a.ps, a.dPW, a.dIH, pbuf + (size_t)buf * a.buf_bytes, a.plane_stride, v0 - a.halo, a.L,
                                                  cg * a.planes, a.planes, ptid, kProducerThreads);
                    else   // fp32 NCHW sources are converted by the producer threads themselves
                        fill_planes<LOADER>(a.ps, a.dPW, a.dIH, pbuf + (size_t)buf * a.buf_bytes, a.plane_stride, v0 - a.halo, a.L, ptid, kProducerThreads);
                }
                cp_async_wait_all();
                fence_proxy_async();
                mbar_arrive(&p_full[buf]);
                t_fill += clock64() - tq;
            }
        }
        if (a.dbg && ptid == 0) {
            unsigned long long* o = a.dbg + (size_t)blockIdx.x * 8;
            o[5] = t_fill;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, ncols);
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
static const bool g_debug = getenv("CVAE_DEBUG") != nullptr;
// experiment switches, read once at load time (never on the launch path)
static const int g_rotate = getenv("CVAE_NO_ROTATE") ? 0 : 1;
static const int g_dbg_flags = getenv("CVAE_DBG_FLAGS") ? atoi(getenv("CVAE_DBG_FLAGS")) : 0;
static unsigned long long* g_dbg_counters = nullptr;

template <int LOADER, int EPI, int N, int KW>
static int launch_pipe(const ConvArgs& a, size_t smem, cudaStream_t stream) {
    auto kern = conv_pipe_kernel<LOADER, EPI, N, KW>;
    CVAE_OPT_IN_SMEM(kern, smem);
    const int items = a.num_chunks * a.n_blocks;
    int gx = sm_count();
    if (gx > items) gx = items;
    // balance: every CTA gets the same number of items where possible
    const int waves = (items + gx - 1) / gx;
    gx = (items + waves - 1) / waves;
    if (g_debug)
        fprintf(stderr, "conv_pipe<L%d,E%d,N%d> B=%d %dx%d planes=%dx%d ksteps=%d kpg=%d ksps=%d stages=%d%s tm=%d chunks=%d "
                        "nblk=%d smem=%zu grid=%d\n", LOADER, EPI, N, a.B, a.H, a.W, a.planes, a.ncg, a.ksteps, a.kpg, a.ksps,
                a.nstages, a.resident ? "(resident)" : "", a.tm, a.num_chunks, a.n_blocks, smem, gx);
    cvae::launch(kern, gx, pipe_threads<LOADER>(), smem, stream, a);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

}  // namespace cvae

using namespace cvae;

// Profiling aid: when set, every pipelined conv CTA writes 8 cycle counters (MMA thread: total, wait for
// accumulator, planes, weights, items; producers: fill, wait; epilogue) to buf[blockIdx.x * 8 ..].
extern "C" void cvae_conv_debug_counters(void* device_buf) { g_dbg_counters = (unsigned long long*)device_buf; }

extern "C" int cvae_conv_ksteps(int ksize, int src_channels, int ktab) {
    if (ktab == CVAE_KTAB_PAIR8) return 13;
    return ksize * ksize * (src_channels / 16);
}

// Tiling policy of the pipelined kernel for one (channel-group size, weight-stage size) choice; fails with
// CVAE_EINVAL when the plane buffers and a useful weight ring do not fit shared memory.
static int plan_pipe(const cvae_conv_desc* d, ConvArgs& a, int all_planes, int max_group_planes, size_t stage_cap, size_t* smem_out,
                     int* n_out) {
    // ---- tiling policy of the pipelined kernel -----------------------------------------------------
    const long total_v = (long)a.B * a.IH * a.PW - (long)a.pad * a.PW;  // pixels from first to last valid row
    const int total_tiles = (int)((total_v + 127) / 128);
    const int sms = sm_count();
    // channel groups of <= 128 channels (16 planes)
    a.planes = all_planes;
    a.ncg = 1;
    while (a.planes > max_group_planes && a.planes % 4 == 0) { a.planes /= 2; a.ncg *= 2; }
    if (d->ktab == CVAE_KTAB_PAIR8) a.kpg = 13;
    else a.kpg = d->ksize * d->ksize * (a.planes / 2);

    // Work decomposition.  N block = the packed block (<= 128 columns; narrower blocks only cost shared-memory
    // bandwidth: an SS-mode MMA reads 4 KB of A regardless of N).  tm (tiles per pass, sharing every weight
    // stage) from a small cost model fitted to tools/conv_bench.py sweeps on B200: per CTA
    //   max(MMA cycles, weight-stream cycles at ~17 B/clk/SM of shared L2 reads) + per-item pipeline fill.
    int N = d->n_block > 0 ? d->n_block : a.nb_pack;
    int tm = d->tm;
    if (tm <= 0) {
        double best = 1e30;
        tm = 1;
        const double mma_cyc = (N / 2.0 > 32.0 + N / 4.0 ? N / 2.0 : 32.0 + N / 4.0) * 1.3;
        for (int t : {1, 2, 3, 4, 6, 8}) {
            if (2 * t * N > 512 || (t > 1 && t > total_tiles)) continue;
            const size_t L = ((size_t)t * 128 + 2 * a.halo + 8) | 1;
            const size_t buf = ((size_t)a.planes * L * 16 + 1023) & ~(size_t)1023;
            if (2 * buf + 32 * 1024 > kDynSmemMax) continue;
            const long chunks = (total_tiles + t - 1) / t;
            const long items = chunks * (d->n_total / N);
            long gx = items < sms ? items : sms;
            const long per_cta = (items + gx - 1) / gx;
            const double t_mma = (double)per_cta * t * a.ksteps * mma_cyc;
            double bw = 2500.0 / (double)gx;
            if (bw > 40.0) bw = 40.0;
            const double t_w = (double)per_cta * a.ksteps * N * 32.0 / bw;
            const double est = (t_mma > t_w ? t_mma : t_w) + (double)per_cta * (2500.0 + 600.0 * t);
            if (est < best) { best = est; tm = t; }
        }
    }
    CVAE_REQUIRE(N >= 16 && N <= 128 && a.nb_pack % N == 0, CVAE_EINVAL, "conv_gemm: n_block %d", N);
    if (tm > total_tiles) tm = total_tiles;
    while (tm > 1 && 2 * tm * N > 512) --tm;   // two accumulator sets must fit the 512 TMEM columns
    if (tm == 5) tm = 4;
    if (tm == 7) tm = 6;
    const bool pair8 = d->ktab == CVAE_KTAB_PAIR8;
    CVAE_REQUIRE(pair8 == (d->loader == CVAE_LOAD_NCHW3), CVAE_EINVAL, "conv_gemm: the paired-tap K order goes with the frame loader");
    if (pair8) tm = tm >= 8 ? 8 : (tm >= 4 ? 4 : (tm >= 2 ? 2 : 1));
    CVAE_REQUIRE(tm >= 1 && 2 * tm * N <= 512, CVAE_EINVAL, "conv_gemm: tm %d x n_block %d exceeds tensor memory", tm, N);
    a.n_blocks = d->n_total / N;
    a.tm = tm;
    a.L = (tm * 128 + 2 * a.halo + 8) | 1;   // odd: the lanes copying the planes of one pixel spread over the banks
    a.plane_stride = a.L * 16;
    a.buf_bytes = (int)((((size_t)a.planes * a.plane_stride) + 1023) & ~(size_t)1023);
    a.num_chunks = (total_tiles + tm - 1) / tm;
    CVAE_REQUIRE((size_t)2 * a.buf_bytes < (1u << 18), CVAE_EINVAL, "conv_gemm: planes exceed descriptor range");

    // weight stages: ~4-8 KB each, as many as fit; resident when the whole matrix fits
    const int G = a.planes / 2;
    const int run = (a.ncg == 1 || d->ktab == CVAE_KTAB_PAIR8) ? a.kpg : G;  // contiguous K steps in global memory
    CVAE_REQUIRE((G & (G - 1)) == 0, CVAE_EINVAL, "conv_gemm: channel pairs per group must be a power of two");
    a.rotate = g_rotate;
    a.kgroup = (G % 4 == 0) ? 4 : (G % 2 == 0 ? 2 : 1);
    while (a.kgroup > 1 && (size_t)a.kgroup * N * 32 > stage_cap) a.kgroup /= 2;
    a.ksps = a.kgroup;
    for (int c = a.kgroup; c <= run; c += a.kgroup)
        if (run % c == 0 && a.kpg % c == 0 && (size_t)c * N * 32 <= stage_cap) a.ksps = c;
    if (pair8) a.ksps = a.kgroup = 13;   // one resident 13-step block
    const size_t stage_bytes = (size_t)a.ksps * N * 32;
    const long budget = (long)kDynSmemMax - 2 * (long)a.buf_bytes - (long)a.kpg * 8 - 64;
    CVAE_REQUIRE(budget >= (long)(2 * stage_bytes), CVAE_EINVAL, "conv_gemm: shape does not fit shared memory");
    int nstages = (int)(budget / (long)stage_bytes);
    if (nstages > kMaxStages) nstages = kMaxStages;
    const int stages_total = a.ncg * (a.kpg / a.ksps);
    a.resident = (a.n_blocks == 1 && stages_total <= nstages) ? 1 : 0;
    if (a.resident) nstages = stages_total;
    a.nstages = nstages;
    *smem_out = 2 * (size_t)a.buf_bytes + (size_t)nstages * stage_bytes + (size_t)a.kpg * 8 + 64;
    *n_out = N;
    return CVAE_OK;

}

extern "C" int64_t cvae_conv_gemm_workspace_bytes(const cvae_conv_desc* d) {
    if (!d) return -1;
    return d->ktab == CVAE_KTAB_BLOCK64 ? conv_wa_workspace_bytes(d) : 0;
}

extern "C" int cvae_conv_gemm(const cvae_conv_desc* d, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(d != nullptr, CVAE_EINVAL, "conv_gemm: null descriptor");
    CVAE_REQUIRE(d->ksize == 5 || d->ksize == 3, CVAE_EINVAL, "conv_gemm: ksize %d", d->ksize);
    CVAE_REQUIRE(d->batch > 0 && d->height > 0 && d->width > 0, CVAE_EINVAL, "conv_gemm: empty shape");
    CVAE_REQUIRE(d->n_total % 16 == 0 && d->n_total > 0, CVAE_EINVAL, "conv_gemm: n_total %d", d->n_total);
    CVAE_REQUIRE(d->src && d->wpack && d->out, CVAE_EINVAL, "conv_gemm: null tensor");
    if (d->ktab == CVAE_KTAB_BLOCK64) return conv_wa_dispatch(d, stream);
    if (d->ktab == CVAE_KTAB_PAIR8)
        CVAE_REQUIRE(d->src_channels == 8 && d->ksize == 5, CVAE_EINVAL, "conv_gemm: PAIR8 needs 8 channels, 5x5");
    else
        CVAE_REQUIRE(d->src_channels % 16 == 0, CVAE_EINVAL, "conv_gemm: src_channels %d", d->src_channels);

    ConvArgs a{};
    a.B = d->batch; a.H = d->height; a.W = d->width; a.KW = d->ksize; a.pad = d->ksize / 2;
    a.PW = a.W + a.pad; a.IH = a.H + a.pad;
    a.dPW = make_fastdiv(a.PW); a.dIH = make_fastdiv(a.IH);
    a.nb_pack = d->n_total < 128 ? d->n_total : 128;
    CVAE_REQUIRE(d->n_total % a.nb_pack == 0, CVAE_EINVAL, "conv_gemm: n_total %d not a multiple of %d", d->n_total, a.nb_pack);
    a.c_total = d->n_total;
    a.ksteps = cvae_conv_ksteps(d->ksize, d->src_channels, d->ktab);
    a.ktab_mode = d->ktab;
    a.halo = a.pad * a.PW + a.pad;
    a.wpack = (const __nv_bfloat16*)d->wpack; a.out = d->out;
    a.bias = d->bias; a.act = (const __nv_bfloat16*)d->act; a.stats = d->stats;
    a.fault = fault_flag();
    a.dbg = g_dbg_counters;
    a.dbg_flags = g_dbg_flags;
    CVAE_REQUIRE(a.fault != nullptr, CVAE_ECUDA, "conv_gemm: fault flag unavailable");
    const int all_planes = d->src_channels / 8;
    a.ps = PlaneSrc{a.B, a.H, a.W, a.pad, a.PW, a.IH, all_planes,
                    (d->loader == CVAE_LOAD_S2D) ? d->src_channels / 4 : d->src_channels, 0, d->src, d->src2};


    // ---- tiling policy of the pipelined kernel: largest channel groups (fewest plane refills, longest contiguous
    // weight runs) and weight stages that fit; tools/conv_bench.py sweeps showed 16-plane groups and 32 KB stages
    // ahead wherever they fit (the MMA thread pays a fixed cost per stage) ------------------------------------
    size_t smem = 0;
    int N = 0;
    {
        const int gp_env = kMaxGroupPlanes, gps[2] = {gp_env > 0 ? gp_env : 16, 8};
        const size_t caps[2] = {kStageBytesMax ? kStageBytesMax : 32 * 1024, 16 * 1024};
        int rc = CVAE_EINVAL;
        for (int gi = 0; gi < 2 && rc != CVAE_OK; ++gi)
            for (int ci = 0; ci < 2 && rc != CVAE_OK; ++ci) {
                ConvArgs trial = a;
                rc = plan_pipe(d, trial, all_planes, gps[gi], caps[ci], &smem, &N);
                if (rc == CVAE_OK) a = trial;
            }
        if (rc != CVAE_OK) return rc;
    }

#define CVAE_CASE(L_, E_, N_, K_) \
    if (d->loader == (L_) && d->epilogue == (E_) && N == (N_) && d->ksize == (K_)) return launch_pipe<L_, E_, N_, K_>(a, smem, stream);
#define CVAE_CASES(L_, E_, K_) CVAE_CASE(L_, E_, 16, K_) CVAE_CASE(L_, E_, 32, K_) CVAE_CASE(L_, E_, 64, K_) CVAE_CASE(L_, E_, 128, K_)
    CVAE_CASES(CVAE_LOAD_NHWC, CVAE_EPI_STATS, 5)
    CVAE_CASES(CVAE_LOAD_NHWC, CVAE_EPI_BIAS_RELU, 5)
    CVAE_CASES(CVAE_LOAD_NHWC, CVAE_EPI_PHASE_BIAS_RELU, 3)
    CVAE_CASE(CVAE_LOAD_NHWC, CVAE_EPI_PHASE_BIAS_TANH, 16, 3)
    CVAE_CASES(CVAE_LOAD_NHWC, CVAE_EPI_PLAIN, 5)
    CVAE_CASES(CVAE_LOAD_S2D, CVAE_EPI_MASK, 3)
    CVAE_CASE(CVAE_LOAD_NCHW3, CVAE_EPI_STATS, 32, 5)
    CVAE_CASE(CVAE_LOAD_S2D_NCHW3_DTANH, CVAE_EPI_MASK, 32, 3)
#undef CVAE_CASES
#undef CVAE_CASE
    set_error("conv_gemm: no kernel for loader %d epilogue %d N %d", d->loader, d->epilogue, N);
    return CVAE_EINVAL;
}
