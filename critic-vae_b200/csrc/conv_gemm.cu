// Implicit-GEMM 5x5 / 3x3 convolution on tcgen05 for 64x64-and-smaller feature maps.
//
// Layout idea ("halo planes"): the CTA's slice of the input is staged ONCE in shared memory as
// planes [channel/8][virtual pixel][8 channels], i.e. in the no-swizzle UMMA core-matrix layout
// with pixels as rows.  Images are laid out in a linear "virtual pixel" space with `pad` shared
// zero columns per row and `pad` shared zero rows per image, so the A operand of every filter tap
// is the SAME shared-memory tile read through a descriptor whose start address is shifted by
// (dy*PW + dx) pixels.  No im2col, no re-fetch per tap: 25x (9x) reuse out of shared memory.
// Weights stream through a small mbarrier ring filled by cp.async.bulk and are reused by all
// `tm` 128-pixel tiles of the pass, whose accumulators sit side by side in TMEM.
//
// Reference ops replaced: nn.Conv2d(5,1,2) at vae_nets.py:69,74,79,84,117,121,125,129,133, the
// nn.Upsample(2) at :119,123,127,131 (folded: conv5x5(up2(x)) == depth_to_space(conv3x3_4C(x)))
// and their autograd data-gradients.
#include "common.cuh"
#include "umma.cuh"
#include "planes.cuh"

namespace cvae {

struct ConvArgs {
    int B, H, W, pad, KW;
    int PW, IH;          // W + pad, H + pad
    int planes;          // 8-channel planes of the A operand
    int n_blocks;        // N split (blockIdx.y)
    int c_total;         // n_total
    int ksteps, ksps;    // K=16 steps in total / per weight stage
    int tm;              // tiles per pass
    int num_chunks;
    int halo;            // pad*PW + pad
    int L;               // pixels per plane (tm*128 + 2*halo + 8)
    int plane_stride;    // bytes
    int ktab_mode;
    int nstages;
    PlaneSrc ps;         // where the A-operand planes come from
    const __nv_bfloat16* wpack;
    void* out;
    const float* bias;
    const __nv_bfloat16* act;
    double* stats;
    int* fault;
};

static constexpr int kThreads = 192;  // warps 0-3: loader + epilogue, 4: weight producer, 5: MMA

// --------------------------------------------------------------------------------------------
// kernel
// --------------------------------------------------------------------------------------------
template <int LOADER, int EPI, int N>
__global__ void __launch_bounds__(kThreads) conv_gemm_kernel(const ConvArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int kMaxStages = 6;
    __shared__ uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_acc;
    __shared__ uint32_t tmem_slot;
    __shared__ float stat_scratch[(EPI == CVAE_EPI_STATS) ? 4 * 32 * 17 : 1];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = blockIdx.y;
    const uint32_t stage_bytes = (uint32_t)a.ksps * N * 32;

    uint8_t* planes = smem;
    uint8_t* wring = smem + (((size_t)a.planes * a.plane_stride + 1023) & ~(size_t)1023);
    uint2* ktab = reinterpret_cast<uint2*>(wring + (size_t)a.nstages * stage_bytes);

    // ---- one-time setup -------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < a.nstages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(&bar_acc, 1);
        mbar_fence_init();
    }
    uint32_t ncols = 32;
    while (ncols < (uint32_t)(a.tm * N)) ncols <<= 1;
    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    for (int i = tid; i < a.ksteps; i += kThreads) {
        uint2 e;
        if (a.ktab_mode == CVAE_KTAB_GENERIC) {
            const int cpairs = a.planes >> 1;
            const int tap = i / cpairs, cp = i - tap * cpairs;
            const int dy = tap / a.KW - a.pad, dx = tap % a.KW - a.pad;
            e.x = (uint32_t)(a.halo + dy * a.PW + dx) * 16u + (uint32_t)cp * 2u * a.plane_stride;
            e.y = (uint32_t)a.plane_stride;
        } else {  // PAIR8: 5x5 taps of an 8-channel source, two taps per K step
            if (i < 10) {
                const int ky = i >> 1, kx = (i & 1) * 2;
                e.x = (uint32_t)(a.halo + (ky - 2) * a.PW + (kx - 2)) * 16u;
                e.y = 16u;
            } else if (i < 12) {
                const int ky = (i - 10) * 2;
                e.x = (uint32_t)(a.halo + (ky - 2) * a.PW + 2) * 16u;
                e.y = (uint32_t)a.PW * 16u;
            } else {
                e.x = (uint32_t)(a.halo + 2 * a.PW + 2) * 16u;
                e.y = 16u;
            }
        }
        ktab[i] = e;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t idesc = umma_idesc_bf16(N, kMajorK, kMajorK);
    const __nv_bfloat16* wsrc = a.wpack + (size_t)nb * a.ksteps * N * 16;
    const int stages_per_chunk = a.ksteps / a.ksps;

    uint32_t ring_stage = 0, ring_phase = 0;  // producer and MMA thread walk the ring in lock step
    uint32_t acc_phase = 0;
    bool alive = true;

    float s1[(EPI == CVAE_EPI_STATS) ? N / 16 : 1], s2[(EPI == CVAE_EPI_STATS) ? N / 16 : 1];
#pragma unroll
    for (int g = 0; g < ((EPI == CVAE_EPI_STATS) ? N / 16 : 1); ++g) s1[g] = s2[g] = 0.f;

    for (int chunk = blockIdx.x; chunk < a.num_chunks; chunk += gridDim.x) {
        const int v0 = a.pad * a.PW + chunk * a.tm * 128;  // first output pixel of this pass
        fill_planes<LOADER>(a.ps, planes, a.plane_stride, v0 - a.halo, a.L, tid, kThreads);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();

        if (warp == 4) {
            if (lane == 0) {
                for (int st = 0; st < stages_per_chunk && alive; ++st) {
                    alive = mbar_wait(&bar_empty[ring_stage], ring_phase ^ 1, a.fault);
                    mbar_expect_tx(&bar_full[ring_stage], stage_bytes);
                    bulk_g2s(wring + (size_t)ring_stage * stage_bytes,
                             reinterpret_cast<const uint8_t*>(wsrc) + (size_t)st * stage_bytes,
                             stage_bytes, &bar_full[ring_stage]);
                    if (++ring_stage == (uint32_t)a.nstages) { ring_stage = 0; ring_phase ^= 1; }
                }
            }
            __syncwarp();
        } else if (warp == 5) {
            if (lane == 0) {
                const uint32_t planes_addr = smem_u32(planes);
                const uint32_t wring_addr = smem_u32(wring);
                for (int st = 0; st < stages_per_chunk && alive; ++st) {
                    alive = mbar_wait(&bar_full[ring_stage], ring_phase, a.fault);
                    tc_fence_after();
                    const uint32_t wb = wring_addr + ring_stage * stage_bytes;
                    // One elected thread feeds the tensor pipe, so the issue loop must stay at a handful
                    // of instructions per MMA: descriptors are built once per K step and the tile index
                    // only bumps the 16-byte-granular start-address field (2048 B per 128-pixel tile).
                    for (int ks = 0; ks < a.ksps; ++ks) {
                        const int kidx = st * a.ksps + ks;
                        const uint2 e = ktab[kidx];
                        uint64_t da = smem_desc(planes_addr + e.x, e.y, 128u);
                        const uint64_t db = smem_desc(wb + (uint32_t)ks * N * 32u, 128u, 256u);
                        const uint32_t accumulate = kidx > 0 ? 1u : 0u;
                        uint32_t tcol = tmem_base;
#pragma unroll 4
                        for (int t = 0; t < a.tm; ++t) {
                            umma_bf16(tcol, da, db, idesc, accumulate);
                            da += 128;   // (2048 >> 4)
                            tcol += N;
                        }
                    }
                    umma_commit(&bar_empty[ring_stage]);
                    if (++ring_stage == (uint32_t)a.nstages) { ring_stage = 0; ring_phase ^= 1; }
                }
                umma_commit(&bar_acc);
            }
            __syncwarp();
        } else {
            // ---- epilogue: warp w owns TMEM lanes [32w, 32w+32) = rows of every tile -----------
            mbar_wait(&bar_acc, acc_phase, a.fault);
            tc_fence_after();
            for (int t = 0; t < a.tm; ++t) {
                const int v = v0 + t * 128 + warp * 32 + lane;
                int vrow = v / a.PW;
                int vcol = v - vrow * a.PW;
                int n = vrow / a.IH;
                int r = vrow - n * a.IH;
                const bool valid = (vcol < a.W) && (r >= a.pad) && (n < a.B);
                const int h = r - a.pad, w = vcol;
                const size_t pix = ((size_t)n * a.H + h) * a.W + w;
#pragma unroll
                for (int g = 0; g < N / 16; ++g) {
                    uint32_t raw[16];
                    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * N + g * 16), raw);
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
                            uint4* o = reinterpret_cast<uint4*>(
                                reinterpret_cast<__nv_bfloat16*>(a.out) + pix * a.c_total + c0);
                            o[0] = make_uint4(p[0], p[1], p[2], p[3]);
                            o[1] = make_uint4(p[4], p[5], p[6], p[7]);
                        }
                    } else if constexpr (EPI == CVAE_EPI_BIAS_RELU || EPI == CVAE_EPI_PLAIN ||
                                         EPI == CVAE_EPI_MASK) {
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
                            uint4* o = reinterpret_cast<uint4*>(
                                reinterpret_cast<__nv_bfloat16*>(a.out) + pix * a.c_total + c0);
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
                            uint4* o = reinterpret_cast<uint4*>(
                                reinterpret_cast<__nv_bfloat16*>(a.out) + opix * cout + co);
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
            acc_phase ^= 1;
        }
        tc_fence_before();
        __syncthreads();  // accumulators drained, planes and ring quiescent: next pass may overwrite
        tc_fence_after();
    }

    if constexpr (EPI == CVAE_EPI_STATS) {
        if (warp < 4 && lane < 16) {
#pragma unroll
            for (int g = 0; g < N / 16; ++g) {
                atomicAdd(a.stats + nb * N + g * 16 + lane, (double)s1[g]);
                atomicAdd(a.stats + a.c_total + nb * N + g * 16 + lane, (double)s2[g]);
            }
        }
    }
    if (warp == 0) tmem_free(tmem_base, ncols);
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
template <int LOADER, int EPI, int N>
static int launch(const ConvArgs& a, size_t smem, cudaStream_t stream) {
    auto kern = conv_gemm_kernel<LOADER, EPI, N>;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        CVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    uint32_t ncols = 32;
    while (ncols < (uint32_t)(a.tm * N)) ncols <<= 1;
    int per_sm = (int)((228 * 1024) / (smem + 2048));          // shared-memory limit (1 KB static + 1 KB reserved per CTA)
    if (per_sm > (int)(512 / ncols)) per_sm = (int)(512 / ncols);  // TMEM limit
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    int gx = sm_count() * per_sm / a.n_blocks;
    if (gx < 1) gx = 1;
    if (gx > a.num_chunks) gx = a.num_chunks;
    dim3 grid(gx, a.n_blocks);
    static const bool debug = getenv("CVAE_DEBUG") != nullptr;
    if (debug)
        fprintf(stderr, "conv_gemm<L%d,E%d,N%d> B=%d %dx%d planes=%d ksteps=%d ksps=%d stages=%d tm=%d chunks=%d smem=%zu "
                        "cols=%u per_sm=%d grid=(%d,%d)\n", LOADER, EPI, N, a.B, a.H, a.W, a.planes, a.ksteps, a.ksps,
                a.nstages, a.tm, a.num_chunks, smem, ncols, per_sm, gx, a.n_blocks);
    kern<<<grid, kThreads, smem, stream>>>(a);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

}  // namespace cvae

using namespace cvae;

extern "C" int cvae_conv_ksteps(int ksize, int src_channels, int ktab) {
    if (ktab == CVAE_KTAB_PAIR8) return 13;
    return ksize * ksize * (src_channels / 16);
}

extern "C" int cvae_conv_gemm(const cvae_conv_desc* d, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(d != nullptr, CVAE_EINVAL, "conv_gemm: null descriptor");
    CVAE_REQUIRE(d->ksize == 5 || d->ksize == 3, CVAE_EINVAL, "conv_gemm: ksize %d", d->ksize);
    CVAE_REQUIRE(d->batch > 0 && d->height > 0 && d->width > 0, CVAE_EINVAL, "conv_gemm: empty shape");
    CVAE_REQUIRE(d->n_total % 16 == 0 && d->n_total > 0, CVAE_EINVAL, "conv_gemm: n_total %d", d->n_total);
    CVAE_REQUIRE(d->src && d->wpack && d->out, CVAE_EINVAL, "conv_gemm: null tensor");
    if (d->ktab == CVAE_KTAB_PAIR8)
        CVAE_REQUIRE(d->src_channels == 8 && d->ksize == 5, CVAE_EINVAL, "conv_gemm: PAIR8 needs 8 channels, 5x5");
    else
        CVAE_REQUIRE(d->src_channels % 16 == 0, CVAE_EINVAL, "conv_gemm: src_channels %d", d->src_channels);

    ConvArgs a{};
    a.B = d->batch; a.H = d->height; a.W = d->width; a.KW = d->ksize; a.pad = d->ksize / 2;
    a.PW = a.W + a.pad; a.IH = a.H + a.pad;
    a.planes = d->src_channels / 8;
    const int N = d->n_total < 128 ? d->n_total : 128;
    CVAE_REQUIRE(d->n_total % N == 0, CVAE_EINVAL, "conv_gemm: n_total %d not a multiple of %d", d->n_total, N);
    a.n_blocks = d->n_total / N;
    a.c_total = d->n_total;
    a.ksteps = cvae_conv_ksteps(d->ksize, d->src_channels, d->ktab);
    a.ktab_mode = d->ktab;
    a.halo = a.pad * a.PW + a.pad;
    a.wpack = (const __nv_bfloat16*)d->wpack; a.out = d->out;
    a.bias = d->bias; a.act = (const __nv_bfloat16*)d->act; a.stats = d->stats;
    a.fault = fault_flag();
    CVAE_REQUIRE(a.fault != nullptr, CVAE_ECUDA, "conv_gemm: fault flag unavailable");
    a.ps = PlaneSrc{a.B, a.H, a.W, a.pad, a.PW, a.IH, a.planes,
                    (d->loader == CVAE_LOAD_S2D) ? d->src_channels / 4 : d->src_channels, 0, d->src, d->src2};

    // Tiling policy.  Phases inside a CTA are sequential (load planes -> MMA -> epilogue), so overlap
    // comes from co-resident CTAs: aim for >= 2 CTAs per SM (<= 112 KB of shared memory and <= 256
    // TMEM columns each), with as many 128-pixel tiles per pass as fit -- every tile of a pass reuses
    // the same weight stage, so `tm` divides the L2 -> SM weight traffic.
    if (d->ktab == CVAE_KTAB_PAIR8) a.ksps = 13;
    else {
        a.ksps = d->src_channels / 16;                       // one tap
        while (a.ksps * N * 32 > 8 * 1024 && a.ksps % 2 == 0) a.ksps /= 2;
    }
    CVAE_REQUIRE(a.ksteps % a.ksps == 0, CVAE_EINVAL, "conv_gemm: internal stage split");
    const size_t stage_bytes = (size_t)a.ksps * N * 32;

    const long total_v = (long)a.B * a.IH * a.PW - (long)a.pad * a.PW;  // pixels from first to last valid row
    const int total_tiles = (int)((total_v + 127) / 128);
    auto smem_for = [&](int tm, int nstages) {
        const size_t L = (size_t)tm * 128 + 2 * a.halo + 8;
        return (((size_t)a.planes * L * 16 + 1023) & ~(size_t)1023) + nstages * stage_bytes + (size_t)a.ksteps * 8 + 64;
    };
    int tm = d->tm > 0 ? d->tm : (256 / N > 0 ? 256 / N : 1);
    if (tm > 8) tm = 8;
    if (tm > total_tiles) tm = total_tiles;
    a.nstages = 3;
    const size_t two_per_sm = 112 * 1024, one_per_sm = 200 * 1024;
    if (d->tm <= 0) {
        while (tm > 1 && smem_for(tm, 3) > two_per_sm) --tm;
        if (smem_for(tm, 3) > two_per_sm && smem_for(tm, 2) <= two_per_sm) a.nstages = 2;
    }
    while (tm > 1 && (smem_for(tm, a.nstages) > one_per_sm || tm * N > 512)) --tm;
    CVAE_REQUIRE(smem_for(tm, a.nstages) <= one_per_sm && tm * N <= 512, CVAE_EINVAL, "conv_gemm: shape does not fit shared memory");
    a.tm = tm;
    a.L = tm * 128 + 2 * a.halo + 8;
    a.plane_stride = a.L * 16;
    const size_t smem = smem_for(tm, a.nstages);
    a.num_chunks = (total_tiles + a.tm - 1) / a.tm;
    CVAE_REQUIRE((size_t)a.planes * a.plane_stride < (1u << 18), CVAE_EINVAL, "conv_gemm: planes exceed descriptor range");

#define CVAE_CASE(L_, E_, N_)                                                     \
    if (d->loader == (L_) && d->epilogue == (E_) && N == (N_))                    \
        return launch<L_, E_, N_>(a, smem, stream);
    CVAE_CASE(CVAE_LOAD_NCHW3, CVAE_EPI_STATS, 32)
    CVAE_CASE(CVAE_LOAD_NHWC, CVAE_EPI_STATS, 64)
    CVAE_CASE(CVAE_LOAD_NHWC, CVAE_EPI_STATS, 128)
    CVAE_CASE(CVAE_LOAD_NHWC, CVAE_EPI_BIAS_RELU, 128)
    CVAE_CASE(CVAE_LOAD_NHWC, CVAE_EPI_PHASE_BIAS_RELU, 128)
    CVAE_CASE(CVAE_LOAD_NHWC, CVAE_EPI_PHASE_BIAS_TANH, 16)
    CVAE_CASE(CVAE_LOAD_NHWC, CVAE_EPI_PLAIN, 32)
    CVAE_CASE(CVAE_LOAD_NHWC, CVAE_EPI_PLAIN, 64)
    CVAE_CASE(CVAE_LOAD_NHWC, CVAE_EPI_PLAIN, 128)
    CVAE_CASE(CVAE_LOAD_S2D, CVAE_EPI_MASK, 128)
    CVAE_CASE(CVAE_LOAD_S2D, CVAE_EPI_MASK, 64)
    CVAE_CASE(CVAE_LOAD_S2D, CVAE_EPI_MASK, 32)
    CVAE_CASE(CVAE_LOAD_S2D_NCHW3_DTANH, CVAE_EPI_MASK, 32)
#undef CVAE_CASE
    set_error("conv_gemm: no kernel for loader %d epilogue %d N %d", d->loader, d->epilogue, N);
    return CVAE_EINVAL;
}
