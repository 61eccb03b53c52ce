// Weight-gradient GEMM on tcgen05:  acc_g[m][n] = sum_v A[v + a_g][m] * B[v + b_g][n],  K = pixels.
//
// Both operands are halo planes ([channel/8][virtual pixel][8 ch], see planes.cuh) read MN-major:
// the 16 bytes of a pixel slot are 8 M/N elements and consecutive pixel slots are K.  One
// accumulator group per filter tap, all groups side by side in TMEM (up to 512 columns), each
// group being the same B tile read through a descriptor shifted by the tap offset.
//   normal : A = dY planes (M blocks = planes, SBO = plane stride), B = X planes.
//   shift  : A = ONE 8-channel plane whose 16 M blocks are 16 one-pixel shifts (SBO = 16 B); used
//            where one side has <= 8 channels (encoder conv 0 frames, decoder conv 4 gradient).
// The bias gradient rides along as a 16-column pseudo-group against a plane of ones.
// Split-K over pixel chunks: each CTA accumulates its chunks in TMEM and writes one fp32 partial;
// wgrad_fold_kernel sums the partials and scatters into the reference's OIHW layout (folding the
// 3x3 phase taps of the up-sample-folded decoder convs back onto the 5x5 filter).
//
// Replaces autograd's weight/bias gradients of nn.Conv2d at vae_nets.py:69,74,79,84,117-133.
#include "common.cuh"
#include "umma.cuh"
#include "planes.cuh"

namespace cvae {

static constexpr int kMaxGroups = 28;
static constexpr int kWgThreads = 160;  // warps 0-3 loaders + epilogue, warp 4 MMA issuer

struct WgGroup {
    int a_off;    // bytes, relative to the A region of the buffer (includes plane / m-block base)
    int b_off;    // bytes, relative to the B region of the buffer
    int n;        // UMMA N of this group
    int out_off;  // float offset of this group's [M_total][n] block inside one split's partial
};

struct WgradArgs {
    PlaneSrc pa, pb;
    int shift_a;
    int m_blocks;      // blockIdx.z
    int m_rows;        // real rows per M block that are stored (<= 128)
    int m_total;       // m_blocks * m_rows
    int groups_total, gpc;  // all groups (incl. the ones pseudo-group) / groups per CTA
    int kc, num_chunks, splits;
    int a_first, a_count, b_first, b_count;  // pixel-slot window per chunk, relative to chunk start
    int stride_a, stride_b;                  // plane strides (bytes)
    int a_region, buf_bytes;                 // bytes
    int ones_planes;                         // 2 when a ones tile follows the B planes
    int v_begin;
    int split_floats;                        // floats per split in `partial`
    WgGroup g[kMaxGroups];
    float* partial;
    int* fault;
};

template <int LA, int LB>
__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_kernel(const WgradArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_full[2], bar_empty[2], bar_acc;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int split = blockIdx.x, gset = blockIdx.y, mb = blockIdx.z;
    const int g0 = gset * a.gpc;
    const int g1 = min(g0 + a.gpc, a.groups_total);

    if (tid == 0) {
        mbar_init(&bar_full[0], 1);
        mbar_init(&bar_full[1], 1);
        mbar_init(&bar_empty[0], 1);
        mbar_init(&bar_empty[1], 1);
        mbar_init(&bar_acc, 1);
        mbar_fence_init();
    }
    int cols = 0;
    for (int g = g0; g < g1; ++g) cols += a.g[g].n;
    uint32_t ncols = 32;
    while (ncols < (uint32_t)cols) ncols <<= 1;
    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    // the ones tile never changes: write it once into both buffers
    if (a.ones_planes) {
        for (int b = 0; b < 2; ++b) {
            uint8_t* bp = smem + (size_t)b * a.buf_bytes + a.a_region + (size_t)a.pb.planes * a.stride_b;
            for (int q = 0; q < a.ones_planes; ++q)
                fill_ones_plane(bp + (size_t)q * a.stride_b, a.b_count, tid, kWgThreads);
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    const int my_chunks = (a.num_chunks - split + a.splits - 1) / a.splits;  // chunks split, split+S, ...

    if (warp < 4) {
        // ------------------------------ loaders ------------------------------------------------
        bool alive = true;
        for (int i = 0; i < my_chunks; ++i) {
            const int buf = i & 1;
            const int c0 = a.v_begin + (split + i * a.splits) * a.kc;
            if (alive) alive = mbar_wait(&bar_empty[buf], ((i >> 1) & 1) ^ 1, a.fault);
            uint8_t* A = smem + (size_t)buf * a.buf_bytes;
            uint8_t* Bp = A + a.a_region;
            fill_planes<LA>(a.pa, A, a.stride_a, c0 + a.a_first, a.a_count, tid, 128);
            fill_planes<LB>(a.pb, Bp, a.stride_b, c0 + a.b_first, a.b_count, tid, 128);
            fence_proxy_async();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 0) mbar_arrive(&bar_full[buf]);
        }
        // ------------------------------ epilogue -----------------------------------------------
        mbar_wait(&bar_acc, 0, a.fault);
        tc_fence_after();
        const int row = warp * 32 + lane;
        float* base = a.partial + (size_t)split * a.split_floats;
        uint32_t col = 0;
        for (int g = g0; g < g1; ++g) {
            const int n = a.g[g].n;
            float* o = base + a.g[g].out_off + ((size_t)mb * a.m_rows + row) * n;
            for (int c = 0; c < n; c += 16) {
                uint32_t raw[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + col + c, raw);
                tmem_wait_ld();
                if (row < a.m_rows) {
                    float4* o4 = reinterpret_cast<float4*>(o + c);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        o4[q] = make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]),
                                            __uint_as_float(raw[4 * q + 2]), __uint_as_float(raw[4 * q + 3]));
                }
            }
            col += n;
        }
    } else {
        // ------------------------------ MMA issuer ---------------------------------------------
        if (elect_one()) {   // elect.sync, not `lane == 0`: the compiler then emits UTCHMMA without an ELECT/BRA.U.ANY loop
            bool alive = true;
            const uint32_t smem_base = smem_u32(smem);
            const uint32_t sbo_a = a.shift_a ? 16u : (uint32_t)a.stride_a;
            const uint32_t a_mb = a.shift_a ? 0u : (uint32_t)mb * 16u * a.stride_a;
            for (int i = 0; i < my_chunks && alive; ++i) {
                const int buf = i & 1;
                alive = mbar_wait(&bar_full[buf], (i >> 1) & 1, a.fault);
                tc_fence_after();
                const uint32_t A = smem_base + (uint32_t)buf * a.buf_bytes;
                const uint32_t Bp = A + a.a_region;
                uint32_t col = 0;
                for (int g = g0; g < g1; ++g) {
                    const uint32_t idesc = umma_idesc_bf16(a.g[g].n, kMajorMN, kMajorMN);
                    const uint32_t as = A + a_mb + a.g[g].a_off, bs = Bp + a.g[g].b_off;
                    for (int k = 0; k < a.kc / 16; ++k) {
                        const uint64_t da = smem_desc(as + k * 256u, 128u, sbo_a);
                        const uint64_t db = smem_desc(bs + k * 256u, 128u, (uint32_t)a.stride_b);
                        umma_bf16(tmem_base + col, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
                    }
                    col += a.g[g].n;
                }
                umma_commit(&bar_empty[buf]);
            }
            umma_commit(&bar_acc);
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, ncols);
}

// --------------------------------------------------------------------------------------------
// fold: partials -> OIHW fp32 gradient (+ bias gradient)
// --------------------------------------------------------------------------------------------
struct FoldArgs {
    int kind;          // cvae_wgrad_kind
    int cout, cin;     // of the reference conv weight [cout][cin][5][5]
    int splits, split_floats;
    int m_total;       // rows per group in the partial
    int n;             // columns per normal group
    int bias_off;      // float offset of the ones pseudo-group ([m_total][16]) inside a split, -1: none
    const float* partial;
    float* dw;         // [cout][cin][5][5]
    float* dbias;      // [cout]
};

// low-res tap (t in 0..2) that 5x5 tap k (0..4) folds onto for output phase a
__device__ __forceinline__ int phase_tap(int a, int k) {
    const int d = k - 2;
    if (a == 0) return d <= -1 ? 0 : (d <= 1 ? 1 : 2);
    return d <= -2 ? 0 : (d <= 0 ? 1 : 2);
}

// One split's contribution to dw[co][ci][ky][kx] (p = that split's partial).
__device__ __forceinline__ float fold_term(const FoldArgs& f, const float* __restrict__ p, int co, int ci, int ky, int kx) {
    if (f.kind == CVAE_WGRAD_5X5) {           // group = tap, row = co, col = ci
        return __ldg(p + ((size_t)(ky * 5 + kx) * f.m_total + co) * f.n + ci);
    } else if (f.kind == CVAE_WGRAD_PHASE) {  // group = 3x3 tap, row = (a,b,co), col = ci
        float acc = 0.f;
#pragma unroll
        for (int ab = 0; ab < 4; ++ab) {
            const int t = phase_tap(ab >> 1, ky) * 3 + phase_tap(ab & 1, kx);
            acc += __ldg(p + ((size_t)t * f.m_total + ab * f.cout + co) * f.n + ci);
        }
        return acc;
    } else if (f.kind == CVAE_WGRAD_SHIFT_FRAMES) {  // group = ky, row = (kx, ch), col = co
        return __ldg(p + ((size_t)ky * f.m_total + kx * 8 + ci) * f.n + co);
    } else {  // CVAE_WGRAD_SHIFT_PHASE12: group = (plane p, ty), row = (j, e), col = ci; tx = 1 - j
        float acc = 0.f;
#pragma unroll
        for (int ab = 0; ab < 4; ++ab) {
            const int ch = ab * 3 + co, pl = ch >> 3, e = ch & 7;
            const int ty = phase_tap(ab >> 1, ky), tx = phase_tap(ab & 1, kx);
            const int j = 2 - tx;  // tx index 0..2 <-> offset tx-1 = 1 - j
            acc += __ldg(p + ((size_t)(pl * 3 + ty) * f.m_total + j * 8 + e) * f.n + ci);
        }
        return acc;
    }
}
__device__ __forceinline__ float fold_bias_term(const FoldArgs& f, const float* __restrict__ p, int co) {
    if (f.kind == CVAE_WGRAD_5X5) return __ldg(p + f.bias_off + (size_t)co * 16);
    if (f.kind == CVAE_WGRAD_PHASE) {
        float acc = 0.f;
        for (int ab = 0; ab < 4; ++ab) acc += __ldg(p + f.bias_off + (size_t)(ab * f.cout + co) * 16);
        return acc;
    }
    // frames kind: ones live in channel 3 of the frame plane: row (kx = 2, ch 3) of group ky = 2
    if (f.kind == CVAE_WGRAD_SHIFT_FRAMES) return __ldg(p + ((size_t)2 * f.m_total + 2 * 8 + 3) * f.n + co);
    float acc = 0.f;
    for (int ab = 0; ab < 4; ++ab) {
        const int ch = ab * 3 + co, pl = ch >> 3, e = ch & 7;
        // ones pseudo-groups: one per plane, rows (j, e); any j sums the same pixels
        acc += __ldg(p + f.bias_off + ((size_t)pl * f.m_total + 0 * 8 + e) * 16);
    }
    return acc;
}

// Fold: block = 32 consecutive outputs x 8 split lanes.  Outputs are enumerated with the partial's
// fastest index innermost, so every warp reads one contiguous 128-byte row segment per split; the
// split loop is spread over the 8 warps and carries two independent accumulators, then the 8 lane
// sums are added in a fixed order (bit-reproducible).  The OIHW write is a strided 4-byte scatter
// (10 MB per step in total).
static constexpr int kFoldLanes = 8;
__global__ void __launch_bounds__(256) wgrad_fold_kernel(const FoldArgs f) {
    __shared__ float red[kFoldLanes][33];
    const int total_w = 25 * f.cout * f.cin;
    const int total = total_w + (f.dbias ? f.cout : 0);
    const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int idx = blockIdx.x * 32 + o;
    int co = 0, ci = 0, ky = 0, kx = 0;
    const bool is_w = idx < total_w, is_b = !is_w && idx < total;
    if (is_w) {
        int r = idx;
        if (f.kind == CVAE_WGRAD_5X5 || f.kind == CVAE_WGRAD_PHASE) {            // (tap, co, ci)
            ci = r % f.cin; r /= f.cin; co = r % f.cout; r /= f.cout; ky = r / 5; kx = r - ky * 5;
        } else if (f.kind == CVAE_WGRAD_SHIFT_FRAMES) {                          // (ky, kx, ci, co)
            co = r % f.cout; r /= f.cout; ci = r % f.cin; r /= f.cin; kx = r % 5; ky = r / 5;
        } else {                                                                 // (co, ky, kx, ci)
            ci = r % f.cin; r /= f.cin; kx = r % 5; r /= 5; ky = r % 5; co = r / 5;
        }
    } else if (is_b) {
        co = idx - total_w;
    }
    float a0 = 0.f, a1 = 0.f;
    const size_t ss = (size_t)f.split_floats;
    if (is_w) {
        int s = sl;
        for (; s + kFoldLanes < f.splits; s += 2 * kFoldLanes) {
            a0 += fold_term(f, f.partial + (size_t)s * ss, co, ci, ky, kx);
            a1 += fold_term(f, f.partial + (size_t)(s + kFoldLanes) * ss, co, ci, ky, kx);
        }
        if (s < f.splits) a0 += fold_term(f, f.partial + (size_t)s * ss, co, ci, ky, kx);
    } else if (is_b) {
        for (int s = sl; s < f.splits; s += kFoldLanes) a0 += fold_bias_term(f, f.partial + (size_t)s * ss, co);
    }
    red[sl][o] = a0 + a1;
    __syncthreads();
    if (sl == 0 && (is_w || is_b)) {
        float acc = 0.f;
#pragma unroll
        for (int l = 0; l < kFoldLanes; ++l) acc += red[l][o];
        if (is_w) f.dw[((size_t)co * f.cin + ci) * 25 + ky * 5 + kx] = acc;
        else f.dbias[co] = acc;
    }
}

template <int LA, int LB>
static int launch_wgrad(const WgradArgs& a, size_t smem, dim3 grid, cudaStream_t stream) {
    auto kern = conv_wgrad_kernel<LA, LB>;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        CVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    kern<<<grid, kWgThreads, smem, stream>>>(a);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

}  // namespace cvae

using namespace cvae;

extern "C" int64_t cvae_conv_wgrad_workspace_bytes(const cvae_wgrad_desc* d) {
    if (!d) return -1;
    // generous upper bound: splits <= 2 * SM count, every group m_total x (n + 16) floats
    int m_total, n, groups;
    switch (d->kind) {
        case CVAE_WGRAD_5X5: m_total = d->cout; n = d->cin; groups = 25; break;
        case CVAE_WGRAD_PHASE: m_total = 4 * d->cout; n = d->cin; groups = 9; break;
        case CVAE_WGRAD_SHIFT_FRAMES: m_total = 128; n = d->cout; groups = 5; break;
        case CVAE_WGRAD_SHIFT_PHASE12: m_total = 128; n = d->cin; groups = 8; break;
        default: return -1;
    }
    if (m_total < 128) m_total = 128;
    return (int64_t)2 * 160 * ((int64_t)groups * m_total * n + (int64_t)2 * m_total * 16) * 4;
}

extern "C" int cvae_conv_wgrad(const cvae_wgrad_desc* d, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(d != nullptr, CVAE_EINVAL, "conv_wgrad: null descriptor");
    CVAE_REQUIRE(d->x && d->dy && d->dw && d->workspace, CVAE_EINVAL, "conv_wgrad: null tensor");
    CVAE_REQUIRE(d->batch > 0, CVAE_EINVAL, "conv_wgrad: empty batch");

    WgradArgs a{};
    FoldArgs f{};
    const int H = d->height, W = d->width;  // grid the GEMM's K pixels live on
    const int pad = (d->kind == CVAE_WGRAD_5X5 || d->kind == CVAE_WGRAD_SHIFT_FRAMES) ? 2 : 1;
    const int PW = W + pad, IH = H + pad, halo = pad * PW + pad;
    // The PHASE12 variant shifts the A operand (dY) by up to 15 pixels relative to the summation
    // index, so its first chunk starts 16 slots early (pure padding) to cover the first real pixels.
    a.v_begin = pad * PW - (d->kind == CVAE_WGRAD_SHIFT_PHASE12 ? 16 : 0);
    const long total_v = (long)d->batch * IH * PW - a.v_begin;
    int n = 0, ngroups = 0, a_planes = 0, b_planes = 0;
    a.shift_a = 0;
    a.ones_planes = 0;
    int la = CVAE_LOAD_NHWC, lb = CVAE_LOAD_NHWC;

    if (d->kind == CVAE_WGRAD_5X5 || d->kind == CVAE_WGRAD_PHASE) {
        const int mtot = (d->kind == CVAE_WGRAD_5X5) ? d->cout : 4 * d->cout;
        CVAE_REQUIRE(mtot % 8 == 0 && d->cin % 16 == 0 && d->cin <= 256, CVAE_EINVAL,
                     "conv_wgrad: unsupported channels %d/%d", d->cout, d->cin);
        a_planes = mtot / 8; b_planes = d->cin / 8; n = d->cin;
        ngroups = (d->kind == CVAE_WGRAD_5X5) ? 25 : 9;
        a.m_blocks = (mtot + 127) / 128;
        a.m_rows = mtot < 128 ? mtot : 128;
        la = (d->kind == CVAE_WGRAD_5X5) ? CVAE_LOAD_NHWC : CVAE_LOAD_S2D;
        a.pa = PlaneSrc{d->batch, H, W, pad, PW, IH, a_planes, (d->kind == CVAE_WGRAD_5X5) ? mtot : d->cout, 0, d->dy, nullptr};
        a.pb = PlaneSrc{d->batch, H, W, pad, PW, IH, b_planes, d->cin, 0, d->x, nullptr};
        a.ones_planes = 2;
        a.a_first = 0; a.b_first = -halo;
    } else if (d->kind == CVAE_WGRAD_SHIFT_FRAMES) {
        // A = frames (3 + ones channel, one plane, 16 shifts), B = dY (cout channels)
        CVAE_REQUIRE(d->cin == 3 && d->cout % 16 == 0 && d->cout <= 256, CVAE_EINVAL, "conv_wgrad: frames kind");
        a.shift_a = 1; a_planes = 1; b_planes = d->cout / 8; n = d->cout; ngroups = 5;
        a.m_blocks = 1; a.m_rows = 128;
        la = CVAE_LOAD_NCHW3;
        a.pa = PlaneSrc{d->batch, H, W, pad, PW, IH, 1, 3, 1, d->x, nullptr};
        a.pb = PlaneSrc{d->batch, H, W, pad, PW, IH, b_planes, d->cout, 0, d->dy, nullptr};
        a.a_first = -2 * PW - 2; a.b_first = 0;
    } else if (d->kind == CVAE_WGRAD_SHIFT_PHASE12) {
        // A = d_recon*(1-recon^2) space-to-depth, 12(+4) channels in two planes, 16 shifts each;
        // B = X (cin channels) + ones
        CVAE_REQUIRE(d->cout == 3 && d->cin % 16 == 0 && d->cin <= 256 && d->dy2, CVAE_EINVAL, "conv_wgrad: phase12 kind");
        a.shift_a = 1; a_planes = 2; b_planes = d->cin / 8; n = d->cin; ngroups = 6;
        a.m_blocks = 1; a.m_rows = 128;
        la = CVAE_LOAD_S2D_NCHW3_DTANH;
        a.pa = PlaneSrc{d->batch, H, W, pad, PW, IH, 2, 3, 0, d->dy, d->dy2};
        a.pb = PlaneSrc{d->batch, H, W, pad, PW, IH, b_planes, d->cin, 0, d->x, nullptr};
        a.ones_planes = 2;
        a.a_first = 0; a.b_first = -halo;
    } else {
        set_error("conv_wgrad: unknown kind %d", d->kind);
        return CVAE_EINVAL;
    }
    a.m_total = a.m_blocks * a.m_rows;

    // chunk size: biggest of 512/256/128 pixels whose double buffer fits
    const size_t cap = 200 * 1024;
    int kc = 512;
    for (;; kc /= 2) {
        CVAE_REQUIRE(kc >= 64, CVAE_EINVAL, "conv_wgrad: shape does not fit shared memory");
        a.kc = kc;
        if (a.shift_a) {
            a.a_count = kc + 16 + (d->kind == CVAE_WGRAD_SHIFT_FRAMES ? 4 * PW : 0);
            a.b_count = kc + (d->kind == CVAE_WGRAD_SHIFT_FRAMES ? 0 : 2 * halo + 2);
        } else {
            a.a_count = kc;
            a.b_count = kc + 2 * halo;
        }
        a.stride_a = a.a_count * 16;
        a.stride_b = a.b_count * 16;
        const int a_planes_alloc = a.shift_a ? a_planes : (a_planes < 16 ? 16 : a_planes);
        a.a_region = (a_planes_alloc * a.stride_a + 1023) & ~1023;
        a.buf_bytes = (a.a_region + (b_planes + a.ones_planes) * a.stride_b + 1023) & ~1023;
        if ((size_t)2 * a.buf_bytes <= cap && a.buf_bytes < (1 << 17)) break;
    }
    a.num_chunks = (int)((total_v + kc - 1) / kc);

    // groups
    int gi = 0, out_off = 0;
    if (d->kind == CVAE_WGRAD_5X5 || d->kind == CVAE_WGRAD_PHASE) {
        const int K = (d->kind == CVAE_WGRAD_5X5) ? 5 : 3;
        for (int t = 0; t < K * K; ++t) {
            const int dy = t / K - pad, dx = t % K - pad;
            a.g[gi++] = WgGroup{0, (halo + dy * PW + dx) * 16, n, out_off};
            out_off += a.m_total * n;
        }
        f.bias_off = out_off;
        a.g[gi++] = WgGroup{0, b_planes * a.stride_b + halo * 16, 16, out_off};
        out_off += a.m_total * 16;
    } else if (d->kind == CVAE_WGRAD_SHIFT_FRAMES) {
        for (int ky = 0; ky < 5; ++ky) {
            a.g[gi++] = WgGroup{((ky - 2) * PW - 2 - a.a_first) * 16, 0, n, out_off};
            out_off += a.m_total * n;
        }
        f.bias_off = -1;
    } else {
        for (int p = 0; p < 2; ++p)
            for (int ty = -1; ty <= 1; ++ty) {
                a.g[gi++] = WgGroup{p * a.stride_a, (ty * PW + 1 - a.b_first) * 16, n, out_off};
                out_off += a.m_total * n;
            }
        f.bias_off = out_off;
        for (int p = 0; p < 2; ++p) {
            a.g[gi++] = WgGroup{p * a.stride_a, b_planes * a.stride_b, 16, out_off};
            out_off += a.m_total * 16;
        }
    }
    a.groups_total = gi;
    a.split_floats = out_off;
    // groups per CTA: fill the 512 TMEM columns
    a.gpc = 512 / n;
    if (a.gpc < 1) a.gpc = 1;
    {   // keep the column total of every set <= 512 including a trailing 16-column pseudo-group
        while (a.gpc > 1) {
            bool ok = true;
            for (int g0 = 0; g0 < a.groups_total && ok; g0 += a.gpc) {
                int cols = 0;
                for (int g = g0; g < g0 + a.gpc && g < a.groups_total; ++g) cols += a.g[g].n;
                ok = cols <= 512;
            }
            if (ok) break;
            --a.gpc;
        }
    }
    const int gsets = (a.groups_total + a.gpc - 1) / a.gpc;
    int splits = d->splits > 0 ? d->splits : (sm_count() / (gsets * a.m_blocks));
    if (splits < 1) splits = 1;
    if (splits > a.num_chunks) splits = a.num_chunks;
    a.splits = splits;
    a.partial = (float*)d->workspace;
    a.fault = fault_flag();
    CVAE_REQUIRE(a.fault != nullptr, CVAE_ECUDA, "conv_wgrad: fault flag unavailable");
    CVAE_REQUIRE((int64_t)splits * a.split_floats * 4 <= cvae_conv_wgrad_workspace_bytes(d), CVAE_EINVAL,
                 "conv_wgrad: workspace too small");

    const size_t smem = (size_t)2 * a.buf_bytes;
    dim3 grid(splits, gsets, a.m_blocks);
    int rc;
    if (la == CVAE_LOAD_NHWC) rc = launch_wgrad<CVAE_LOAD_NHWC, CVAE_LOAD_NHWC>(a, smem, grid, stream);
    else if (la == CVAE_LOAD_S2D) rc = launch_wgrad<CVAE_LOAD_S2D, CVAE_LOAD_NHWC>(a, smem, grid, stream);
    else if (la == CVAE_LOAD_NCHW3) rc = launch_wgrad<CVAE_LOAD_NCHW3, CVAE_LOAD_NHWC>(a, smem, grid, stream);
    else rc = launch_wgrad<CVAE_LOAD_S2D_NCHW3_DTANH, CVAE_LOAD_NHWC>(a, smem, grid, stream);
    if (rc != CVAE_OK) return rc;

    f.kind = d->kind; f.cout = d->cout; f.cin = d->cin; f.splits = splits; f.split_floats = a.split_floats;
    f.m_total = a.m_total; f.n = n; f.partial = a.partial; f.dw = (float*)d->dw; f.dbias = (float*)d->dbias;
    {
        const int total = d->cout * d->cin * 25 + d->cout;
        wgrad_fold_kernel<<<(total + 31) / 32, 256, 0, stream>>>(f);
    }
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}
