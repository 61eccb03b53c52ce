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
// warps 0 .. kWgLoadWarps-1 load planes (the address arithmetic of a 16-byte-granular gather is latency bound with
// one warp per scheduler, so there are several per scheduler); warps 0-3 also run the epilogue; the last warp issues MMAs
static constexpr int kWgLoadWarps = 12;
static constexpr int kWgLoaders = kWgLoadWarps * 32;
static constexpr int kWgThreads = kWgLoaders + 32;

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
    int a_planes_cta;                        // A planes one CTA loads (its M block only)
    FastDiv dPW, dIH;
    int v_begin;
    int split_floats;                        // floats per split in `partial`
    WgGroup g[kMaxGroups];
    float* partial;
    int* fault;
    unsigned long long* dbg;   // optional per-CTA cycle counters [ctas][8] (cvae_wgrad_debug_counters)
};

// kWgKS consecutive K steps of one accumulator group, fully unrolled: every descriptor is a base plus an
// immediate, so the MMAs issue back to back from the uniform datapath (tools/umma_rate.cu: a descriptor
// rebuilt from vector registers costs >100 cycles per MMA, three times the N = 32 MMA itself).
static constexpr int kWgKS = 4;
__device__ __forceinline__ void wg_issue(uint32_t tcol, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate_first) {
#pragma unroll
    for (int j = 0; j < kWgKS; ++j)
        umma_bf16(tcol, ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)(j * 16)),
                  ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)(j * 16)), idesc, j == 0 ? accumulate_first : 1u);
}

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

    if (warp < kWgLoadWarps) {
        // ------------------------------ loaders ------------------------------------------------
        // bf16 sources go global -> shared with cp.async (zero-fill for the padding), all copies of a chunk in
        // flight at once; fp32 sources (frames, d_recon) are converted by the threads themselves.  The fill of
        // chunk i overlaps the MMAs of chunk i-1 (two buffers).
        constexpr bool kAsyncA = (LA == CVAE_LOAD_NHWC || LA == CVAE_LOAD_S2D);
        bool alive = true;
        long long t_we = 0, t_issue = 0, t_land = 0, tq = 0;
        const bool prof = a.dbg != nullptr;
        for (int i = 0; i < my_chunks; ++i) {
            const int buf = i & 1;
            const int c0 = a.v_begin + (split + i * a.splits) * a.kc;
            if (prof) tq = clock64();
            if (alive) alive = mbar_wait(&bar_empty[buf], ((i >> 1) & 1) ^ 1, a.fault);
            if (prof) { const long long t = clock64(); t_we += t - tq; tq = t; }
            uint8_t* A = smem + (size_t)buf * a.buf_bytes;
            uint8_t* Bp = A + a.a_region;
            if constexpr (kAsyncA)
                fill_planes_async<LA>(a.pa, a.dPW, a.dIH, A, a.stride_a, c0 + a.a_first, a.a_count, mb * 16, a.a_planes_cta, tid, kWgLoaders);
            fill_planes_async<LB>(a.pb, a.dPW, a.dIH, Bp, a.stride_b, c0 + a.b_first, a.b_count, 0, a.pb.planes, tid, kWgLoaders);
            if constexpr (!kAsyncA) fill_planes<LA>(a.pa, A, a.stride_a, c0 + a.a_first, a.a_count, tid, kWgLoaders);
            if (prof) { const long long t = clock64(); t_issue += t - tq; tq = t; }
            cp_async_wait_all();
            fence_proxy_async();
            asm volatile("bar.sync 1, %0;" ::"n"(kWgLoaders) : "memory");
            if (tid == 0) mbar_arrive(&bar_full[buf]);
            if (prof) t_land += clock64() - tq;
        }
        if (prof && tid == 0) {
            unsigned long long* o = a.dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8;
            o[3] = t_we; o[4] = t_issue; o[5] = t_land; o[6] = my_chunks;
        }
        long long t_ep = prof ? clock64() : 0;
        // ------------------------------ epilogue -----------------------------------------------
        mbar_wait(&bar_acc, 0, a.fault);
        tc_fence_after();
        // warp w reads TMEM lanes 32 (w % 4) .. +31 (hardware rule); the accumulator groups are dealt round-robin
        // to the kWgLoadWarps / 4 warps that share a lane quarter
        const int row = (warp & 3) * 32 + lane;
        float* base = a.partial + (size_t)split * a.split_floats;
        uint32_t col = 0;
        for (int g = g0; g < g1; ++g) {
            const int n = a.g[g].n;
            if ((g - g0) % (kWgLoadWarps / 4) != (warp >> 2)) { col += n; continue; }
            float* o = base + a.g[g].out_off + ((size_t)mb * a.m_rows + row) * n;
            for (int c = 0; c < n; c += 16) {
                uint32_t raw[16];
                tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + col + c, raw);
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
        if (prof && tid == 0) {
            unsigned long long* o = a.dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8;
            o[7] = (unsigned long long)(clock64() - t_ep);   // wait for the last MMAs + TMEM -> global
        }
    } else {
        // ------------------------------ MMA issuer ---------------------------------------------
        if (elect_one()) {   // elect.sync, not `lane == 0`: the compiler then emits UTCHMMA without an ELECT/BRA.U.ANY loop
            bool alive = true;
            const uint32_t smem_base = smem_u32(smem);
            // descriptors: low word = (address >> 4) | LBO (128 B) << 16, high word = SBO | version; the K loop
            // advances both start addresses by 256 B = 16 descriptor units, issued in unrolled groups of kWgKS
            const uint32_t lbo = (128u >> 4) << 16;
            const uint32_t a_hi = (((a.shift_a ? 16u : (uint32_t)a.stride_a) >> 4) & 0x3FFFu) | (1u << 14);
            const uint32_t b_hi = (((uint32_t)a.stride_b >> 4) & 0x3FFFu) | (1u << 14);
            const int kgroups = a.kc / (16 * kWgKS);
            const bool prof = a.dbg != nullptr;
            long long t_wf = 0, tq = 0;
            const long long t0 = prof ? clock64() : 0;
            for (int i = 0; i < my_chunks && alive; ++i) {
                const int buf = i & 1;
                if (prof) tq = clock64();
                alive = mbar_wait(&bar_full[buf], (i >> 1) & 1, a.fault);
                if (prof) t_wf += clock64() - tq;
                tc_fence_after();
                const uint32_t A16 = ((smem_base + (uint32_t)buf * a.buf_bytes) & 0x3FFFFu) >> 4;
                const uint32_t B16 = A16 + ((uint32_t)a.a_region >> 4);
                uint32_t col = tmem_base;
                for (int g = g0; g < g1; ++g) {
                    const uint32_t idesc = umma_idesc_bf16(a.g[g].n, kMajorMN, kMajorMN);
                    uint32_t a_lo = (A16 + ((uint32_t)a.g[g].a_off >> 4)) | lbo;
                    uint32_t b_lo = (B16 + ((uint32_t)a.g[g].b_off >> 4)) | lbo;
                    wg_issue(col, a_lo, b_lo, a_hi, b_hi, idesc, i > 0 ? 1u : 0u);
                    for (int kg = 1; kg < kgroups; ++kg) {
                        a_lo += 16u * kWgKS;
                        b_lo += 16u * kWgKS;
                        wg_issue(col, a_lo, b_lo, a_hi, b_hi, idesc, 1u);
                    }
                    col += (uint32_t)a.g[g].n;
                }
                umma_commit(&bar_empty[buf]);
            }
            umma_commit(&bar_acc);
            if (prof) {
                unsigned long long* o = a.dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8;
                o[0] = (unsigned long long)(clock64() - t0); o[1] = t_wf;
            }
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
    int dbg_linear;
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

// Fold: block = 256 / L consecutive outputs x L split lanes (L = 8 for many splits, fewer for few).  Outputs are enumerated with the partial's
// fastest index innermost, so every warp reads one contiguous 128-byte row segment per split; the
// split loop is spread over the 8 warps and carries two independent accumulators, then the 8 lane
// sums are added in a fixed order (bit-reproducible).  The OIHW write is a strided 4-byte scatter
// (10 MB per step in total).
template <int kFoldLanes>
__global__ void __launch_bounds__(256) wgrad_fold_kernel(const FoldArgs f) {
    constexpr int kOut = 256 / kFoldLanes;   // outputs per block
    __shared__ float red[kFoldLanes][kOut + 1];
    const int total_w = 25 * f.cout * f.cin;
    const int total = total_w + (f.dbias ? f.cout : 0);
    const int o = threadIdx.x % kOut, sl = threadIdx.x / kOut;
    const int idx = blockIdx.x * kOut + o;
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
        if (is_w) f.dw[f.dbg_linear ? (size_t)idx : ((size_t)co * f.cin + ci) * 25 + ky * 5 + kx] = acc;
        else f.dbias[co] = acc;
    }
}

// Vectorised fold for the 5X5 / PHASE kinds (ci fastest in the partial, cin % 4 == 0): a thread owns 4
// consecutive ci of one (tap, co) and reads float4s, so a warp covers a 512-byte row segment per split.
// Blocks past the weight range compute the bias gradient (scalar path).
template <int kFoldLanes>
__global__ void __launch_bounds__(256) wgrad_fold_rows_kernel(const FoldArgs f, int weight_blocks) {
    constexpr int kOut = 256 / kFoldLanes;   // float4 outputs per block
    __shared__ float4 red[kFoldLanes][kOut];
    const int o = threadIdx.x % kOut, sl = threadIdx.x / kOut;
    const size_t ss = (size_t)f.split_floats;
    if ((int)blockIdx.x >= weight_blocks) {   // bias: 256 / L outputs per block, scalar
        const int co = ((int)blockIdx.x - weight_blocks) * kOut + o;
        float acc = 0.f;
        if (co < f.cout)
            for (int s = sl; s < f.splits; s += kFoldLanes) acc += fold_bias_term(f, f.partial + (size_t)s * ss, co);
        red[sl][o].x = acc;
        __syncthreads();
        if (sl == 0 && co < f.cout) {
            float t = 0.f;
#pragma unroll
            for (int l = 0; l < kFoldLanes; ++l) t += red[l][o].x;
            f.dbias[co] = t;
        }
        return;
    }
    const int cin4 = f.cin >> 2;
    const int total4 = 25 * f.cout * cin4;
    const int idx = blockIdx.x * kOut + o;
    const bool live = idx < total4;
    int ci = 0, co = 0, ky = 0, kx = 0;
    if (live) {
        int r = idx;
        ci = (r % cin4) * 4; r /= cin4; co = r % f.cout; r /= f.cout; ky = r / 5; kx = r - ky * 5;
    }
    // offsets of the (up to 4) partial rows that fold onto this output
    size_t off[4];
    int nsrc = 1;
    if (f.kind == CVAE_WGRAD_5X5) {
        off[0] = ((size_t)(ky * 5 + kx) * f.m_total + co) * f.n + ci;
    } else {
        nsrc = 4;
#pragma unroll
        for (int ab = 0; ab < 4; ++ab) {
            const int t = phase_tap(ab >> 1, ky) * 3 + phase_tap(ab & 1, kx);
            off[ab] = ((size_t)t * f.m_total + ab * f.cout + co) * f.n + ci;
        }
    }
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    auto add = [](float4& d, const float4 v) { d.x += v.x; d.y += v.y; d.z += v.z; d.w += v.w; };
    if (live) {
        for (int q = 0; q < nsrc; ++q) {
            const float* p = f.partial + off[q];
            int s = sl;
            for (; s + kFoldLanes < f.splits; s += 2 * kFoldLanes) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(p + (size_t)s * ss));
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(s + kFoldLanes) * ss));
                add(a0, v0);
                add(a1, v1);
            }
            if (s < f.splits) add(a0, __ldg(reinterpret_cast<const float4*>(p + (size_t)s * ss)));
        }
    }
    add(a0, a1);
    red[sl][o] = a0;
    __syncthreads();
    if (sl == 0 && live) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int l = 0; l < kFoldLanes; ++l) add(t, red[l][o]);
        float* d = f.dw + ((size_t)co * f.cin + ci) * 25 + ky * 5 + kx;
        d[0] = t.x; d[25] = t.y; d[50] = t.z; d[75] = t.w;
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

static unsigned long long* g_wg_dbg = nullptr;
// Profiling aid: when set, every weight-gradient CTA writes 8 cycle counters to buf[cta * 8 ..]: MMA thread
// total / wait-for-data; loaders wait-for-buffer / issue / landing wait / chunks; epilogue.
extern "C" void cvae_wgrad_debug_counters(void* device_buf) { g_wg_dbg = (unsigned long long*)device_buf; }

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
        a.ones_planes = d->dbias ? 2 : 0;   // no bias gradient wanted (conv in front of BatchNorm): skip the ones group
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

    a.a_planes_cta = a.shift_a ? a_planes : (a_planes < 16 ? a_planes : 16);
    a.dPW = make_fastdiv(PW);
    a.dIH = make_fastdiv(IH);

    // groups per CTA: fill the 512 TMEM columns (every group is n columns wide; the bias pseudo-groups 16)
    const int bias_groups = (d->kind == CVAE_WGRAD_SHIFT_FRAMES) ? 0 : (d->kind == CVAE_WGRAD_SHIFT_PHASE12 ? 2 : (d->dbias ? 1 : 0));
    ngroups = (d->kind == CVAE_WGRAD_5X5) ? 25 : (d->kind == CVAE_WGRAD_PHASE ? 9 : (d->kind == CVAE_WGRAD_SHIFT_FRAMES ? 5 : 6));
    a.groups_total = ngroups + bias_groups;
    a.gpc = 512 / n;
    if (a.gpc < 1) a.gpc = 1;
    while (a.gpc > 1) {   // keep the column total of every set <= 512 including trailing 16-column pseudo-groups
        bool ok = true;
        for (int g0 = 0; g0 < a.groups_total && ok; g0 += a.gpc) {
            int cols = 0;
            for (int g = g0; g < g0 + a.gpc && g < a.groups_total; ++g) cols += (g < ngroups) ? n : 16;
            ok = cols <= 512;
        }
        if (ok) break;
        --a.gpc;
    }
    const int gsets = (a.groups_total + a.gpc - 1) / a.gpc;
    a.gpc = (a.groups_total + gsets - 1) / gsets;   // same number of sets, balanced
    int splits = d->splits > 0 ? d->splits : (sm_count() / (gsets * a.m_blocks));
    if (splits < 1) splits = 1;

    // chunk size: the largest multiple of 16 pixels whose double buffer fits, but small enough that every
    // CTA still gets ~4 chunks to pipeline
    const size_t cap = 216 * 1024;
    long per_cta = (total_v + splits - 1) / splits;
    const int kq = 16 * kWgKS;   // the MMA issuer unrolls kWgKS K steps
    int kc_want = (int)(((per_cta + 3) / 4 + kq - 1) / kq * kq);
    if (kc_want < 64) kc_want = 64;
    if (kc_want > 512) kc_want = 512;
    if (getenv("CVAE_WG_KC")) kc_want = atoi(getenv("CVAE_WG_KC")) / kq * kq;
    int kc = kc_want;
    for (;; kc -= kq) {
        CVAE_REQUIRE(kc >= kq, CVAE_EINVAL, "conv_wgrad: shape does not fit shared memory");
        a.kc = kc;
        if (a.shift_a) {
            a.a_count = kc + 16 + (d->kind == CVAE_WGRAD_SHIFT_FRAMES ? 4 * PW : 0);
            a.b_count = kc + (d->kind == CVAE_WGRAD_SHIFT_FRAMES ? 0 : 2 * halo + 2);
        } else {
            a.a_count = kc;
            a.b_count = kc + 2 * halo;
        }
        // odd plane strides (in 16-byte slots): the 8 lanes that copy the 8 planes of one pixel hit 8 different
        // bank groups instead of one
        a.stride_a = (a.a_count | 1) * 16;
        a.stride_b = (a.b_count | 1) * 16;
        // an M = 128 MMA reads 16 plane-strided core-matrix groups: keep them inside the buffer even when
        // fewer planes are real (their accumulator rows are never stored)
        const int a_planes_alloc = a.shift_a ? a_planes : 16;
        a.a_region = (a_planes_alloc * a.stride_a + 1023) & ~1023;
        a.buf_bytes = (a.a_region + (b_planes + a.ones_planes) * a.stride_b + 1023) & ~1023;
        if ((size_t)2 * a.buf_bytes <= cap && a.buf_bytes < (1 << 17)) break;
    }
    a.num_chunks = (int)((total_v + kc - 1) / kc);
    if (splits > a.num_chunks) splits = a.num_chunks;
    a.splits = splits;

    // groups
    int gi = 0, out_off = 0;
    if (d->kind == CVAE_WGRAD_5X5 || d->kind == CVAE_WGRAD_PHASE) {
        const int K = (d->kind == CVAE_WGRAD_5X5) ? 5 : 3;
        for (int t = 0; t < K * K; ++t) {
            const int dy = t / K - pad, dx = t % K - pad;
            a.g[gi++] = WgGroup{0, (halo + dy * PW + dx) * 16, n, out_off};
            out_off += a.m_total * n;
        }
        f.bias_off = -1;
        if (d->dbias) {
            f.bias_off = out_off;
            a.g[gi++] = WgGroup{0, b_planes * a.stride_b + halo * 16, 16, out_off};
            out_off += a.m_total * 16;
        }
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
    CVAE_REQUIRE(gi == a.groups_total, CVAE_EINVAL, "conv_wgrad: internal group count");
    a.split_floats = out_off;
    a.partial = (float*)d->workspace;
    a.fault = fault_flag();
    a.dbg = g_wg_dbg;
    CVAE_REQUIRE(a.fault != nullptr, CVAE_ECUDA, "conv_wgrad: fault flag unavailable");
    CVAE_REQUIRE((int64_t)splits * a.split_floats * 4 <= cvae_conv_wgrad_workspace_bytes(d), CVAE_EINVAL,
                 "conv_wgrad: workspace too small");

    const size_t smem = (size_t)2 * a.buf_bytes;
    dim3 grid(splits, gsets, a.m_blocks);
    if (getenv("CVAE_DEBUG"))
        fprintf(stderr, "conv_wgrad kind %d B=%d %dx%d cout=%d cin=%d: n=%d groups=%d gpc=%d gsets=%d mblocks=%d splits=%d kc=%d chunks=%d "
                        "buf=%d B smem=%zu\n", d->kind, d->batch, H, W, d->cout, d->cin, n, a.groups_total, a.gpc, gsets, a.m_blocks, splits,
                a.kc, a.num_chunks, a.buf_bytes, (size_t)2 * a.buf_bytes);
    int rc;
    if (la == CVAE_LOAD_NHWC) rc = launch_wgrad<CVAE_LOAD_NHWC, CVAE_LOAD_NHWC>(a, smem, grid, stream);
    else if (la == CVAE_LOAD_S2D) rc = launch_wgrad<CVAE_LOAD_S2D, CVAE_LOAD_NHWC>(a, smem, grid, stream);
    else if (la == CVAE_LOAD_NCHW3) rc = launch_wgrad<CVAE_LOAD_NCHW3, CVAE_LOAD_NHWC>(a, smem, grid, stream);
    else rc = launch_wgrad<CVAE_LOAD_S2D_NCHW3_DTANH, CVAE_LOAD_NHWC>(a, smem, grid, stream);
    if (rc != CVAE_OK) return rc;

    if (d->fold_stream != nullptr && d->fold_stream != stream_) {   // fold on a second stream, ordered after the GEMM
        cudaEvent_t ev;
        CVAE_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e1 = cudaEventRecord(ev, stream);
        cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent((cudaStream_t)d->fold_stream, ev, 0) : e1;
        cudaEventDestroy(ev);   // released once the recorded work has completed
        CVAE_REQUIRE(e2 == cudaSuccess, CVAE_ECUDA, "conv_wgrad: fold stream dependency failed: %s", cudaGetErrorString(e2));
        stream = (cudaStream_t)d->fold_stream;
    }
    f.dbg_linear = getenv("CVAE_FOLD_LINEAR") ? 1 : 0;
    f.kind = d->kind; f.cout = d->cout; f.cin = d->cin; f.splits = splits; f.split_floats = a.split_floats;
    f.m_total = a.m_total; f.n = n; f.partial = a.partial; f.dw = (float*)d->dw; f.dbias = (float*)d->dbias;
    {
        const int total = d->cout * d->cin * 25 + d->cout;
        if (d->kind == CVAE_WGRAD_5X5 || d->kind == CVAE_WGRAD_PHASE) {
            const int total4 = d->cout * d->cin * 25 / 4;
#define CVAE_FOLD_ROWS(L_)                                                                            \
    {                                                                                                 \
        const int wb = (total4 + 256 / (L_) - 1) / (256 / (L_));                                      \
        const int bb = f.dbias ? (d->cout + 256 / (L_) - 1) / (256 / (L_)) : 0;                       \
        wgrad_fold_rows_kernel<L_><<<wb + bb, 256, 0, stream>>>(f, wb);                               \
    }
            if (splits >= 48) CVAE_FOLD_ROWS(8)
            else if (splits >= 16) CVAE_FOLD_ROWS(4)
            else CVAE_FOLD_ROWS(2)
#undef CVAE_FOLD_ROWS
        } else {
            wgrad_fold_kernel<8><<<(total + 31) / 32, 256, 0, stream>>>(f);
        }
    }
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}
