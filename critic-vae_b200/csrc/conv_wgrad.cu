// Weight-gradient GEMM on tcgen05:  acc_g[m][n] = sum_v A[v + a_g][m] * B[v + b_g][n],  K = pixels.
//
// Two operand layouts, one GEMM structure:
//   conv_wgrad_tma_kernel (further down): layers with >= 64 channels on both operands; 128-byte-swizzled
//                          MN-major tiles written by TMA boxes of the NHWC tensors.
//   conv_wgrad_kernel    : everything else (32-channel and fp32 sources).  Both operands are halo planes
// ([channel/8][virtual pixel][8 ch], see planes.cuh) read MN-major:
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
#include "tma.cuh"

namespace cvae {

static constexpr int kMaxGroups = 28;
// experiment switches, read once at load time (never on the launch path)
static const bool g_wg_debug = getenv("CVAE_DEBUG") != nullptr;
static const bool g_wg_no_tma = getenv("CVAE_WG_NO_TMA") != nullptr;
static const bool g_wg_no_stack = getenv("CVAE_WG_NO_STACK") != nullptr;
static const int g_wg_stack_pw = getenv("CVAE_WG_STACK_PW") ? atoi(getenv("CVAE_WG_STACK_PW")) : 0;   // experiments: force the tap-stacked plan
static const int g_wg_stack_ra = getenv("CVAE_WG_STACK_RA") ? atoi(getenv("CVAE_WG_STACK_RA")) : 0;
static const int g_wg_kc = getenv("CVAE_WG_KC") ? atoi(getenv("CVAE_WG_KC")) : 0;
static const int g_fold_linear = getenv("CVAE_FOLD_LINEAR") ? 1 : 0;
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
    grid_dependency_sync();     // everything above touched only this CTA's shared memory and TMEM
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
            if constexpr (!kAsyncA) fill_planes<LA>(a.pa, a.dPW, a.dIH, A, a.stride_a, c0 + a.a_first, a.a_count, tid, kWgLoaders);
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
    int stack;         // tap-stacked partial (launch_wgrad_stack): a group holds the taps of one (pair of) filter row(s), its n columns are (dx, ci)
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
    } else if (f.kind == CVAE_WGRAD_SHIFT_FRAMES) {
        if (f.stack) {   // launch_wgrad_frames: group 0 = filter rows 3..0 as row blocks 0..3 of 32 channels, group 1 = row 4; col = (kx, ch)
            const int g = ky < 4 ? 0 : 1, j = ky < 4 ? 3 - ky : 0;
            return __ldg(p + ((size_t)g * f.m_total + j * f.cout + co) * f.n + kx * 8 + ci);
        }
        // group = ky, row = (kx, ch), col = co
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
    grid_dependency_sync();
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
    grid_dependency_sync();
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
        if (f.stack) {   // group g = filter rows (2g, 2g+1) as (row block 1, row block 0) of 64 channels; group 2 = filter row 4 in block 0
            const int g = ky < 4 ? (ky >> 1) : 2, j = ky < 4 ? 1 - (ky & 1) : 0;
            off[0] = ((size_t)g * f.m_total + j * f.cout + co) * f.n + kx * f.cin + ci;
        } else {
            off[0] = ((size_t)(ky * 5 + kx) * f.m_total + co) * f.n + ci;
        }
    } else {
        nsrc = 4;
#pragma unroll
        for (int ab = 0; ab < 4; ++ab) {
            const int ty = phase_tap(ab >> 1, ky), tx = phase_tap(ab & 1, kx);
            off[ab] = f.stack ? ((size_t)ty * f.m_total + ab * f.cout + co) * f.n + tx * f.cin + ci
                              : ((size_t)(ty * 3 + tx) * f.m_total + ab * f.cout + co) * f.n + ci;
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
    CVAE_OPT_IN_SMEM(kern, smem);
    cvae::launch(kern, grid, kWgThreads, smem, stream, a);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}


// =============================================================================================
// TMA-fed variant for layers whose operands have >= 64 channels on both sides (E2, E3, D0, D1).
//
// Operand tiles are MN-major with 128-byte rows ([pixel][64 channels], SWIZZLE_128B), written by
// cp.async.bulk.tensor boxes {64 ch, W + pad, rows, images} of the NHWC tensors: the zero padding columns / rows of
// the virtual pixel space come from out-of-bounds fill, so a box lands as consecutive virtual pixels.  A filter tap is
// the B descriptor's start address moved by dy * PW + dx rows; tools/umma_probe_swz.py verified that tcgen05.mma
// applies the 128-byte swizzle to ABSOLUTE shared-memory address bits (base offset 0), so a tile written by TMA can be
// read from any 128-byte row.  tools/tma_probe.cu: such boxes arrive at ~80 B/clk/SM against ~15 B/clk for the
// 16-byte-granular planes.
// Chunk = NB whole images (small maps) or R rows of one image; everything TMA never writes (margins around the B box,
// the K tail that rounds a chunk up to 16 pixels) is zeroed once at kernel start and stays zero.
// =============================================================================================
struct WgTmaArgs {
    int rows_mode;              // 0: chunk = NB whole images, 1: chunk = RA rows of one image
    int NB, RA, pad;            // images per box; rows per A box (B box has RA + 2 pad rows in rows mode)
    int blocks_per_image;       // rows mode: row blocks per image
    int num_chunks, splits;
    int kc;                     // K per chunk (multiple of 16)
    int b_blocks;               // 64-channel column blocks of B (N / 64)
    int a_block_bytes, b_block_bytes;   // column-block strides (multiples of 1024)
    int b_box_row;              // 128-byte row inside a B block where the box lands
    int b_base_row;             // B row that pairs with A row 0 for the centre tap
    int buf_bytes, a_region, nbuf;   // nbuf: 2..4 chunk buffers (TMA latency is longer than the MMAs of a short chunk)
    uint32_t tx_bytes;          // bytes of all boxes of one chunk
    int phase_maps;             // A column block q of M block mb: 1 -> tensor map mb*2+q (channel 0), 0 -> map 0, channel (mb*2+q)*64
    // Tap-stacked variant for 32-channel operands (launch_wgrad_stack): 64-byte rows (SWIZZLE_64B) and accumulator groups whose
    // N columns are several taps side by side -- column block i of B is the SAME tile one pixel (one row of the tile) further on.
    int a_loads;                // A boxes per chunk: 2 (two 64-channel blocks), 4 (four 32-channel phase blocks) or 1 (block 1 = block 0 one virtual row on)
    int a_row0;                 // rows mode: chunks start this many image rows early (1 when block 1 is the shifted block 0)
    int a_kstep, b_kstep;       // descriptor units (16 B) per K = 16 step: 128 for 128-byte rows, 64 for 64-byte rows
    int b_row_units;            // 16-byte units per B row (8 or 4)
    uint32_t a_hi, b_hi;        // high descriptor words (SBO, version, swizzle type) of A and B
    // Frame-operand variant (launch_wgrad_frames, encoder conv 0): B is not a TMA tile but one no-swizzle plane
    // [virtual pixel][3 frame channels + 5 zeros] that the eight epilogue warps convert from the fp32 NCHW frames
    // themselves while the GEMM runs (b_fill = 1, b_blocks = 0); column block i of B is that plane one pixel further on.
    int b_fill, b_count, b_pad_rows;   // slots per chunk; virtual rows above each image in the plane's pixel numbering
    PlaneSrc pb;
    FastDiv dPW, dIH;
    int ones_off, ones_stride;  // bias pseudo-group operand: no-swizzle plane pair of ones after the buffers (0: none)
    int m_rows, groups_total, gpc, split_floats;
    int tap_row[kMaxGroups];    // dy * PW + dx per group
    int gn[kMaxGroups];         // UMMA N per group
    int gout[kMaxGroups];       // float offset of the group's block in one split's partial
    float* partial;
    int* fault;
};

static constexpr int kWtThreads = 320;   // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue

__global__ void __launch_bounds__(kWtThreads, 1)
conv_wgrad_tma_kernel(const WgTmaArgs a, const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                      const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                      const __grid_constant__ CUtensorMap mapB) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_full[4], bar_empty[4], bar_acc;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int split = blockIdx.x, gset = blockIdx.y, mb = blockIdx.z;
    const int g0 = gset * a.gpc, g1 = min(g0 + a.gpc, a.groups_total);

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(&bar_full[i], a.b_fill ? 1 + (kWtThreads / 32 - 2) : 1); mbar_init(&bar_empty[i], 1); }
        mbar_init(&bar_acc, 1);
        mbar_fence_init();
    }
    int cols = 0;
    for (int g = g0; g < g1; ++g) cols += a.gn[g];
    uint32_t ncols = 32;
    while (ncols < (uint32_t)cols) ncols <<= 1;
    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    // zero everything once (margins and K tails stay zero), then the ones tile
    const int total16 = (a.nbuf * a.buf_bytes) >> 4;
    for (int i = tid; i < total16; i += kWtThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (a.ones_off) {
        for (int q = 0; q < 2; ++q) fill_ones_plane(smem + a.ones_off + (size_t)q * a.ones_stride, a.kc, tid, kWtThreads);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    grid_dependency_sync();     // everything above touched only this CTA's shared memory and TMEM
    const uint32_t tmem_base = tmem_slot;
    const int my_chunks = (a.num_chunks - split + a.splits - 1) / a.splits;
    const uint32_t smem_base = smem_u32(smem);

    if (warp == 0) {
        // ------------------------------ TMA producer --------------------------------------------
        if (elect_one()) {
            const CUtensorMap* mapsA[4] = {&mapA0, &mapA1, &mapA2, &mapA3};
            bool alive = true;
            for (int i = 0; i < my_chunks && alive; ++i) {
                const int buf = i % a.nbuf;
                const int chunk = split + i * a.splits;
                alive = mbar_wait(&bar_empty[buf], ((i / a.nbuf) & 1) ^ 1, a.fault);
                int n, ha, hb;
                if (a.rows_mode) {
                    n = chunk / a.blocks_per_image;
                    ha = (chunk - n * a.blocks_per_image) * a.RA - a.a_row0;
                    hb = ha - a.pad;
                } else {
                    n = chunk * a.NB;
                    ha = hb = -a.pad;
                }
                mbar_expect_tx(&bar_full[buf], a.tx_bytes);
                const uint32_t A = smem_base + (uint32_t)buf * a.buf_bytes;
                const uint32_t Bp = A + a.a_region + (uint32_t)(a.b_box_row * a.b_row_units) * 16u;
                for (int q = 0; q < a.a_loads; ++q) {
                    const int blk = mb * 2 + q;
                    tma_load_4d(A + (uint32_t)q * a.a_block_bytes, a.phase_maps ? mapsA[blk] : mapsA[0], a.phase_maps ? 0 : blk * 64, 0, ha, n,
                                &bar_full[buf]);
                }
                for (int q = 0; q < a.b_blocks; ++q)
                    tma_load_4d(Bp + (uint32_t)q * a.b_block_bytes, &mapB, q * 64, 0, hb, n, &bar_full[buf]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ----------------------------------------------
        if (elect_one()) {
            bool alive = true;
            // swizzled MN-major descriptors: LBO = column-block stride, SBO = 1024 (8 rows), layout type 2; the K loop
            // moves both start addresses by 16 rows = 2048 B = 128 descriptor units
            const uint32_t hi_sw = a.a_hi;
            const uint32_t a_lbo = (((uint32_t)a.a_block_bytes >> 4) & 0x3FFFu) << 16;
            const uint32_t b_lbo = (((uint32_t)a.b_block_bytes >> 4) & 0x3FFFu) << 16;
            const uint32_t hi_ones = (((uint32_t)a.ones_stride >> 4) & 0x3FFFu) | (1u << 14);   // no swizzle: SBO = plane stride
            const uint32_t ones_lbo = (128u >> 4) << 16;
            const int ksteps = a.kc >> 4;
            const uint32_t a_step = (uint32_t)a.a_kstep;
            for (int i = 0; i < my_chunks && alive; ++i) {
                const int buf = i % a.nbuf;
                alive = mbar_wait(&bar_full[buf], (i / a.nbuf) & 1, a.fault);
                tc_fence_after();
                const uint32_t A16 = (((smem_base + (uint32_t)buf * a.buf_bytes) & 0x3FFFFu) >> 4);
                const uint32_t B16 = A16 + ((uint32_t)a.a_region >> 4);
                uint32_t col = tmem_base;
                for (int g = g0; g < g1; ++g) {
                    const int n = a.gn[g];
                    const uint32_t idesc = umma_idesc_bf16(n, kMajorMN, kMajorMN);
                    uint32_t a_lo = A16 | a_lbo;
                    uint32_t b_lo, b_hi, b_step;
                    if (n == 16) {   // bias pseudo-group against the ones tile
                        b_lo = (((smem_base + (uint32_t)a.ones_off) & 0x3FFFFu) >> 4) | ones_lbo;
                        b_hi = hi_ones;
                        b_step = 16u;    // 16 pixel slots of 16 B
                    } else {
                        b_lo = (B16 + (uint32_t)((a.b_base_row + a.tap_row[g]) * a.b_row_units)) | b_lbo;
                        b_hi = a.b_hi;
                        b_step = (uint32_t)a.b_kstep;
                    }
                    uint32_t acc = i > 0 ? 1u : 0u;
                    int k = 0;
                    for (; k + 4 <= ksteps; k += 4) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            umma_bf16(col, ((uint64_t)hi_sw << 32) | (uint64_t)(a_lo + (uint32_t)j * a_step),
                                      ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)j * b_step), idesc, j == 0 ? acc : 1u);
                        acc = 1u;
                        a_lo += 4u * a_step;
                        b_lo += 4u * b_step;
                    }
                    for (; k < ksteps; ++k) {
                        umma_bf16(col, ((uint64_t)hi_sw << 32) | (uint64_t)a_lo, ((uint64_t)b_hi << 32) | (uint64_t)b_lo, idesc, acc);
                        acc = 1u;
                        a_lo += a_step;
                        b_lo += b_step;
                    }
                    col += (uint32_t)n;
                }
                umma_commit(&bar_empty[buf]);
            }
            umma_commit(&bar_acc);
        }
        __syncwarp();
    } else {
        if (a.b_fill) {
            // ------------------------------ B plane loaders (the epilogue warps, idle until the last chunk) ------------
            bool alive = true;
            const int ltid = tid - 64;
            for (int i = 0; i < my_chunks && alive; ++i) {
                const int buf = i % a.nbuf;
                const int chunk = split + i * a.splits;
                alive = mbar_wait(&bar_empty[buf], ((i / a.nbuf) & 1) ^ 1, a.fault);
                const int n = chunk / a.blocks_per_image;
                const int ha = (chunk - n * a.blocks_per_image) * a.RA - a.a_row0;
                // slot 0 = frame pixel (row ha + 1, column -2) of image n: the first tap of group 0 for A row 0
                const int v_first = (n * a.pb.IH + ha + 1 + a.b_pad_rows) * a.pb.PW - 2;
                fill_planes<CVAE_LOAD_NCHW3>(a.pb, a.dPW, a.dIH, smem + (size_t)buf * a.buf_bytes + a.a_region, 0, v_first, a.b_count, ltid,
                                             kWtThreads - 64);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[buf]);
            }
        }
        // ------------------------------ epilogue (8 warps: lane quarter = warp % 4, groups dealt by parity) -------------
        mbar_wait(&bar_acc, 0, a.fault);
        tc_fence_after();
        const int quarter = warp & 3, par = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        float* base = a.partial + (size_t)split * a.split_floats;
        uint32_t col = 0;
        for (int g = g0; g < g1; ++g) {
            const int n = a.gn[g];
            if (((g - g0) & 1) != par) { col += n; continue; }
            float* o = base + a.gout[g] + ((size_t)mb * a.m_rows + row) * n;
            for (int c = 0; c < n; c += 16) {
                uint32_t raw[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + col + c, raw);
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
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, ncols);
}

TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (TensorMapEncodeFn)p;
    }
    return fn;
}

// bf16 tensor viewed as {channels, w, h, n} with element strides (in elements) for w, h, n; box {64, bw, bh, bn}, SWIZZLE_128B
static bool encode_map(CUtensorMap* m, const void* base, int channels, int W, int H, int B, long sw, long sh, long sn, int bw, int bh, int bn) {
    return encode_map_4d(m, base, channels, W, H, B, sw, sh, sn, 64, bw, bh, bn, CU_TENSOR_MAP_SWIZZLE_128B);
}

// Plans and launches the TMA variant; returns 1 when the shape is not eligible (the caller falls back to the plane kernel).
static int launch_wgrad_tma(const cvae_wgrad_desc* d, const WgradArgs& base, int n, int gsets, int H, int W, int pad, cudaStream_t stream,
                            int* splits_out) {
    const int mtot = (d->kind == CVAE_WGRAD_5X5) ? d->cout : 4 * d->cout;
    if (g_wg_no_tma) return 1;
    if (!(d->kind == CVAE_WGRAD_5X5 || d->kind == CVAE_WGRAD_PHASE)) return 1;
    if (mtot % 128 != 0 || d->cin % 64 != 0 || d->cin > 256) return 1;
    if (d->kind == CVAE_WGRAD_PHASE && d->cout != 64) return 1;
    if (d->kind == CVAE_WGRAD_5X5 && d->cout % 64 != 0) return 1;
    if (!tensor_map_encoder()) return 1;
    const int PW = W + pad, IH = H + pad, halo = pad * PW + pad;
    WgTmaArgs t{};
    t.pad = pad;
    t.b_blocks = d->cin / 64;
    const bool bias = d->dbias != nullptr;
    const int m_blocks = mtot / 128;
    int splits = *splits_out;   // in: the most splits the workspace was sized for; out: the number used
    if (splits < 1) splits = 1;
    const size_t cap = 212 * 1024;
    auto plan = [&](int rows_mode, int NB, int RA) -> bool {
        const int kreal = rows_mode ? RA * PW : NB * IH * PW;
        const int kc = (kreal + 15) / 16 * 16;
        const int box_rows_b = rows_mode ? (RA + 2 * pad) * PW : kreal;
        const int b_box_row = rows_mode ? 8 : (halo + 7) / 8 * 8;
        const int b_rows = b_box_row + box_rows_b + (kc - kreal) + halo + 8;
        const int a_block = (kc * 128 + 1023) & ~1023, b_block = (b_rows * 128 + 1023) & ~1023;
        const size_t buf = (size_t)2 * a_block + (size_t)t.b_blocks * b_block;
        const size_t ones = bias ? (size_t)2 * (((kc * 16) + 127) & ~127) : 0;
        if (2 * buf + ones > cap || a_block >= (1 << 18) || b_block >= (1 << 18)) return false;
        int nbuf = (int)((cap - ones) / buf);
        t.nbuf = nbuf > 4 ? 4 : nbuf;
        t.rows_mode = rows_mode; t.NB = NB; t.RA = RA; t.kc = kc;
        t.b_box_row = b_box_row;
        t.b_base_row = b_box_row + (rows_mode ? pad * PW : 0);
        t.a_block_bytes = a_block; t.b_block_bytes = b_block;
        t.a_region = 2 * a_block;
        t.buf_bytes = (int)buf;
        t.ones_off = bias ? (int)(t.nbuf * buf) : 0;
        t.ones_stride = (kc * 16 + 127) & ~127;
        t.tx_bytes = (uint32_t)(2 * kreal * 128 + t.b_blocks * box_rows_b * 128);
        t.blocks_per_image = rows_mode ? (H + RA - 1) / RA : 1;
        t.num_chunks = rows_mode ? d->batch * t.blocks_per_image : (d->batch + NB - 1) / NB;
        return true;
    };
    // Candidates: NB whole images per chunk (pad rows included) or RA rows of one image.  Pick the one that spends the
    // fewest K steps per image among those with a useful chunk length (>= 96 pixels) and >= 2 chunks per CTA.
    bool ok = false;
    {
        const int per_cta_images = (d->batch + splits - 1) / splits;
        int best_mode = -1, best_nb = 0, best_ra = 0;
        double best_cost = 1e30;
        auto consider = [&](int rows_mode, int NB, int RA) {
            if (!plan(rows_mode, NB, RA)) return;
            const double k_per_image = rows_mode ? (double)t.blocks_per_image * t.kc : (double)t.kc / NB;
            const double chunks_per_cta = (double)t.num_chunks / splits;
            double cost = k_per_image;
            if (t.kc < 96) cost *= 96.0 / t.kc;            // per-chunk overhead of short chunks
            if (chunks_per_cta < 2.0) cost *= 2.0;         // no load / MMA overlap
            if (t.nbuf < 3) cost *= 1.15;                  // two buffers do not hide the TMA latency of short chunks
            if (cost < best_cost) { best_cost = cost; best_mode = rows_mode; best_nb = NB; best_ra = RA; }
        };
        for (int NB = 1; NB <= 256 && NB <= per_cta_images && IH <= 256; ++NB) consider(0, NB, IH);
        for (int RA = 1; RA <= H; ++RA) consider(1, 1, RA);
        if (best_mode >= 0) ok = plan(best_mode, best_nb, best_ra);
    }
    if (!ok) return 1;
    if (splits > t.num_chunks) splits = t.num_chunks;
    t.splits = splits;

    // tensor maps
    CUtensorMap mA[4], mB;
    const int ba_rows = t.rows_mode ? t.RA : IH, bb_rows = t.rows_mode ? t.RA + 2 * pad : IH;
    bool enc_ok = true;
    if (d->kind == CVAE_WGRAD_5X5) {
        enc_ok = encode_map(&mA[0], d->dy, d->cout, W, H, d->batch, d->cout, (long)W * d->cout, (long)H * W * d->cout, PW, ba_rows, t.NB);
        mA[1] = mA[2] = mA[3] = mA[0];
        t.phase_maps = 0;
    } else {   // dY at [B][2H][2W][cout]: one strided view per output phase (a, b)
        const long c = d->cout;
        for (int ab = 0; ab < 4 && enc_ok; ++ab) {
            const __nv_bfloat16* b0 = (const __nv_bfloat16*)d->dy + ((long)(ab >> 1) * 2 * W + (ab & 1)) * c;
            enc_ok = encode_map(&mA[ab], b0, d->cout, W, H, d->batch, 2 * c, 2L * 2 * W * c, 4L * H * W * c, PW, ba_rows, t.NB);
        }
        t.phase_maps = 1;
    }
    enc_ok = enc_ok && encode_map(&mB, d->x, d->cin, W, H, d->batch, d->cin, (long)W * d->cin, (long)H * W * d->cin, PW, bb_rows, t.NB);
    if (!enc_ok) return 1;

    t.m_rows = 128;
    t.a_loads = 2; t.a_row0 = 0; t.a_kstep = 128; t.b_kstep = 128; t.b_row_units = 8;
    t.a_hi = t.b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);     // SBO 1024 B (8 rows of 128 B), version 1, SWIZZLE_128B
    t.groups_total = base.groups_total; t.gpc = base.gpc; t.split_floats = base.split_floats;
    for (int g = 0; g < base.groups_total; ++g) {
        t.gn[g] = base.g[g].n;
        t.gout[g] = base.g[g].out_off;
        t.tap_row[g] = base.g[g].n == 16 ? 0 : base.g[g].b_off / 16 - halo;
    }
    t.partial = base.partial; t.fault = base.fault;
    const size_t smem = (size_t)t.nbuf * t.buf_bytes + (t.ones_off ? (size_t)2 * t.ones_stride : 0);
    CVAE_OPT_IN_SMEM(conv_wgrad_tma_kernel, smem);
    if (g_wg_debug)
        fprintf(stderr, "conv_wgrad_tma kind %d %dx%d cout=%d cin=%d: %s NB=%d RA=%d kc=%d nbuf=%d chunks=%d splits=%d gsets=%d mblocks=%d buf=%d smem=%zu\n",
                d->kind, H, W, d->cout, d->cin, t.rows_mode ? "rows" : "images", t.NB, t.RA, t.kc, t.nbuf, t.num_chunks, splits, gsets, m_blocks,
                t.buf_bytes, smem);
    dim3 grid(splits, gsets, m_blocks);
    cvae::launch(conv_wgrad_tma_kernel, grid, kWtThreads, smem, stream, t, mA[0], mA[1], mA[2], mA[3], mB);
    CVAE_LAUNCH_CHECK();
    *splits_out = splits;
    return CVAE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Tap-stacked variant of the TMA kernel for 32-channel inputs (encoder conv 1: 5x5, 32 -> 64; decoder conv 3 in phase
// form: 3x3, 32 -> 4 x 32).  With N = 32 columns per tap the tensor pipe runs at a third of its rate and every tap is its
// own MMA; here the B tile has 64-byte rows ([virtual pixel][32 ch], SWIZZLE_64B) and ONE accumulator group covers all
// the dx of a filter row: column block i of the MN-major B descriptor is the same tile one pixel further on (LBO = 64 B),
// so N = 5 x 32 = 160 (3 x 32 = 96).  With 64 output channels the M = 128 rows are stacked as well: row block 1 of A is
// row block 0 one virtual image row further on (LBO = PW x 128 B), which makes it the filter row above -- 25 taps in
// 3 MMAs per K step instead of 25.  Chunks are RA image rows; with the shifted A block they start one row early so that
// both blocks see every image row exactly once, and RA x PW must then be a multiple of 16 (PW is widened with extra
// zero columns -- out-of-bounds box columns -- until a good RA exists).
// ---------------------------------------------------------------------------------------------------------------
struct WgStackShape { int ndx, ngroups, n, mstack; long split_floats; };
static bool wgrad_stack_shape(const cvae_wgrad_desc* d, WgStackShape& s) {
    if (d->cin != 32) return false;
    if (d->kind == CVAE_WGRAD_5X5 && d->cout == 64) { s.ndx = 5; s.mstack = 1; }
    else if (d->kind == CVAE_WGRAD_PHASE && d->cout == 32) { s.ndx = 3; s.mstack = 0; }
    else return false;
    s.ngroups = 3;
    s.n = s.ndx * 32;
    s.split_floats = (long)s.ngroups * 128 * s.n + (d->dbias ? 128 * 16 : 0);
    return true;
}

// returns 1 when the shape is not eligible (the caller goes on with the other variants)
static int launch_wgrad_stack(const cvae_wgrad_desc* d, int H, int W, int pad, cudaStream_t stream, int max_splits, WgStackShape& s, int* splits_out,
                              int* bias_off_out) {
    if (g_wg_no_tma || g_wg_no_stack || !wgrad_stack_shape(d, s) || !tensor_map_encoder()) return 1;
    const bool bias = d->dbias != nullptr;
    WgTmaArgs t{};
    t.pad = pad;
    t.rows_mode = 1; t.NB = 1;
    t.b_blocks = 1;
    t.a_row0 = s.mstack;
    t.a_loads = s.mstack ? 1 : 4;
    const int a_row_bytes = s.mstack ? 128 : 64;
    const size_t cap = 212 * 1024;
    int splits = max_splits < 1 ? 1 : max_splits;
    int best_pw = 0, best_ra = 0;
    double best_cost = 1e30;
    int PW = 0, halo = 0;
    auto plan = [&](int pw, int RA) -> bool {
        const int kreal = RA * pw;
        if (s.mstack && (kreal & 15)) return false;
        const int kc = (kreal + 15) / 16 * 16;
        if (pw > 256 || RA + 2 * pad > 256) return false;
        const int hl = pad * pw + pad;
        const int a_rows = kc + (s.mstack ? pw : 0);
        const int a_block = s.mstack ? pw * 128 : ((kc * 64 + 1023) & ~1023);       // LBO of A
        const int a_region = s.mstack ? ((a_rows * 128 + 1023) & ~1023) : 4 * a_block;
        const int b_box_row = 8;
        const int b_rows = b_box_row + (RA + 2 * pad) * pw + (kc - kreal) + hl + 16;
        const int b_bytes = (b_rows * 64 + 1023) & ~1023;
        const size_t buf = (size_t)a_region + b_bytes;
        const size_t ones = bias ? (size_t)2 * (((kc * 16) + 127) & ~127) : 0;
        if (2 * buf + ones > cap || a_region >= (1 << 18) || b_bytes >= (1 << 18)) return false;
        int nbuf = (int)((cap - ones) / buf);
        t.nbuf = nbuf > 4 ? 4 : nbuf;
        t.RA = RA; t.kc = kc;
        t.b_box_row = b_box_row;
        t.b_base_row = b_box_row + pad * pw;
        t.a_block_bytes = a_block; t.b_block_bytes = 64;      // B: the next column block is the next pixel
        t.a_region = a_region;
        t.buf_bytes = (int)buf;
        t.ones_off = bias ? (int)(t.nbuf * buf) : 0;
        t.ones_stride = (kc * 16 + 127) & ~127;
        t.tx_bytes = (uint32_t)((s.mstack ? (RA + 1) * pw * 128 : 4 * RA * pw * 64) + (RA + 2 * pad) * pw * 64);
        t.blocks_per_image = (H + s.mstack + RA - 1) / RA;
        t.num_chunks = d->batch * t.blocks_per_image;
        PW = pw; halo = hl;
        return true;
    };
    for (int pw = W + pad; pw <= W + pad + 8; ++pw)
        for (int RA = 1; RA <= H + s.mstack; ++RA) {
            if (!plan(pw, RA)) continue;
            const int per_cta = (t.num_chunks + splits - 1) / splits;
            double cost = (double)per_cta * (t.kc + 64);          // K per CTA plus a per-chunk overhead worth ~64 pixels
            if (per_cta < 2) cost *= 2.0;                          // (measured, profiles/r02_wgrad_stack_sweep.log: two buffers are enough here)
            if (cost < best_cost) { best_cost = cost; best_pw = pw; best_ra = RA; }
        }
    if (g_wg_stack_pw > 0 && g_wg_stack_ra > 0 && plan(W + pad + g_wg_stack_pw - 1, g_wg_stack_ra)) { best_pw = W + pad + g_wg_stack_pw - 1; best_ra = g_wg_stack_ra; }
    if (best_pw == 0 || !plan(best_pw, best_ra)) return 1;
    if (splits > t.num_chunks) splits = t.num_chunks;
    t.splits = splits;

    CUtensorMap mA[4], mB;
    bool ok = true;
    if (s.mstack) {
        ok = encode_map_4d(&mA[0], d->dy, d->cout, W, H, d->batch, d->cout, (long)W * d->cout, (long)H * W * d->cout, 64, PW, t.RA + 1, 1,
                           CU_TENSOR_MAP_SWIZZLE_128B);
        mA[1] = mA[2] = mA[3] = mA[0];
        t.phase_maps = 0;
        t.a_kstep = 128;
        t.a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    } else {   // dY at [B][2H][2W][32]: one strided 32-channel view per output phase
        const long c = d->cout;
        for (int ab = 0; ab < 4 && ok; ++ab) {
            const __nv_bfloat16* b0 = (const __nv_bfloat16*)d->dy + ((long)(ab >> 1) * 2 * W + (ab & 1)) * c;
            ok = encode_map_4d(&mA[ab], b0, d->cout, W, H, d->batch, 2 * c, 2L * 2 * W * c, 4L * H * W * c, 32, PW, t.RA, 1, CU_TENSOR_MAP_SWIZZLE_64B);
        }
        t.phase_maps = 1;
        t.a_kstep = 64;
        t.a_hi = (512u >> 4) | (1u << 14) | (4u << 29);
    }
    ok = ok && encode_map_4d(&mB, d->x, d->cin, W, H, d->batch, d->cin, (long)W * d->cin, (long)H * W * d->cin, 32, PW, t.RA + 2 * pad, 1,
                             CU_TENSOR_MAP_SWIZZLE_64B);
    if (!ok) return 1;
    t.b_kstep = 64; t.b_row_units = 4;
    t.b_hi = (512u >> 4) | (1u << 14) | (4u << 29);           // SBO 512 B (8 rows of 64 B), version 1, SWIZZLE_64B

    t.m_rows = 128;
    t.groups_total = s.ngroups + (bias ? 1 : 0);
    t.gpc = t.groups_total;
    t.split_floats = (int)s.split_floats;
    for (int g = 0; g < s.ngroups; ++g) {
        const int dy = s.mstack ? (g == 0 ? -1 : g) : g - 1;      // block 0's filter row; block 1 (stacked M) is the row above
        t.gn[g] = s.n;
        t.gout[g] = g * 128 * s.n;
        t.tap_row[g] = dy * PW - pad;
    }
    *bias_off_out = -1;
    if (bias) {
        t.gn[s.ngroups] = 16;
        t.gout[s.ngroups] = s.ngroups * 128 * s.n;
        t.tap_row[s.ngroups] = 0;
        *bias_off_out = t.gout[s.ngroups];
    }
    t.partial = (float*)d->workspace; t.fault = fault_flag();
    if (t.fault == nullptr) return 1;
    const size_t smem = (size_t)t.nbuf * t.buf_bytes + (t.ones_off ? (size_t)2 * t.ones_stride : 0);
    CVAE_OPT_IN_SMEM(conv_wgrad_tma_kernel, smem);
    if (g_wg_debug)
        fprintf(stderr, "conv_wgrad_stack kind %d %dx%d cout=%d cin=%d: PW=%d RA=%d kc=%d nbuf=%d chunks=%d splits=%d n=%d buf=%d smem=%zu\n", d->kind, H, W,
                d->cout, d->cin, PW, t.RA, t.kc, t.nbuf, t.num_chunks, splits, s.n, t.buf_bytes, smem);
    cvae::launch(conv_wgrad_tma_kernel, dim3(splits, 1, 1), kWtThreads, smem, stream, t, mA[0], mA[1], mA[2], mA[3], mB);
    CVAE_LAUNCH_CHECK();
    *splits_out = splits;
    return CVAE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Encoder conv 0 (3 -> 32 channels, 64x64 frames): dW[co][ch][ky][kx] = sum_q dY[q][co] * F[q + (ky - 2) PW + kx - 2][ch].
// The plane kernel runs it as 5 MMAs of N = 32 per K step with 15 useful rows out of 128 (16 pixel shifts x 8 channel
// slots) and is bound by the N = 32 MMA rate (profiles/r02_e0w_counters.log: 117 k cycles, 9 k of them waiting for data).
// Here the roles are swapped and both operands are stacked:
//   A = dY, TMA box of [virtual pixel][32 co] rows (64 B, SWIZZLE_64B); its four 32-row blocks are the SAME tile 0, 1, 2, 3
//       image rows further on (LBO = PW x 64 B), i.e. filter rows ky_g, ky_g - 1, ky_g - 2, ky_g - 3;
//   B = the frames as ONE no-swizzle plane [virtual pixel][3 ch + 5 zeros], converted from fp32 NCHW by the epilogue warps;
//       its six 8-column blocks are the plane 0..5 pixels further on (SBO = 16 B), i.e. kx = 0..4 (+ one unused).
// Two MMAs of N = 48 per K step (ky_g = 3: filter rows 3..0; ky_g = 4: row 4 in block 0) instead of five of N = 32.
// Chunks are RA image rows starting three rows early (all four blocks see every row once); RA x PW is a multiple of 16.
// ---------------------------------------------------------------------------------------------------------------
static constexpr int kFramesN = 48, kFramesGroups = 2;
static bool wgrad_frames_shape(const cvae_wgrad_desc* d) {
    return d->kind == CVAE_WGRAD_SHIFT_FRAMES && d->cout == 32 && d->cin == 3 && d->dbias == nullptr;
}

static int launch_wgrad_frames(const cvae_wgrad_desc* d, int H, int W, int pad, cudaStream_t stream, int max_splits, int* splits_out) {
    if (g_wg_no_tma || g_wg_no_stack || !wgrad_frames_shape(d) || !tensor_map_encoder()) return 1;
    WgTmaArgs t{};
    t.pad = pad;
    t.rows_mode = 1; t.NB = 1;
    t.b_blocks = 0;                                  // no TMA for B
    t.b_fill = 1;
    t.a_row0 = 3;
    t.a_loads = 1;
    const int pad_rows = 16;                         // virtual rows above each image in the plane's pixel numbering (>= every overshoot)
    const size_t cap = 212 * 1024;
    int splits = max_splits < 1 ? 1 : max_splits;
    int best_pw = 0, best_ra = 0;
    double best_cost = 1e30;
    int PW = 0;
    auto plan = [&](int pw, int RA) -> bool {
        const int kc = RA * pw;
        if ((kc & 15) || pw > 256 || RA + 3 > 256) return false;
        const int a_region = (((RA + 3) * pw * 64) + 1023) & ~1023;
        const int b_count = (kc + pw + 5 + 3 + 15) & ~15;
        const int b_bytes = (b_count * 16 + 1023) & ~1023;
        const size_t buf = (size_t)a_region + b_bytes;
        if (2 * buf > cap || a_region >= (1 << 18)) return false;
        int nbuf = (int)(cap / buf);
        t.nbuf = nbuf > 4 ? 4 : nbuf;
        t.RA = RA; t.kc = kc;
        t.a_block_bytes = pw * 64;                   // LBO of A: the next 32-row block is the next image row
        t.a_region = a_region;
        t.b_block_bytes = 128;                       // LBO of B (no swizzle): 8 pixels of 16 B along K
        t.b_count = b_count;
        t.buf_bytes = (int)buf;
        t.tx_bytes = (uint32_t)((RA + 3) * pw * 64);
        t.blocks_per_image = (H + 3 + RA - 1) / RA;
        t.num_chunks = d->batch * t.blocks_per_image;
        PW = pw;
        return true;
    };
    for (int pw = W + pad; pw <= W + pad + 8; ++pw)
        for (int RA = 1; RA <= H + 3; ++RA) {
            if (!plan(pw, RA)) continue;
            if ((t.blocks_per_image - 1) * RA - 3 + 1 + (t.kc + pw + 8) / pw + 1 - H >= pad_rows) continue;   // rows below the image must stay virtual
            const int per_cta = (t.num_chunks + splits - 1) / splits;
            double cost = (double)per_cta * (t.kc + 64);
            if (per_cta < 2) cost *= 2.0;
            if (cost < best_cost) { best_cost = cost; best_pw = pw; best_ra = RA; }
        }
    if (g_wg_stack_pw > 0 && g_wg_stack_ra > 0 && plan(W + pad + g_wg_stack_pw - 1, g_wg_stack_ra)) { best_pw = W + pad + g_wg_stack_pw - 1; best_ra = g_wg_stack_ra; }
    if (best_pw == 0 || !plan(best_pw, best_ra)) return 1;
    if (splits > t.num_chunks) splits = t.num_chunks;
    t.splits = splits;

    CUtensorMap mA[4], mB;
    if (!encode_map_4d(&mA[0], d->dy, d->cout, W, H, d->batch, d->cout, (long)W * d->cout, (long)H * W * d->cout, 32, PW, t.RA + 3, 1,
                       CU_TENSOR_MAP_SWIZZLE_64B))
        return 1;
    mA[1] = mA[2] = mA[3] = mB = mA[0];
    t.phase_maps = 0;
    t.a_kstep = 64;
    t.a_hi = (512u >> 4) | (1u << 14) | (4u << 29);  // SBO 512 B (8 rows of 64 B), version 1, SWIZZLE_64B
    t.b_kstep = 16; t.b_row_units = 1;               // plane slots of 16 B
    t.b_hi = (16u >> 4) | (1u << 14);                // SBO 16 B: the next 8 columns are the plane one pixel on; no swizzle
    t.b_box_row = 0; t.b_base_row = 0;
    t.b_pad_rows = pad_rows;
    t.pb = PlaneSrc{d->batch, H, W, pad_rows, PW, H + pad_rows, 1, 3, 0, d->x, nullptr};
    t.dPW = make_fastdiv(PW);
    t.dIH = make_fastdiv(H + pad_rows);

    t.m_rows = 128;
    t.groups_total = kFramesGroups;
    t.gpc = kFramesGroups;
    t.split_floats = kFramesGroups * 128 * kFramesN;
    for (int g = 0; g < kFramesGroups; ++g) {
        t.gn[g] = kFramesN;
        t.gout[g] = g * 128 * kFramesN;
        t.tap_row[g] = g * PW;                       // group 1's block 0 is one filter row further down
    }
    t.partial = (float*)d->workspace; t.fault = fault_flag();
    if (t.fault == nullptr) return 1;
    const size_t smem = (size_t)t.nbuf * t.buf_bytes;
    CVAE_OPT_IN_SMEM(conv_wgrad_tma_kernel, smem);
    if (g_wg_debug)
        fprintf(stderr, "conv_wgrad_frames %dx%d: PW=%d RA=%d kc=%d nbuf=%d chunks=%d splits=%d buf=%d smem=%zu\n", H, W, PW, t.RA, t.kc, t.nbuf,
                t.num_chunks, splits, t.buf_bytes, smem);
    cvae::launch(conv_wgrad_tma_kernel, dim3(splits, 1, 1), kWtThreads, smem, stream, t, mA[0], mA[1], mA[2], mA[3], mB);
    CVAE_LAUNCH_CHECK();
    *splits_out = splits;
    return CVAE_OK;
}

}  // namespace cvae

using namespace cvae;

static unsigned long long* g_wg_dbg = nullptr;
// Profiling aid: when set, every weight-gradient CTA writes 8 cycle counters to buf[cta * 8 ..]: MMA thread
// total / wait-for-data; loaders wait-for-buffer / issue / landing wait / chunks; epilogue.
extern "C" void cvae_wgrad_debug_counters(void* device_buf) { g_wg_dbg = (unsigned long long*)device_buf; }

// GEMM shape of one weight-gradient call, shared by the workspace query and the launcher so the two can never
// disagree: accumulator groups (filter taps + bias pseudo-groups), how they are dealt to CTAs (gsets sets of gpc
// groups filling the 512 TMEM columns), the floats of one split-K partial, and the largest split count either
// kernel variant (plane or TMA) will use.
struct WgShape {
    int n, ngroups, bias_groups, groups_total, m_blocks, m_rows, m_total, gpc, gsets, max_splits;
    long split_floats;
};
static bool wgrad_shape(const cvae_wgrad_desc* d, WgShape& s) {
    switch (d->kind) {
        case CVAE_WGRAD_5X5:
        case CVAE_WGRAD_PHASE: {
            const int mtot = (d->kind == CVAE_WGRAD_5X5) ? d->cout : 4 * d->cout;
            s.n = d->cin; s.ngroups = (d->kind == CVAE_WGRAD_5X5) ? 25 : 9; s.bias_groups = d->dbias ? 1 : 0;
            s.m_blocks = (mtot + 127) / 128; s.m_rows = mtot < 128 ? mtot : 128;
            break;
        }
        case CVAE_WGRAD_SHIFT_FRAMES: s.n = d->cout; s.ngroups = 5; s.bias_groups = 0; s.m_blocks = 1; s.m_rows = 128; break;
        case CVAE_WGRAD_SHIFT_PHASE12: s.n = d->cin; s.ngroups = 6; s.bias_groups = 2; s.m_blocks = 1; s.m_rows = 128; break;
        default: return false;
    }
    if (s.n <= 0 || s.n > 256) return false;
    s.m_total = s.m_blocks * s.m_rows;
    s.groups_total = s.ngroups + s.bias_groups;
    // groups per CTA: fill the 512 TMEM columns (every group is n columns wide; the bias pseudo-groups 16)
    s.gpc = 512 / s.n;
    if (s.gpc < 1) s.gpc = 1;
    while (s.gpc > 1) {   // keep the column total of every set <= 512 including trailing 16-column pseudo-groups
        bool ok = true;
        for (int g0 = 0; g0 < s.groups_total && ok; g0 += s.gpc) {
            int cols = 0;
            for (int g = g0; g < g0 + s.gpc && g < s.groups_total; ++g) cols += (g < s.ngroups) ? s.n : 16;
            ok = cols <= 512;
        }
        if (ok) break;
        --s.gpc;
    }
    s.gsets = (s.groups_total + s.gpc - 1) / s.gpc;
    s.gpc = (s.groups_total + s.gsets - 1) / s.gsets;   // same number of sets, balanced
    s.max_splits = d->splits > 0 ? d->splits : (sm_count() / (s.gsets * s.m_blocks));
    if (s.max_splits < 1) s.max_splits = 1;
    s.split_floats = (long)s.ngroups * s.m_total * s.n + (long)s.bias_groups * s.m_total * 16;
    return true;
}

// splits x floats-per-split of the variant the launcher would pick for exactly this descriptor
static int64_t wgrad_workspace_exact(const cvae_wgrad_desc* d) {
    WgShape s;
    if (!wgrad_shape(d, s)) return -1;
    int64_t bytes = (int64_t)s.max_splits * s.split_floats * 4;   // the plane and the TMA variant use <= max_splits splits
    if (wgrad_frames_shape(d)) {
        const int64_t sb = (int64_t)(d->splits > 0 ? d->splits : sm_count()) * kFramesGroups * 128 * kFramesN * 4;
        if (sb > bytes) bytes = sb;
    }
    WgStackShape st;
    if (wgrad_stack_shape(d, st)) {                               // the tap-stacked variant: one CTA per SM, wider partials
        const int64_t sb = (int64_t)(d->splits > 0 ? d->splits : sm_count()) * st.split_floats * 4;
        if (sb > bytes) bytes = sb;
    }
    return bytes;
}

extern "C" int64_t cvae_conv_wgrad_workspace_bytes(const cvae_wgrad_desc* d) {
    if (!d) return -1;
    // The bias pseudo-group changes both the partial size and the split count: callers routinely size the workspace before
    // they fill in dbias, so the answer covers both forms of the call.
    cvae_wgrad_desc with = *d, without = *d;
    with.dbias = (void*)(uintptr_t)8;
    without.dbias = nullptr;
    const int64_t a = wgrad_workspace_exact(&with), b = wgrad_workspace_exact(&without);
    if (a < 0 || b < 0) return -1;
    return a > b ? a : b;
}

extern "C" int cvae_conv_wgrad(const cvae_wgrad_desc* d, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(d != nullptr, CVAE_EINVAL, "conv_wgrad: null descriptor");
    CVAE_REQUIRE(d->x && d->dy && d->dw && d->workspace, CVAE_EINVAL, "conv_wgrad: null tensor");
    CVAE_REQUIRE(d->batch > 0, CVAE_EINVAL, "conv_wgrad: empty batch");
    if (d->workspace_bytes > 0) {
        const int64_t need = wgrad_workspace_exact(d);
        CVAE_REQUIRE(need >= 0 && need <= d->workspace_bytes, CVAE_EINVAL, "conv_wgrad: workspace of %lld bytes, this call needs %lld",
                     (long long)d->workspace_bytes, (long long)need);
    }

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

    WgShape shape;
    CVAE_REQUIRE(wgrad_shape(d, shape), CVAE_EINVAL, "conv_wgrad: unsupported shape");
    CVAE_REQUIRE(shape.n == n && shape.m_total == a.m_total && shape.m_blocks == a.m_blocks, CVAE_EINVAL, "conv_wgrad: internal shape mismatch");
    ngroups = shape.ngroups;
    a.groups_total = shape.groups_total;
    a.gpc = shape.gpc;
    const int gsets = shape.gsets;
    int splits = shape.max_splits;

    // chunk size: the largest multiple of 16 pixels whose double buffer fits, but small enough that every
    // CTA still gets ~4 chunks to pipeline
    const size_t cap = 216 * 1024;
    long per_cta = (total_v + splits - 1) / splits;
    const int kq = 16 * kWgKS;   // the MMA issuer unrolls kWgKS K steps
    int kc_want = (int)(((per_cta + 3) / 4 + kq - 1) / kq * kq);
    if (kc_want < 64) kc_want = 64;
    if (kc_want > 512) kc_want = 512;
    if (g_wg_kc > 0) kc_want = g_wg_kc / kq * kq;
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
    CVAE_REQUIRE((long)a.split_floats == shape.split_floats, CVAE_EINVAL, "conv_wgrad: internal partial size mismatch");

    const size_t smem = (size_t)2 * a.buf_bytes;
    dim3 grid(splits, gsets, a.m_blocks);
    if (g_wg_debug)
        fprintf(stderr, "conv_wgrad kind %d B=%d %dx%d cout=%d cin=%d: n=%d groups=%d gpc=%d gsets=%d mblocks=%d splits=%d kc=%d chunks=%d "
                        "buf=%d B smem=%zu\n", d->kind, d->batch, H, W, d->cout, d->cin, n, a.groups_total, a.gpc, gsets, a.m_blocks, splits,
                a.kc, a.num_chunks, a.buf_bytes, (size_t)2 * a.buf_bytes);
    WgStackShape stack{};
    int stack_splits = 0, stack_bias_off = -1;
    int rc = launch_wgrad_stack(d, H, W, pad, stream, d->splits > 0 ? d->splits : sm_count(), stack, &stack_splits, &stack_bias_off);
    if (rc < 0) return rc;
    bool stacked = rc == CVAE_OK;
    bool frames_stacked = false;
    if (!stacked) {
        rc = launch_wgrad_frames(d, H, W, pad, stream, d->splits > 0 ? d->splits : sm_count(), &stack_splits);
        if (rc < 0) return rc;
        if (rc == CVAE_OK) {
            stacked = frames_stacked = true;
            stack.n = kFramesN;
            stack.split_floats = kFramesGroups * 128 * kFramesN;
        }
    }
    int tma_splits = shape.max_splits;
    if (!stacked) rc = launch_wgrad_tma(d, a, n, gsets, H, W, pad, stream, &tma_splits);
    if (rc < 0) return rc;
    if (stacked) {                                // tap-stacked TMA variant: its own partial layout (FoldArgs::stack)
        splits = stack_splits;
        n = stack.n;
        a.split_floats = (int)stack.split_floats;
        a.m_total = 128;
        f.bias_off = stack_bias_off;
        f.stack = 1;
    } else if (rc == CVAE_OK) splits = tma_splits;   // TMA variant launched with its own chunking
    else if (la == CVAE_LOAD_NHWC) rc = launch_wgrad<CVAE_LOAD_NHWC, CVAE_LOAD_NHWC>(a, smem, grid, stream);
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
    f.dbg_linear = g_fold_linear;
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
        cvae::launch(wgrad_fold_rows_kernel<L_>, wb + bb, 256, 0, stream, f, wb);                               \
    }
            if (splits >= 48) CVAE_FOLD_ROWS(8)
            else if (splits >= 16) CVAE_FOLD_ROWS(4)
            else CVAE_FOLD_ROWS(2)
#undef CVAE_FOLD_ROWS
        } else {
            // few outputs (2.4 k), many splits: 32 split lanes per output keep the chain of dependent loads short (this fold
            // of encoder conv 0 is the last kernel in front of the optimizer)
            cvae::launch(wgrad_fold_kernel<32>, (total + 7) / 8, 256, 0, stream, f);
        }
    }
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}
