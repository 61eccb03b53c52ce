// The three Linear layers around the latent (0.17 % of the network FLOPs): fc_mu || fc_var
// (vae_nets.py:98-99,108-109) on the NHWC-flattened 4x4x256 bottleneck, and decoder_input
// (vae_nets.py:137,143-144).  CUDA-core fp32 kernels: K = 4096 / 33 and N = 64 / 4096 are a poor
// fit for 128-row UMMA tiles, and these are latency-, not throughput-critical.
//
// Weight layouts come from pack.cu: wfc fp32 [4096 k'][64] (k' = pixel*256 + channel),
// wdec fp32 [34][4096 k'] (rows 0..32 = W^T, row 33 = bias).
#include <cooperative_groups.h>

#include "common.cuh"
#include "umma.cuh"

namespace cg = cooperative_groups;

namespace cvae {

// ---- fc forward: ml[b][j] = sum_k a[b][k] * wfc[k][j] + bias[j];  a bf16 [B][4096] ---------------
// A thread-block cluster of 8 CTAs shares one block of 16 rows: CTA r of the cluster owns the K slice
// [512 r, 512 r + 512), pulls its 128 KB weight slice and its 16 activation rows into shared memory with
// bulk async copies (one mbarrier), and leaves its [16][64] partial in its own shared memory; after
// cluster.sync() CTA r sums the eight partials of rows 2r, 2r+1 over distributed shared memory in rank
// order and adds the bias -- one launch, no atomics, bit-reproducible.
static constexpr int kFcRows = 16, kFcSplit = 8, kFcSlice = 4096 / kFcSplit;
static constexpr size_t kFcSmem = (size_t)kFcSlice * 64 * 4 + (size_t)kFcRows * kFcSlice * 2;
// [16 rows][64] partial of one K slice: thread -> output j = tid % 64, rows 4 (tid / 64) .. + 3 (a warp shares the rows: the
// activation reads are broadcasts).  Shared by fc_fwd_kernel and bottleneck_fwd_kernel (bit-identical results).
template <int ROWS = 16>
__device__ __forceinline__ void fc_slice_partial(const uint32_t* __restrict__ sa, const float* __restrict__ sw, float (*part)[64]) {
    constexpr int RPT = ROWS / 4;     // rows per thread (ROWS = 16: the layout the stand-alone kernel has always had)
    const int j = threadIdx.x & 63, rq = threadIdx.x >> 6;
    float acc[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) acc[i] = 0.f;
#pragma unroll 2
    for (int k8 = 0; k8 < kFcSlice / 8; ++k8) {
        uint4 xr[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) xr[i] = *reinterpret_cast<const uint4*>(sa + (rq * RPT + i) * (kFcSlice / 2) + k8 * 4);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const float wv = sw[(k8 * 8 + kk) * 64 + j];
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const uint32_t u = (kk >> 1) == 0 ? xr[i].x : ((kk >> 1) == 1 ? xr[i].y : ((kk >> 1) == 2 ? xr[i].z : xr[i].w));
                acc[i] = fmaf(wv, (kk & 1) ? bf16_hi(u) : bf16_lo(u), acc[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) part[rq * RPT + i][j] = acc[i];
}

__global__ void __cluster_dims__(kFcSplit, 1, 1) __launch_bounds__(256)
fc_fwd_kernel(int B, const __nv_bfloat16* __restrict__ a, const float* __restrict__ wfc,
              const float* __restrict__ bmu, const float* __restrict__ bvar, float* __restrict__ ml, int* fault) {
    grid_dependency_sync();
    extern __shared__ __align__(128) uint8_t fc_smem[];
    float* sw = reinterpret_cast<float*>(fc_smem);                                        // [512][64]
    uint32_t* sa = reinterpret_cast<uint32_t*>(fc_smem + (size_t)kFcSlice * 64 * 4);      // [16][256] bf16 pairs
    __shared__ float part[kFcRows][64];
    __shared__ uint64_t bar;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b0 = (blockIdx.x / kFcSplit) * kFcRows, k0 = rank * kFcSlice;
    const int rows = min(kFcRows, B - b0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = threadIdx.x; i < (kFcRows - rows) * (kFcSlice / 2); i += 256) sa[rows * (kFcSlice / 2) + i] = 0u;
    __syncthreads();
    if (threadIdx.x < 32) {
        if (elect_one()) {
            mbar_expect_tx(&bar, (uint32_t)(kFcSlice * 64 * 4 + rows * kFcSlice * 2));
            for (int c = 0; c < 8; ++c)
                bulk_g2s(fc_smem + (size_t)c * 16384, reinterpret_cast<const uint8_t*>(wfc + (size_t)k0 * 64) + (size_t)c * 16384, 16384, &bar);
            for (int r = 0; r < rows; ++r)
                bulk_g2s(sa + r * (kFcSlice / 2), a + (size_t)(b0 + r) * 4096 + k0, kFcSlice * 2, &bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0, fault);
    fc_slice_partial(sa, sw, part);
    cluster.sync();
    if (threadIdx.x < 128) {
        const int r = rank * 2 + (threadIdx.x >> 6), jj = threadIdx.x & 63;
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < kFcSplit; ++q) s += cluster.map_shared_rank(&part[0][0], q)[r * 64 + jj];
        if (b0 + r < B) ml[(size_t)(b0 + r) * 64 + jj] = s + (jj < 32 ? bmu[jj] : bvar[jj - 32]);
    }
    cluster.sync();   // nobody leaves while a peer may still read its partial
}

// fc backward (data) for ONE output k' and R rows.  The 64 terms are added four at a time starting at quad
// rot = (k' / 2) % 16 and wrapping around: with the weight rows in shared memory (256-byte pitch) the 8 lanes of a quarter
// warp then read 8 different quads = all 32 banks (no rotation: one bank group, 8-way conflict).  fc_bwd_data_kernel reads
// its rows from global memory and uses the same order, so the two paths stay bit-identical.
template <int R, bool kShared>
__device__ __forceinline__ void fc_bwd_rows(const float4* __restrict__ wr, int rot, const float (*sd)[64], float* acc) {
#pragma unroll 4
    for (int qq = 0; qq < 16; ++qq) {
        const int q = (qq + rot) & 15;
        const float4 w = kShared ? wr[q] : __ldg(wr + q);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 d = *reinterpret_cast<const float4*>(&sd[r][4 * q]);
            acc[r] += w.x * d.x + w.y * d.y + w.z * d.z + w.w * d.w;
        }
    }
}

// ---- fc backward (data): da[b][k'] = sum_j dml[b][j] * wfc[k'][j]  (gradient w.r.t. the Tanh output) ----
__global__ void fc_bwd_data_kernel(int B, const float* __restrict__ dml, const float* __restrict__ wfc,
                                   __nv_bfloat16* __restrict__ da) {
    grid_dependency_sync();
    __shared__ __align__(16) float sd[8][64];
    const int b0 = blockIdx.x * 8;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        const int r = i >> 6, jj = i & 63;
        sd[r][jj] = (b0 + r < B) ? dml[(size_t)(b0 + r) * 64 + jj] : 0.f;
    }
    __syncthreads();
    for (int k = blockIdx.y * blockDim.x + threadIdx.x; k < 4096; k += gridDim.y * blockDim.x) {
        float acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = 0.f;
        fc_bwd_rows<8, false>(reinterpret_cast<const float4*>(wfc + (size_t)k * 64), (k >> 1) & 15, sd, acc);
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (b0 + r < B) da[(size_t)(b0 + r) * 4096 + k] = __float2bfloat16_rn(acc[r]);
    }
}

// ---- fc backward (weights): dW[j][k] = sum_b dml[b][j] * a[b][k'], written in the reference's
//      [32][4096] (k = c*16 + p) layout for fc_mu and fc_var; biases = column sums of dml. --------
// block = 32 k' x 8 batch lanes (one warp per lane: the dml row is a warp-uniform broadcast load, the
// activations a coalesced 64-byte segment); the 8 lane sums are combined through shared memory in lane
// order, so the result is bit-reproducible.
static constexpr int kWLanes = 8;
__global__ void __launch_bounds__(256) fc_bwd_weight_kernel(int B, const float* __restrict__ dml, const __nv_bfloat16* __restrict__ a,
                                                            float* __restrict__ dwmu, float* __restrict__ dwvar,
                                                            float* __restrict__ dbmu, float* __restrict__ dbvar) {
    grid_dependency_sync();
    __shared__ float red[32][65];
    const int kq = threadIdx.x & 31, bl = threadIdx.x >> 5;
    const int kp = blockIdx.x * 32 + kq;  // k' in NHWC order
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;
    for (int b = bl; b < B; b += kWLanes) {
        const float x = __bfloat162float(a[(size_t)b * 4096 + kp]);
        const float4* dr = reinterpret_cast<const float4*>(dml + (size_t)b * 64);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 d = __ldg(dr + q);
            acc[4 * q] = fmaf(x, d.x, acc[4 * q]); acc[4 * q + 1] = fmaf(x, d.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(x, d.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(x, d.w, acc[4 * q + 3]);
        }
    }
    for (int l = 0; l < kWLanes; ++l) {
        if (bl == l) {
#pragma unroll
            for (int j = 0; j < 64; ++j) red[kq][j] = (l == 0 ? 0.f : red[kq][j]) + acc[j];
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < 32 * 64; t += 256) {
        const int q = t & 31, j = t >> 5;
        const int kpp = blockIdx.x * 32 + q;
        const int k = (kpp & 255) * 16 + (kpp >> 8);
        (j < 32 ? dwmu : dwvar)[(size_t)(j & 31) * 4096 + k] = red[q][j];
    }
    if (blockIdx.x == 0) {   // bias gradients: column sums of dml, 4 partial sums per column in fixed order
        __syncthreads();
        const int j = threadIdx.x & 63, q = threadIdx.x >> 6;
        float sacc = 0.f;
        for (int b = q; b < B; b += 4) sacc += dml[(size_t)b * 64 + j];
        red[q][j] = sacc;
        __syncthreads();
        if (threadIdx.x < 64) {
            const float t = (red[0][j] + red[1][j]) + (red[2][j] + red[3][j]);
            if (j < 32) dbmu[j] = t; else dbvar[j - 32] = t;
        }
    }
}

// ---- decoder_input forward: h[b][k'] = sum_i zc[b][i] * wdec[i][k'] + wdec[33][k'] -> bf16 NHWC ------
__global__ void decin_fwd_kernel(int B, const float* __restrict__ zc, const float* __restrict__ wdec,
                                 __nv_bfloat16* __restrict__ h) {
    grid_dependency_sync();
    __shared__ float sz[8][33];
    const int b0 = blockIdx.y * 8;
    for (int i = threadIdx.x; i < 8 * 33; i += blockDim.x) {
        const int r = i / 33;
        sz[r][i % 33] = (b0 + r < B) ? zc[(size_t)(b0 + r) * 33 + i % 33] : 0.f;
    }
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    float acc[8];
    const float bias = __ldg(wdec + 33 * 4096 + k);
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = bias;
    for (int i = 0; i < 33; ++i) {
        const float w = __ldg(wdec + (size_t)i * 4096 + k);
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = fmaf(w, sz[r][i], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
        if (b0 + r < B) h[(size_t)(b0 + r) * 4096 + k] = __float2bfloat16_rn(acc[r]);
}

// ---- decoder_input backward (data): dzc[b][i] = sum_k' dh[b][k'] * wdec[i][k'] ----------------------
// Same shape of solution as fc_fwd_kernel: a cluster of 8 CTAs splits K for a block of 16 batch rows, operands
// arrive by bulk async copies, partial [16][33] tiles are combined over distributed shared memory in rank order.
static constexpr int kDdRows = 16, kDdSplit = 8, kDdSlice = 4096 / kDdSplit, kDdWStride = kDdSlice + 4;  // floats; +4 keeps rows 16 B aligned
static constexpr size_t kDdSmem = (size_t)33 * kDdWStride * 4 + (size_t)kDdRows * kDdSlice * 2;
// [16 rows][33] partial of one K slice: thread -> row r = tid / 16, outputs i = ig, ig + 16 (and 32 for ig == 0).
// Shared by decin_bwd_data_kernel and bottleneck_bwd_kernel (bit-identical results).
__device__ __forceinline__ void decin_slice_partial(const uint32_t* __restrict__ sd, const float* __restrict__ sw, float (*part)[33]) {
    const int r = threadIdx.x >> 4, ig = threadIdx.x & 15;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    const float* w0 = sw + ig * kDdWStride;
    const float* w1 = sw + (ig + 16) * kDdWStride;
    const float* w2 = sw + 32 * kDdWStride;
    const uint32_t* xr = sd + r * (kDdSlice / 2);
    // four k per step: 16-byte weight loads (row stride 516 floats = 4 banks: the 8 lanes of a quarter warp cover all 32
    // banks; 8-byte loads were 2-way conflicted), same order of additions as before
#pragma unroll 2
    for (int k4 = 0; k4 < kDdSlice / 4; ++k4) {
        const uint2 u = *reinterpret_cast<const uint2*>(xr + 2 * k4);
        const float x0 = bf16_lo(u.x), x1 = bf16_hi(u.x), x2 = bf16_lo(u.y), x3 = bf16_hi(u.y);
        const float4 a = *reinterpret_cast<const float4*>(w0 + 4 * k4);
        const float4 b = *reinterpret_cast<const float4*>(w1 + 4 * k4);
        acc0 = fmaf(x0, a.x, acc0); acc0 = fmaf(x1, a.y, acc0); acc0 = fmaf(x2, a.z, acc0); acc0 = fmaf(x3, a.w, acc0);
        acc1 = fmaf(x0, b.x, acc1); acc1 = fmaf(x1, b.y, acc1); acc1 = fmaf(x2, b.z, acc1); acc1 = fmaf(x3, b.w, acc1);
        if (ig == 0) {
            const float4 c = *reinterpret_cast<const float4*>(w2 + 4 * k4);
            acc2 = fmaf(x0, c.x, acc2); acc2 = fmaf(x1, c.y, acc2); acc2 = fmaf(x2, c.z, acc2); acc2 = fmaf(x3, c.w, acc2);
        }
    }
    part[r][ig] = acc0;
    part[r][ig + 16] = acc1;
    if (ig == 0) part[r][32] = acc2;
}

__global__ void __cluster_dims__(kDdSplit, 1, 1) __launch_bounds__(256)
decin_bwd_data_kernel(int B, const __nv_bfloat16* __restrict__ dh, const float* __restrict__ wdec, float* __restrict__ dzc, int* fault) {
    grid_dependency_sync();
    extern __shared__ __align__(128) uint8_t dd_smem[];
    float* sw = reinterpret_cast<float*>(dd_smem);                                            // [33][516]
    uint32_t* sd = reinterpret_cast<uint32_t*>(dd_smem + (size_t)33 * kDdWStride * 4);         // [16][256] bf16 pairs
    __shared__ float part[kDdRows][33];
    __shared__ uint64_t bar;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b0 = (blockIdx.x / kDdSplit) * kDdRows, k0 = rank * kDdSlice;
    const int rows = min(kDdRows, B - b0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = threadIdx.x; i < (kDdRows - rows) * (kDdSlice / 2); i += 256) sd[rows * (kDdSlice / 2) + i] = 0u;
    __syncthreads();
    if (threadIdx.x < 32) {
        if (elect_one()) {
            mbar_expect_tx(&bar, (uint32_t)(33 * kDdSlice * 4 + rows * kDdSlice * 2));
            for (int i = 0; i < 33; ++i) bulk_g2s(sw + i * kDdWStride, wdec + (size_t)i * 4096 + k0, kDdSlice * 4, &bar);
            for (int r = 0; r < rows; ++r) bulk_g2s(sd + r * (kDdSlice / 2), dh + (size_t)(b0 + r) * 4096 + k0, kDdSlice * 2, &bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0, fault);
    decin_slice_partial(sd, sw, part);
    cluster.sync();
    if (threadIdx.x < 66) {
        const int rr = rank * 2 + threadIdx.x / 33, i = threadIdx.x % 33;
        float sacc = 0.f;
#pragma unroll
        for (int q = 0; q < kDdSplit; ++q) sacc += cluster.map_shared_rank(&part[0][0], q)[rr * 33 + i];
        if (b0 + rr < B) dzc[(size_t)(b0 + rr) * 33 + i] = sacc;
    }
    cluster.sync();
}

// ---- decoder_input backward (weights + bias), reference layout dW [4096 k][33], db [4096 k] ---------
// same decomposition as fc_bwd_weight_kernel: 32 k' x 8 batch lanes, fixed-order combine.
__global__ void __launch_bounds__(256) decin_bwd_weight_kernel(int B, const __nv_bfloat16* __restrict__ dh, const float* __restrict__ zc,
                                                               float* __restrict__ dw, float* __restrict__ db) {
    grid_dependency_sync();
    __shared__ float red[32][35];
    const int kq = threadIdx.x & 31, bl = threadIdx.x >> 5;
    const int kp = blockIdx.x * 32 + kq;
    float acc[34];
#pragma unroll
    for (int i = 0; i < 34; ++i) acc[i] = 0.f;
    for (int b = bl; b < B; b += kWLanes) {
        const float g = __bfloat162float(dh[(size_t)b * 4096 + kp]);
        const float* zr = zc + (size_t)b * 33;
#pragma unroll
        for (int i = 0; i < 33; ++i) acc[i] = fmaf(g, __ldg(zr + i), acc[i]);
        acc[33] += g;
    }
    for (int l = 0; l < kWLanes; ++l) {
        if (bl == l) {
#pragma unroll
            for (int i = 0; i < 34; ++i) red[kq][i] = (l == 0 ? 0.f : red[kq][i]) + acc[i];
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < 32 * 34; t += 256) {
        const int q = t / 34, i = t - q * 34;
        const int kpp = blockIdx.x * 32 + q;
        const int k = (kpp & 255) * 16 + (kpp >> 8);
        if (i < 33) dw[(size_t)k * 33 + i] = red[q][i];
        else db[k] = red[q][33];
    }
}

// =============================================================================================
// The bottleneck in one launch per direction (training path).  Both are latency chains of three small layers
// (0.17 % of the FLOPs, ~60 us of the step's main chain as separate launches):
//   forward :  fc_mu || fc_var  ->  reparametrise + critic concat  ->  decoder_input
//   backward:  decoder_input^T  ->  reparametrise backward (+ KL gradient)  ->  (fc_mu || fc_var)^T
// Same decomposition as the separate kernels: a cluster of 8 CTAs owns 16 batch rows; the K = 4096 reduction of the first
// layer is split over the cluster and combined over distributed shared memory in rank order; CTA r finishes rows 2r, 2r+1,
// does their latent arithmetic, publishes the 33 (64) values per row in its shared memory, and after one more cluster
// barrier every CTA reads all 16 rows and produces ITS 512-wide slice of the last layer.  The arithmetic is shared with
// the separate kernels (fc_slice_partial, decin_slice_partial, same expressions), so the results are bit-identical.
// =============================================================================================
// profiling aid (cvae_bottleneck_debug): CTA 0 writes clock64() at its phase boundaries, 8 slots per kernel (fwd 0.., bwd 8..)
__device__ long long* g_bn_dbg = nullptr;
#define BN_STAMP(slot)                                                                        \
    do {                                                                                      \
        if (g_bn_dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) g_bn_dbg[slot] = clock64(); \
    } while (0)

// every CTA: %globaltimer at entry and exit, slots 16 + 4 * blockIdx.x + {0, 1} (forward) / {2, 3} (backward)
__device__ __forceinline__ void bn_wall(int which) {
    if (g_bn_dbg != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_bn_dbg[16 + 4 * blockIdx.x + which] = (long long)t;
    }
}

// ROWS batch rows per cluster: 16, or 20 when 16 would need more clusters than the device holds at once (15 eight-CTA clusters
// at one CTA per SM: a 16th cluster -- batch 256 -- ran as a second wave and doubled the kernel's duration).
template <int ROWS> constexpr size_t bn_fwd_smem() { return (size_t)kFcSlice * 64 * 4 + (size_t)ROWS * kFcSlice * 2 + (size_t)34 * kFcSlice * 4; }
template <int ROWS>
__global__ void __cluster_dims__(kFcSplit, 1, 1) __launch_bounds__(256)
bottleneck_fwd_kernel(int B, const __nv_bfloat16* __restrict__ a, const float* __restrict__ wfc, const float* __restrict__ bmu,
                      const float* __restrict__ bvar, const float* __restrict__ eps, const float* __restrict__ pred,
                      const float* __restrict__ wdec, float* __restrict__ ml, float* __restrict__ zc, __nv_bfloat16* __restrict__ h,
                      int* fault) {
    bn_wall(0);
    extern __shared__ __align__(128) uint8_t bn_smem[];
    float* sw = reinterpret_cast<float*>(bn_smem);                                        // [512][64]
    uint32_t* sa = reinterpret_cast<uint32_t*>(bn_smem + (size_t)kFcSlice * 64 * 4);      // [16][256] bf16 pairs
    float* swd = reinterpret_cast<float*>(bn_smem + (size_t)kFcSlice * 64 * 4 + (size_t)ROWS * kFcSlice * 2);   // [34][512]
    constexpr int NOWN = (ROWS + kFcSplit - 1) / kFcSplit;     // rows this CTA finishes: rank, rank + 8, ...
    __shared__ float part[ROWS][64];
    __shared__ float ml_own[NOWN][64];
    __shared__ float z_own[NOWN][33];
    __shared__ float z_all[ROWS][33];
    __shared__ uint64_t bar[2];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b0 = (blockIdx.x / kFcSplit) * ROWS, k0 = rank * kFcSlice;
    const int rows = min(ROWS, B - b0);
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    for (int i = threadIdx.x; i < (ROWS - rows) * (kFcSlice / 2); i += 256) sa[rows * (kFcSlice / 2) + i] = 0u;
    __syncthreads();
    grid_dependency_sync();
    BN_STAMP(0);
    if (threadIdx.x < 32) {
        if (elect_one()) {
            mbar_expect_tx(&bar[0], (uint32_t)(kFcSlice * 64 * 4 + rows * kFcSlice * 2));
            for (int c = 0; c < 8; ++c)
                bulk_g2s(bn_smem + (size_t)c * 16384, reinterpret_cast<const uint8_t*>(wfc + (size_t)k0 * 64) + (size_t)c * 16384, 16384, &bar[0]);
            for (int r = 0; r < rows; ++r)
                bulk_g2s(sa + r * (kFcSlice / 2), a + (size_t)(b0 + r) * 4096 + k0, kFcSlice * 2, &bar[0]);
            mbar_expect_tx(&bar[1], (uint32_t)(34 * kFcSlice * 4));       // needed two cluster barriers from now
            for (int i = 0; i < 34; ++i) bulk_g2s(swd + i * kFcSlice, wdec + (size_t)i * 4096 + k0, kFcSlice * 4, &bar[1]);
        }
        __syncwarp();
    }
    mbar_wait(&bar[0], 0, fault);
    BN_STAMP(1);
    fc_slice_partial<ROWS>(sa, sw, part);
    BN_STAMP(2);
    cluster.sync();
    BN_STAMP(3);
    if (threadIdx.x < NOWN * 64) {               // rows rank, rank + 8, ..: combine the eight K slices in rank order (as fc_fwd_kernel)
        const int rr = threadIdx.x >> 6, r = rank + kFcSplit * rr, jj = threadIdx.x & 63;
        float s = 0.f;
        if (r < ROWS) {
#pragma unroll
            for (int q = 0; q < kFcSplit; ++q) s += cluster.map_shared_rank(&part[0][0], q)[r * 64 + jj];
            s += (jj < 32 ? bmu[jj] : bvar[jj - 32]);
            if (b0 + r < B) ml[(size_t)(b0 + r) * 64 + jj] = s;
        }
        ml_own[rr][jj] = s;
    }
    __syncthreads();
    if (threadIdx.x < NOWN * 33) {               // z = mu + eps * exp(logvar / 2) | critic value  (latent_fwd_kernel's expressions)
        const int rr = threadIdx.x / 33, d = threadIdx.x % 33, r = rank + kFcSplit * rr;
        float z = 0.f;
        if (r < ROWS && b0 + r < B) {
            if (d < 32) z = fmaf(__ldg(eps + (size_t)(b0 + r) * 32 + d), expf(0.5f * ml_own[rr][32 + d]), ml_own[rr][d]);
            else z = __ldg(pred + b0 + r);
            zc[(size_t)(b0 + r) * 33 + d] = z;
        }
        z_own[rr][d] = z;
    }
    cluster.sync();
    for (int i = threadIdx.x; i < ROWS * 33; i += 256) {
        const int r = i / 33, d = i - r * 33;
        z_all[r][d] = cluster.map_shared_rank(&z_own[0][0], r % kFcSplit)[(r / kFcSplit) * 33 + d];
    }
    __syncthreads();
    BN_STAMP(4);
    mbar_wait(&bar[1], 0, fault);
    BN_STAMP(5);
    {   // decoder_input slice: thread -> outputs k0 + 2 tid, + 1 for all 16 rows (decin_fwd_kernel's accumulation order)
        const int kk = 2 * threadIdx.x;
        float acc0[ROWS], acc1[ROWS];
        const float2 bias = *reinterpret_cast<const float2*>(swd + 33 * kFcSlice + kk);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) { acc0[r] = bias.x; acc1[r] = bias.y; }
        for (int i = 0; i < 33; ++i) {
            const float2 w = *reinterpret_cast<const float2*>(swd + i * kFcSlice + kk);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const float z = z_all[r][i];
                acc0[r] = fmaf(w.x, z, acc0[r]);
                acc1[r] = fmaf(w.y, z, acc1[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
            if (b0 + r < B) *reinterpret_cast<uint32_t*>(h + (size_t)(b0 + r) * 4096 + k0 + kk) = pack_bf16x2(acc0[r], acc1[r]);
    }
    BN_STAMP(6);
    cluster.sync();   // nobody leaves while a peer may still read its partials / latent rows
    BN_STAMP(7);
    bn_wall(1);
}

// decoder_input operands + this CTA's [512][64] fc slice; ROWS * 16 threads (one (row, output pair) per thread in the first phase)
template <int ROWS> constexpr size_t bn_bwd_smem() { return (size_t)33 * kDdWStride * 4 + (size_t)ROWS * kDdSlice * 2 + (size_t)kDdSlice * 64 * 4; }
template <int ROWS>
__global__ void __cluster_dims__(kDdSplit, 1, 1) __launch_bounds__(ROWS * 16)
bottleneck_bwd_kernel(int B, const __nv_bfloat16* __restrict__ dh, const float* __restrict__ wdec, const float* __restrict__ ml,
                      const float* __restrict__ eps, float kld_grad_scale, const float* __restrict__ wfc, float* __restrict__ dzc,
                      float* __restrict__ dml, __nv_bfloat16* __restrict__ da, int* fault) {
    bn_wall(2);
    extern __shared__ __align__(128) uint8_t bb_smem[];
    float* sw = reinterpret_cast<float*>(bb_smem);                                            // [33][516]
    uint32_t* sd = reinterpret_cast<uint32_t*>(bb_smem + (size_t)33 * kDdWStride * 4);         // [16][256] bf16 pairs
    float* swf = reinterpret_cast<float*>(bb_smem + (size_t)33 * kDdWStride * 4 + (size_t)ROWS * kDdSlice * 2);   // [512][64] rows k0 .. of wfc
    constexpr int NOWN = (ROWS + kDdSplit - 1) / kDdSplit, kThreads = ROWS * 16;
    __shared__ float part[ROWS][33];
    __shared__ float dz_own[NOWN][33];
    __shared__ float dml_own[NOWN][64];
    __shared__ __align__(16) float dml_all[ROWS][64];
    __shared__ uint64_t bar, bar_w;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b0 = (blockIdx.x / kDdSplit) * ROWS, k0 = rank * kDdSlice;
    const int rows = min(ROWS, B - b0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar_w, 1); mbar_fence_init(); }
    for (int i = threadIdx.x; i < (ROWS - rows) * (kDdSlice / 2); i += kThreads) sd[rows * (kDdSlice / 2) + i] = 0u;
    __syncthreads();
    grid_dependency_sync();
    BN_STAMP(8);
    if (threadIdx.x < 32) {
        if (elect_one()) {
            mbar_expect_tx(&bar, (uint32_t)(33 * kDdSlice * 4 + rows * kDdSlice * 2));
            for (int i = 0; i < 33; ++i) bulk_g2s(sw + i * kDdWStride, wdec + (size_t)i * 4096 + k0, kDdSlice * 4, &bar);
            for (int r = 0; r < rows; ++r) bulk_g2s(sd + r * (kDdSlice / 2), dh + (size_t)(b0 + r) * 4096 + k0, kDdSlice * 2, &bar);
            mbar_expect_tx(&bar_w, (uint32_t)(kDdSlice * 64 * 4));       // the last phase's weights: needed two cluster barriers from now
            for (int c = 0; c < 8; ++c)
                bulk_g2s(reinterpret_cast<uint8_t*>(swf) + (size_t)c * 16384, reinterpret_cast<const uint8_t*>(wfc + (size_t)k0 * 64) + (size_t)c * 16384,
                         16384, &bar_w);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0, fault);
    BN_STAMP(9);
    decin_slice_partial(sd, sw, part);
    BN_STAMP(10);
    cluster.sync();
    BN_STAMP(11);
    if (threadIdx.x < NOWN * 33) {               // rows rank, rank + 8, .. (as decin_bwd_data_kernel)
        const int rr = threadIdx.x / 33, r = rank + kDdSplit * rr, i = threadIdx.x % 33;
        float sacc = 0.f;
        if (r < ROWS) {
#pragma unroll
            for (int q = 0; q < kDdSplit; ++q) sacc += cluster.map_shared_rank(&part[0][0], q)[r * 33 + i];
            if (dzc && b0 + r < B) dzc[(size_t)(b0 + r) * 33 + i] = sacc;
        }
        dz_own[rr][i] = sacc;
    }
    __syncthreads();
    if (threadIdx.x < NOWN * 32) {               // reparametrise backward + KL gradient (latent_bwd_kernel's expressions, no external terms)
        const int rr = threadIdx.x >> 5, d = threadIdx.x & 31, r = rank + kDdSplit * rr;
        float om = 0.f, ol = 0.f;
        if (r < ROWS && b0 + r < B) {
            const float mu = __ldg(ml + (size_t)(b0 + r) * 64 + d), lv = __ldg(ml + (size_t)(b0 + r) * 64 + 32 + d);
            const float e = __ldg(eps + (size_t)(b0 + r) * 32 + d), dz = dz_own[rr][d];
            const float std_ = expf(0.5f * lv);
            om = dz + 0.f;
            ol = dz * e * 0.5f * std_ + 0.f;
            if (kld_grad_scale != 0.f) {
                om += kld_grad_scale * mu;
                ol += kld_grad_scale * 0.5f * (expf(lv) - 1.f);
            }
            dml[(size_t)(b0 + r) * 64 + d] = om;
            dml[(size_t)(b0 + r) * 64 + 32 + d] = ol;
        }
        dml_own[rr][d] = om;
        dml_own[rr][32 + d] = ol;
    }
    cluster.sync();
    for (int i = threadIdx.x; i < ROWS * 64; i += kThreads) {
        const int r = i >> 6, j = i & 63;
        dml_all[r][j] = cluster.map_shared_rank(&dml_own[0][0], r % kDdSplit)[(r / kDdSplit) * 64 + j];
    }
    __syncthreads();
    BN_STAMP(12);
    mbar_wait(&bar_w, 0, fault);
    if (threadIdx.x < 256) {   // (fc_mu || fc_var)^T slice: thread -> k' = k0 + 2 tid, + 1, all rows, weight rows from shared memory
        const int kl = 2 * threadIdx.x, kp = k0 + kl, rot = (kp >> 1) & 15;
        float acc0[ROWS], acc1[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc0[r] = acc1[r] = 0.f;
        fc_bwd_rows<ROWS, true>(reinterpret_cast<const float4*>(swf + (size_t)kl * 64), rot, dml_all, acc0);
        fc_bwd_rows<ROWS, true>(reinterpret_cast<const float4*>(swf + (size_t)(kl + 1) * 64), rot, dml_all, acc1);
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
            if (b0 + r < B) *reinterpret_cast<uint32_t*>(da + (size_t)(b0 + r) * 4096 + kp) = pack_bf16x2(acc0[r], acc1[r]);
    }
    BN_STAMP(13);
    cluster.sync();
    BN_STAMP(14);
    bn_wall(3);
}

}  // namespace cvae

using namespace cvae;

// How many 8-CTA clusters of a kernel the device holds at once (cudaOccupancyMaxActiveClusters; 15 on a B200 at one CTA
// per SM), cached per kernel.  The fused kernels pick 20 rows per cluster when 16 would need one cluster too many.
template <typename K>
static int max_clusters_of(K kern, int threads, size_t smem) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(128);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = -1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
    return n;
}
// rows per cluster: 16 unless that needs more clusters than fit at once and 20 does not
static int bottleneck_rows(int batch, int max_clusters) {
    static const int forced = getenv("CVAE_BOTTLENECK_ROWS") ? atoi(getenv("CVAE_BOTTLENECK_ROWS")) : 0;      // (experiments; read once)
    if (forced) return forced == 20 ? 20 : 16;
    if (max_clusters <= 0) return 16;
    const int c16 = (batch + 15) / 16, c20 = (batch + 19) / 20;
    return (c16 > max_clusters && c20 <= max_clusters) ? 20 : 16;
}

// profiling aid: 8-CTA clusters of the forward (which = 0) / backward (1) kernel (16-row form) the device holds at once with
// `smem` bytes of dynamic shared memory per CTA (0: the kernel's own size); negative on error
extern "C" int cvae_bottleneck_max_clusters(int which, int64_t smem) {
    if (which) {
        CVAE_OPT_IN_SMEM(bottleneck_bwd_kernel<16>, bn_bwd_smem<16>());
        return max_clusters_of(bottleneck_bwd_kernel<16>, 256, smem > 0 ? (size_t)smem : bn_bwd_smem<16>());
    }
    CVAE_OPT_IN_SMEM(bottleneck_fwd_kernel<16>, bn_fwd_smem<16>());
    return max_clusters_of(bottleneck_fwd_kernel<16>, 256, smem > 0 ? (size_t)smem : bn_fwd_smem<16>());
}

extern "C" int cvae_bottleneck_debug(void* device_buf16) {
    long long* p = (long long*)device_buf16;
    CVAE_CUDA(cudaMemcpyToSymbol(g_bn_dbg, &p, sizeof(p)));
    return CVAE_OK;
}

template <int ROWS>
static int launch_bottleneck_fwd(int batch, const void* act, const float* wfc, const float* bias_mu, const float* bias_var, const float* eps,
                                 const float* pred, const float* wdec, float* mu_logvar, float* z_pred, void* dec_in, int* fault, cudaStream_t stream) {
    CVAE_OPT_IN_SMEM(bottleneck_fwd_kernel<ROWS>, bn_fwd_smem<ROWS>());
    cvae::launch(bottleneck_fwd_kernel<ROWS>, ((batch + ROWS - 1) / ROWS) * kFcSplit, 256, bn_fwd_smem<ROWS>(), stream, batch, (const __nv_bfloat16*)act,
                 wfc, bias_mu, bias_var, eps, pred, wdec, mu_logvar, z_pred, (__nv_bfloat16*)dec_in, fault);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

// vae_nets.py:108-111 + :48-51 + :143-144 in one launch (training path: z is sampled with the caller's eps).
// Outputs: mu_logvar fp32 [B][64], z_pred fp32 [B][33], dec_in bf16 [B][4096] (NHWC 4x4x256).
extern "C" int cvae_bottleneck_fwd(int batch, const void* act, const float* wfc, const float* bias_mu, const float* bias_var,
                                   const float* eps, const float* pred, const float* wdec, float* mu_logvar, float* z_pred,
                                   void* dec_in, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(batch > 0 && act && wfc && bias_mu && bias_var && eps && pred && wdec && mu_logvar && z_pred && dec_in, CVAE_EINVAL,
                 "bottleneck_fwd: bad argument");
    int* fault = fault_flag();
    CVAE_REQUIRE(fault != nullptr, CVAE_ECUDA, "bottleneck_fwd: fault flag unavailable");
    static int cap_of_device[64] = {0};      // cluster capacity per device (0: not asked yet, -1: unknown)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (cap_of_device[dev] == 0) {
        CVAE_OPT_IN_SMEM(bottleneck_fwd_kernel<16>, bn_fwd_smem<16>());
        cap_of_device[dev] = max_clusters_of(bottleneck_fwd_kernel<16>, 256, bn_fwd_smem<16>());
    }
    const int max_clusters = cap_of_device[dev];
    if (bottleneck_rows(batch, max_clusters) == 20)
        return launch_bottleneck_fwd<20>(batch, act, wfc, bias_mu, bias_var, eps, pred, wdec, mu_logvar, z_pred, dec_in, fault, stream);
    return launch_bottleneck_fwd<16>(batch, act, wfc, bias_mu, bias_var, eps, pred, wdec, mu_logvar, z_pred, dec_in, fault, stream);
}

template <int ROWS>
static int launch_bottleneck_bwd(int batch, const void* d_dec_in, const float* wdec, const float* mu_logvar, const float* eps, float kld_grad_scale,
                                 const float* wfc, float* d_z_pred, float* d_mu_logvar, void* d_act, int* fault, cudaStream_t stream) {
    CVAE_OPT_IN_SMEM(bottleneck_bwd_kernel<ROWS>, bn_bwd_smem<ROWS>());
    cvae::launch(bottleneck_bwd_kernel<ROWS>, ((batch + ROWS - 1) / ROWS) * kDdSplit, ROWS * 16, bn_bwd_smem<ROWS>(), stream, batch,
                 (const __nv_bfloat16*)d_dec_in, wdec, mu_logvar, eps, kld_grad_scale, wfc, d_z_pred, d_mu_logvar, (__nv_bfloat16*)d_act, fault);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

// The data-gradient chain of the same three layers in one launch: d_dec_in bf16 [B][4096] -> d_z_pred fp32 [B][33] (optional),
// d_mu_logvar fp32 [B][64] (with the KL gradient kld_grad_scale * (mu | (exp(logvar) - 1) / 2) folded in), d_act bf16 [B][4096].
extern "C" int cvae_bottleneck_bwd(int batch, const void* d_dec_in, const float* wdec, const float* mu_logvar, const float* eps,
                                   float kld_grad_scale, const float* wfc, float* d_z_pred, float* d_mu_logvar, void* d_act,
                                   void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(batch > 0 && d_dec_in && wdec && mu_logvar && eps && wfc && d_mu_logvar && d_act, CVAE_EINVAL, "bottleneck_bwd: bad argument");
    int* fault = fault_flag();
    CVAE_REQUIRE(fault != nullptr, CVAE_ECUDA, "bottleneck_bwd: fault flag unavailable");
    static int cap_of_device[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (cap_of_device[dev] == 0) {
        CVAE_OPT_IN_SMEM(bottleneck_bwd_kernel<16>, bn_bwd_smem<16>());
        cap_of_device[dev] = max_clusters_of(bottleneck_bwd_kernel<16>, 256, bn_bwd_smem<16>());
    }
    const int max_clusters = cap_of_device[dev];
    if (bottleneck_rows(batch, max_clusters) == 20)
        return launch_bottleneck_bwd<20>(batch, d_dec_in, wdec, mu_logvar, eps, kld_grad_scale, wfc, d_z_pred, d_mu_logvar, d_act, fault, stream);
    return launch_bottleneck_bwd<16>(batch, d_dec_in, wdec, mu_logvar, eps, kld_grad_scale, wfc, d_z_pred, d_mu_logvar, d_act, fault, stream);
}

extern "C" int cvae_fc_fwd(int batch, const void* act, const float* wfc, const float* bias_mu,
                           const float* bias_var, float* mu_logvar, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(batch > 0 && act && wfc && bias_mu && bias_var && mu_logvar, CVAE_EINVAL, "fc_fwd: bad argument");
    CVAE_OPT_IN_SMEM(fc_fwd_kernel, kFcSmem);
    int* fault = fault_flag();
    CVAE_REQUIRE(fault != nullptr, CVAE_ECUDA, "fc_fwd: fault flag unavailable");
    cvae::launch(fc_fwd_kernel, ((batch + kFcRows - 1) / kFcRows) * kFcSplit, 256, kFcSmem, stream, batch, (const __nv_bfloat16*)act, wfc, bias_mu,
                                                                                          bias_var, mu_logvar, fault);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_fc_bwd(int batch, const float* d_mu_logvar, const void* act, const float* wfc,
                           void* d_act, float* dw_mu, float* dw_var, float* db_mu, float* db_var, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool want_w = dw_mu || dw_var || db_mu || db_var;
    CVAE_REQUIRE(batch > 0 && d_mu_logvar && (d_act || want_w), CVAE_EINVAL, "fc_bwd: bad argument");
    if (d_act) {
        CVAE_REQUIRE(wfc != nullptr, CVAE_EINVAL, "fc_bwd: data gradient needs the packed weights");
        cvae::launch(fc_bwd_data_kernel, dim3((batch + 7) / 8, 4), 256, 0, stream, batch, d_mu_logvar, wfc, (__nv_bfloat16*)d_act);
        CVAE_LAUNCH_CHECK();
    }
    if (want_w) {
        CVAE_REQUIRE(act && dw_mu && dw_var && db_mu && db_var, CVAE_EINVAL, "fc_bwd: weight gradient outputs go together");
        cvae::launch(fc_bwd_weight_kernel, 4096 / 32, 256, 0, stream, batch, d_mu_logvar, (const __nv_bfloat16*)act, dw_mu, dw_var, db_mu, db_var);
        CVAE_LAUNCH_CHECK();
    }
    return CVAE_OK;
}

extern "C" int cvae_decin_fwd(int batch, const float* z_pred, const float* wdec, void* out, void* stream) {
    CVAE_REQUIRE(batch > 0 && z_pred && wdec && out, CVAE_EINVAL, "decin_fwd: bad argument");
    cvae::launch(decin_fwd_kernel, dim3(4096 / 128, (batch + 7) / 8), 128, 0, (cudaStream_t)stream, batch, z_pred, wdec, (__nv_bfloat16*)out);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_decin_bwd(int batch, const void* d_out, const float* z_pred, const float* wdec,
                              float* d_z_pred, float* dw, float* db, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(batch > 0 && d_out && (d_z_pred || dw || db), CVAE_EINVAL, "decin_bwd: bad argument");
    if (d_z_pred) {
        CVAE_REQUIRE(wdec != nullptr, CVAE_EINVAL, "decin_bwd: data gradient needs the packed weights");
        CVAE_OPT_IN_SMEM(decin_bwd_data_kernel, kDdSmem);
        int* fault = fault_flag();
        CVAE_REQUIRE(fault != nullptr, CVAE_ECUDA, "decin_bwd: fault flag unavailable");
        cvae::launch(decin_bwd_data_kernel, ((batch + kDdRows - 1) / kDdRows) * kDdSplit, 256, kDdSmem, stream, 
            batch, (const __nv_bfloat16*)d_out, wdec, d_z_pred, fault);
        CVAE_LAUNCH_CHECK();
    }
    if (dw || db) {
        CVAE_REQUIRE(z_pred && dw && db, CVAE_EINVAL, "decin_bwd: weight gradient outputs go together");
        cvae::launch(decin_bwd_weight_kernel, 4096 / 32, 256, 0, stream, batch, (const __nv_bfloat16*)d_out, z_pred, dw, db);
        CVAE_LAUNCH_CHECK();
    }
    return CVAE_OK;
}
