// The three Linear layers around the latent (0.17 % of the network FLOPs): fc_mu || fc_var
// (vae_nets.py:98-99,108-109) on the NHWC-flattened 4x4x256 bottleneck, and decoder_input
// (vae_nets.py:137,143-144).  CUDA-core fp32 kernels: K = 4096 / 33 and N = 64 / 4096 are a poor
// fit for 128-row UMMA tiles, and these are latency-, not throughput-critical.
//
// Weight layouts come from pack.cu: wfc fp32 [4096 k'][64] (k' = pixel*256 + channel),
// wdec fp32 [34][4096 k'] (rows 0..32 = W^T, row 33 = bias).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace cvae {

// ---- fc forward: ml[b][j] = sum_k a[b][k] * wfc[k][j] + bias[j];  a bf16 [B][4096] ---------------
// A thread-block cluster of 8 CTAs shares one block of 16 rows: CTA r of the cluster owns the K slice
// [512 r, 512 r + 512) and leaves its [16][64] partial in its own shared memory; after cluster.sync()
// CTA r sums the eight partials of rows 2r, 2r+1 over distributed shared memory in rank order and adds
// the bias -- one launch, no atomics, bit-reproducible.
static constexpr int kFcRows = 16, kFcSplit = 8, kFcSlice = 4096 / kFcSplit;
__global__ void __cluster_dims__(kFcSplit, 1, 1) __launch_bounds__(256)
fc_fwd_kernel(int B, const __nv_bfloat16* __restrict__ a, const float* __restrict__ wfc,
              const float* __restrict__ bmu, const float* __restrict__ bvar, float* __restrict__ ml) {
    __shared__ __align__(16) float sa[kFcSlice][kFcRows];   // [k][row] of this CTA's K slice (32 KB)
    __shared__ float part[kFcRows][64];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b0 = (blockIdx.x / kFcSplit) * kFcRows, k0 = rank * kFcSlice;
    for (int t = threadIdx.x; t < kFcSlice * (kFcRows / 4); t += 256) {
        const int k = t % kFcSlice, rq = t / kFcSlice;
        float4 v;
        float* pv = &v.x;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = b0 + rq * 4 + i;
            pv[i] = r < B ? __bfloat162float(a[(size_t)r * 4096 + k0 + k]) : 0.f;
        }
        *reinterpret_cast<float4*>(&sa[k][rq * 4]) = v;
    }
    __syncthreads();
    const int j = threadIdx.x & 63, rq = threadIdx.x >> 6;   // a warp shares rq: the float4 read is a broadcast
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* w = wfc + (size_t)k0 * 64 + j;
#pragma unroll 8
    for (int k = 0; k < kFcSlice; ++k) {
        const float wv = __ldg(w + (size_t)k * 64);
        const float4 x = *reinterpret_cast<const float4*>(&sa[k][rq * 4]);
        acc[0] = fmaf(wv, x.x, acc[0]); acc[1] = fmaf(wv, x.y, acc[1]);
        acc[2] = fmaf(wv, x.z, acc[2]); acc[3] = fmaf(wv, x.w, acc[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) part[rq * 4 + i][j] = acc[i];
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

// ---- fc backward (data): da[b][k'] = sum_j dml[b][j] * wfc[k'][j]  (gradient w.r.t. the Tanh output) ----
__global__ void fc_bwd_data_kernel(int B, const float* __restrict__ dml, const float* __restrict__ wfc,
                                   __nv_bfloat16* __restrict__ da) {
    __shared__ float sd[8][64];
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
        const float4* wr = reinterpret_cast<const float4*>(wfc + (size_t)k * 64);
#pragma unroll 4
        for (int q = 0; q < 16; ++q) {
            const float4 w = __ldg(wr + q);
#pragma unroll
            for (int r = 0; r < 8; ++r)
                acc[r] += w.x * sd[r][4 * q] + w.y * sd[r][4 * q + 1] + w.z * sd[r][4 * q + 2] + w.w * sd[r][4 * q + 3];
        }
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
// block = 4 batch rows staged in shared memory as bf16; warp w owns outputs i = w, w+8, ...; every
// weight value loaded from L2 is used for the 4 rows.
static constexpr int kDdRows = 4;
__global__ void __launch_bounds__(256) decin_bwd_data_kernel(int B, const __nv_bfloat16* __restrict__ dh, const float* __restrict__ wdec,
                                                             float* __restrict__ dzc) {
    __shared__ __align__(16) uint32_t sdh[kDdRows][2048];   // bf16 pairs
    const int b0 = blockIdx.x * kDdRows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < kDdRows * 512; t += 256) {
        const int r = t >> 9, q = t & 511;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (b0 + r < B) v = __ldg(reinterpret_cast<const uint4*>(dh + (size_t)(b0 + r) * 4096) + q);
        *reinterpret_cast<uint4*>(&sdh[r][q * 4]) = v;
    }
    __syncthreads();
    for (int i = warp; i < 33; i += 8) {
        float acc[kDdRows];
#pragma unroll
        for (int r = 0; r < kDdRows; ++r) acc[r] = 0.f;
        const float2* w = reinterpret_cast<const float2*>(wdec + (size_t)i * 4096);
#pragma unroll 4
        for (int t = 0; t < 64; ++t) {
            const int kk = t * 32 + lane;   // pair index
            const float2 wv = __ldg(w + kk);
#pragma unroll
            for (int r = 0; r < kDdRows; ++r) {
                const uint32_t u = sdh[r][kk];
                acc[r] = fmaf(bf16_lo(u), wv.x, acc[r]);
                acc[r] = fmaf(bf16_hi(u), wv.y, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < kDdRows; ++r) {
            const float sacc = warp_sum(acc[r]);
            if (lane == 0 && b0 + r < B) dzc[(size_t)(b0 + r) * 33 + i] = sacc;
        }
    }
}

// ---- decoder_input backward (weights + bias), reference layout dW [4096 k][33], db [4096 k] ---------
// same decomposition as fc_bwd_weight_kernel: 32 k' x 8 batch lanes, fixed-order combine.
__global__ void __launch_bounds__(256) decin_bwd_weight_kernel(int B, const __nv_bfloat16* __restrict__ dh, const float* __restrict__ zc,
                                                               float* __restrict__ dw, float* __restrict__ db) {
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

}  // namespace cvae

using namespace cvae;

extern "C" int cvae_fc_fwd(int batch, const void* act, const float* wfc, const float* bias_mu,
                           const float* bias_var, float* mu_logvar, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(batch > 0 && act && wfc && bias_mu && bias_var && mu_logvar, CVAE_EINVAL, "fc_fwd: bad argument");
    fc_fwd_kernel<<<((batch + kFcRows - 1) / kFcRows) * kFcSplit, 256, 0, stream>>>(batch, (const __nv_bfloat16*)act, wfc, bias_mu, bias_var, mu_logvar);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_fc_bwd(int batch, const float* d_mu_logvar, const void* act, const float* wfc,
                           void* d_act, float* dw_mu, float* dw_var, float* db_mu, float* db_var, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(batch > 0 && d_mu_logvar && act && wfc && d_act && dw_mu && dw_var && db_mu && db_var, CVAE_EINVAL,
                 "fc_bwd: bad argument");
    fc_bwd_data_kernel<<<dim3((batch + 7) / 8, 4), 256, 0, stream>>>(batch, d_mu_logvar, wfc, (__nv_bfloat16*)d_act);
    CVAE_LAUNCH_CHECK();
    fc_bwd_weight_kernel<<<4096 / 32, 256, 0, stream>>>(batch, d_mu_logvar, (const __nv_bfloat16*)act, dw_mu, dw_var, db_mu, db_var);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_decin_fwd(int batch, const float* z_pred, const float* wdec, void* out, void* stream) {
    CVAE_REQUIRE(batch > 0 && z_pred && wdec && out, CVAE_EINVAL, "decin_fwd: bad argument");
    decin_fwd_kernel<<<dim3(4096 / 128, (batch + 7) / 8), 128, 0, (cudaStream_t)stream>>>(batch, z_pred, wdec, (__nv_bfloat16*)out);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_decin_bwd(int batch, const void* d_out, const float* z_pred, const float* wdec,
                              float* d_z_pred, float* dw, float* db, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(batch > 0 && d_out && z_pred && wdec && d_z_pred && dw && db, CVAE_EINVAL, "decin_bwd: bad argument");
    decin_bwd_data_kernel<<<(batch + kDdRows - 1) / kDdRows, 256, 0, stream>>>(batch, (const __nv_bfloat16*)d_out, wdec, d_z_pred);
    CVAE_LAUNCH_CHECK();
    decin_bwd_weight_kernel<<<4096 / 32, 256, 0, stream>>>(batch, (const __nv_bfloat16*)d_out, z_pred, dw, db);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}
