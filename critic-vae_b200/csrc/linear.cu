// The three Linear layers around the latent (0.17 % of the network FLOPs): fc_mu || fc_var
// (vae_nets.py:98-99,108-109) on the NHWC-flattened 4x4x256 bottleneck, and decoder_input
// (vae_nets.py:137,143-144).  CUDA-core fp32 kernels: K = 4096 / 33 and N = 64 / 4096 are a poor
// fit for 128-row UMMA tiles, and these are latency-, not throughput-critical.
//
// Weight layouts come from pack.cu: wfc fp32 [4096 k'][64] (k' = pixel*256 + channel),
// wdec fp32 [34][4096 k'] (rows 0..32 = W^T, row 33 = bias).
#include "common.cuh"

namespace cvae {

// ---- fc forward: ml[b][j] = sum_k a[b][k] * wfc[k][j] + bias[j];  a bf16 [B][4096] ---------------
// grid B/4; block 256 = 64 j x 4 k-lanes; the K loop is staged through shared memory in slices of
// 512 and reduced in a fixed order, so the result is bit-reproducible run to run (no atomics).
__global__ void fc_fwd_kernel(int B, const __nv_bfloat16* __restrict__ a, const float* __restrict__ wfc,
                              const float* __restrict__ bmu, const float* __restrict__ bvar, float* __restrict__ ml) {
    __shared__ float sa[512][4];        // [k][row] of the current slice
    __shared__ float red[4][4][64];
    const int b0 = blockIdx.x * 4;
    const int j = threadIdx.x & 63, kl = threadIdx.x >> 6;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < 4096; k0 += 512) {
        __syncthreads();
        for (int i = threadIdx.x; i < 4 * 512; i += 256) {
            const int r = i >> 9, k = i & 511;
            sa[k][r] = (b0 + r < B) ? __bfloat162float(a[(size_t)(b0 + r) * 4096 + k0 + k]) : 0.f;
        }
        __syncthreads();
        for (int k = kl; k < 512; k += 4) {
            const float w = __ldg(wfc + (size_t)(k0 + k) * 64 + j);
            const float4 x0 = *reinterpret_cast<const float4*>(&sa[k][0]);
            acc[0] = fmaf(w, x0.x, acc[0]); acc[1] = fmaf(w, x0.y, acc[1]);
            acc[2] = fmaf(w, x0.z, acc[2]); acc[3] = fmaf(w, x0.w, acc[3]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) red[kl][r][j] = acc[r];
    __syncthreads();
    {
        const int r = threadIdx.x >> 6, jj = threadIdx.x & 63;
        if (b0 + r < B) {
            const float s = ((red[0][r][jj] + red[1][r][jj]) + (red[2][r][jj] + red[3][r][jj])) +
                            (jj < 32 ? bmu[jj] : bvar[jj - 32]);
            ml[(size_t)(b0 + r) * 64 + jj] = s;
        }
    }
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
__global__ void fc_bwd_weight_kernel(int B, const float* __restrict__ dml, const __nv_bfloat16* __restrict__ a,
                                     float* __restrict__ dwmu, float* __restrict__ dwvar,
                                     float* __restrict__ dbmu, float* __restrict__ dbvar) {
    __shared__ float sd[32][64];
    const int kp = blockIdx.x * blockDim.x + threadIdx.x;  // k' in NHWC order
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;
    for (int b0 = 0; b0 < B; b0 += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) {
            const int r = i >> 6;
            sd[r][i & 63] = (b0 + r < B) ? dml[(size_t)(b0 + r) * 64 + (i & 63)] : 0.f;
        }
        __syncthreads();
        const int nb = min(32, B - b0);
        for (int r = 0; r < nb; ++r) {
            const float x = __bfloat162float(a[(size_t)(b0 + r) * 4096 + kp]);
#pragma unroll
            for (int j = 0; j < 64; ++j) acc[j] = fmaf(x, sd[r][j], acc[j]);
        }
    }
    const int k = (kp & 255) * 16 + (kp >> 8);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        dwmu[(size_t)j * 4096 + k] = acc[j];
        dwvar[(size_t)j * 4096 + k] = acc[32 + j];
    }
    if (blockIdx.x == 0 && threadIdx.x < 64) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += dml[(size_t)b * 64 + threadIdx.x];
        if (threadIdx.x < 32) dbmu[threadIdx.x] = s; else dbvar[threadIdx.x - 32] = s;
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
__global__ void decin_bwd_data_kernel(int B, const __nv_bfloat16* __restrict__ dh, const float* __restrict__ wdec,
                                      float* __restrict__ dzc) {
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = warp; i < 33; i += nw) {
        float s = 0.f;
        for (int k = lane; k < 4096; k += 32)
            s = fmaf(__bfloat162float(dh[(size_t)b * 4096 + k]), __ldg(wdec + (size_t)i * 4096 + k), s);
        s = warp_sum(s);
        if (lane == 0) dzc[(size_t)b * 33 + i] = s;
    }
}

// ---- decoder_input backward (weights + bias), reference layout dW [4096 k][33], db [4096 k] ---------
__global__ void decin_bwd_weight_kernel(int B, const __nv_bfloat16* __restrict__ dh, const float* __restrict__ zc,
                                        float* __restrict__ dw, float* __restrict__ db) {
    __shared__ float sz[32][33];
    const int kp = blockIdx.x * blockDim.x + threadIdx.x;
    float acc[34];
#pragma unroll
    for (int i = 0; i < 34; ++i) acc[i] = 0.f;
    for (int b0 = 0; b0 < B; b0 += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) {
            const int r = i / 33;
            sz[r][i % 33] = (b0 + r < B) ? zc[(size_t)(b0 + r) * 33 + i % 33] : 0.f;
        }
        __syncthreads();
        const int nb = min(32, B - b0);
        for (int r = 0; r < nb; ++r) {
            const float g = __bfloat162float(dh[(size_t)(b0 + r) * 4096 + kp]);
#pragma unroll
            for (int i = 0; i < 33; ++i) acc[i] = fmaf(g, sz[r][i], acc[i]);
            acc[33] += g;
        }
    }
    const int k = (kp & 255) * 16 + (kp >> 8);
#pragma unroll
    for (int i = 0; i < 33; ++i) dw[(size_t)k * 33 + i] = acc[i];
    db[k] = acc[33];
}

}  // namespace cvae

using namespace cvae;

extern "C" int cvae_fc_fwd(int batch, const void* act, const float* wfc, const float* bias_mu,
                           const float* bias_var, float* mu_logvar, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(batch > 0 && act && wfc && bias_mu && bias_var && mu_logvar, CVAE_EINVAL, "fc_fwd: bad argument");
    fc_fwd_kernel<<<(batch + 3) / 4, 256, 0, stream>>>(batch, (const __nv_bfloat16*)act, wfc, bias_mu, bias_var, mu_logvar);
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
    fc_bwd_weight_kernel<<<4096 / 64, 64, 0, stream>>>(batch, d_mu_logvar, (const __nv_bfloat16*)act, dw_mu, dw_var, db_mu, db_var);
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
    decin_bwd_data_kernel<<<batch, 256, 0, stream>>>(batch, (const __nv_bfloat16*)d_out, wdec, d_z_pred);
    CVAE_LAUNCH_CHECK();
    decin_bwd_weight_kernel<<<4096 / 64, 64, 0, stream>>>(batch, (const __nv_bfloat16*)d_out, z_pred, dw, db);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}
