// HBM-bound pieces around the conv GEMMs: BatchNorm finalize / apply+maxpool+activation forward and
// backward, the fused latent kernel (reparameterise + critic concat, and its backward), Adam.
// All are vectorised (16-byte) grid-stride kernels; reductions go warp-shuffle -> shared -> one
// double atomic per channel per CTA.
#include "common.cuh"

namespace cvae {

struct F8 { float v[8]; };
__device__ __forceinline__ F8 unpack8(uint4 u) {
    F8 f;
    f.v[0] = bf16_lo(u.x); f.v[1] = bf16_hi(u.x); f.v[2] = bf16_lo(u.y); f.v[3] = bf16_hi(u.y);
    f.v[4] = bf16_lo(u.z); f.v[5] = bf16_hi(u.z); f.v[6] = bf16_lo(u.w); f.v[7] = bf16_hi(u.w);
    return f;
}
__device__ __forceinline__ uint4 pack8(const F8& f) {
    return make_uint4(pack_bf16x2(f.v[0], f.v[1]), pack_bf16x2(f.v[2], f.v[3]),
                      pack_bf16x2(f.v[4], f.v[5]), pack_bf16x2(f.v[6], f.v[7]));
}

// ---------------------------------------------------------------------------------------------
// BatchNorm finalize: nn.BatchNorm2d defaults (vae_nets.py:70,75,80,85).  ss = [scale|shift|mean|invstd]
// ---------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(int C, double count, int training, const double* stats,
                                   const float* gamma, const float* beta, const float* conv_bias,
                                   float* rmean, float* rvar, long long* nbt, float momentum, float eps,
                                   float* ss) {
    grid_dependency_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mean, invstd;
    if (training) {
        // the conv ran WITHOUT its bias: a per-channel constant cancels in (x - mean); it only
        // shows up in the running mean.
        const double m = stats[c] / count;
        double var = stats[C + c] / count - m * m;
        if (var < 0) var = 0;
        mean = (float)m;
        invstd = (float)(1.0 / sqrt(var + (double)eps));
        const double unbiased = count > 1 ? var * count / (count - 1) : var;
        rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)(m + (double)conv_bias[c]);
        rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
        if (c == 0 && nbt) nbt[0] += 1;
    } else {
        mean = rmean[c] - conv_bias[c];
        invstd = 1.f / sqrtf(rvar[c] + eps);
    }
    const float sc = gamma[c] * invstd;
    ss[c] = sc;
    ss[C + c] = beta[c] - mean * sc;
    ss[2 * C + c] = mean;
    ss[3 * C + c] = invstd;
}

__device__ __forceinline__ float act_fn(int act, float x) { return act == 0 ? fmaxf(x, 0.f) : tanhf(x); }

// bn_finalize + bn_pool_act_fwd in ONE launch (the finalize kernel is a single small block: its 3 us plus a launch
// boundary, four times per forward pass, sat on the main chain).  Every CTA recomputes scale / shift / mean / invstd of
// all C <= 256 channels from the conv epilogue's statistics into shared memory (identical arithmetic in every CTA);
// block 0 also publishes them for the backward pass and updates the running buffers.
struct BnFusedArgs {
    double count;
    int training;
    const double* stats;
    const float *gamma, *beta, *conv_bias;
    float *rmean, *rvar;
    long long* nbt;
    float momentum, eps;
    float* ss_out;
};
__global__ void __launch_bounds__(256) bn_fused_fwd_kernel(int B, int H, int W, int C, int act, const uint4* __restrict__ x, const BnFusedArgs f,
                                                           uint4* __restrict__ y, uint4* __restrict__ xhat_max, uint16_t* __restrict__ argmax) {
    grid_dependency_sync();
    __shared__ __align__(16) float ss[4 * 256];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mean, invstd;
        if (f.training) {
            const double m = f.stats[c] / f.count;
            double var = f.stats[C + c] / f.count - m * m;
            if (var < 0) var = 0;
            mean = (float)m;
            invstd = (float)(1.0 / sqrt(var + (double)f.eps));
            if (blockIdx.x == 0) {
                const double unbiased = f.count > 1 ? var * f.count / (f.count - 1) : var;
                f.rmean[c] = (1.f - f.momentum) * f.rmean[c] + f.momentum * (float)(m + (double)f.conv_bias[c]);
                f.rvar[c] = (1.f - f.momentum) * f.rvar[c] + f.momentum * (float)unbiased;
                if (c == 0 && f.nbt) f.nbt[0] += 1;
            }
        } else {
            mean = f.rmean[c] - f.conv_bias[c];
            invstd = 1.f / sqrtf(f.rvar[c] + f.eps);
        }
        const float sc = f.gamma[c] * invstd;
        ss[c] = sc;
        ss[C + c] = f.beta[c] - mean * sc;
        ss[2 * C + c] = mean;
        ss[3 * C + c] = invstd;
        if (blockIdx.x == 0) {
            f.ss_out[c] = sc; f.ss_out[C + c] = ss[C + c]; f.ss_out[2 * C + c] = mean; f.ss_out[3 * C + c] = invstd;
        }
    }
    __syncthreads();
    const int cg = C >> 3, Ho = H >> 1, Wo = W >> 1;
    const long long total = (long long)B * Ho * Wo * cg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cg);
        long long p = i / cg;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int n = (int)(p / Ho);
        const float* sc = ss + c8 * 8;
        const float* sh = ss + C + c8 * 8;
        const size_t base = ((size_t)(n * H + 2 * ho) * W + 2 * wo) * cg + c8;
        F8 m, xm;
        uint32_t am = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const F8 v = unpack8(__ldg(x + base + (size_t)(q >> 1) * W * cg + (size_t)(q & 1) * cg));
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float b = fmaf(v.v[e], sc[e], sh[e]);
                if (q == 0 || b > m.v[e]) {   // first maximum wins, like nn.MaxPool2d
                    m.v[e] = b;
                    xm.v[e] = v.v[e];
                    am = (am & ~(3u << (2 * e))) | ((uint32_t)q << (2 * e));
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) m.v[e] = act_fn(act, m.v[e]);
        y[i] = pack8(m);
        if (xhat_max != nullptr) {
            const float* mean = ss + 2 * C + c8 * 8;
            const float* inv = ss + 3 * C + c8 * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) xm.v[e] = (xm.v[e] - mean[e]) * inv[e];
            xhat_max[i] = pack8(xm);
            argmax[i] = (uint16_t)am;
        }
    }
}

// y = act(maxpool2x2(x*scale + shift)); x bf16 NHWC [B][H][W][C] -> y bf16 NHWC [B][H/2][W/2][C].
// Training also saves what the backward needs per pooled element, so it does not have to redo the
// normalisation of all four positions: the normalised value at the arg-max position (bf16) and the
// arg-max position itself (2 bits per channel, one uint16 per 8-channel group).
__global__ void bn_pool_act_fwd_kernel(int B, int H, int W, int C, int act, const uint4* __restrict__ x,
                                       const float* __restrict__ ss, uint4* __restrict__ y,
                                       uint4* __restrict__ xhat_max, uint16_t* __restrict__ argmax) {
    grid_dependency_sync();
    const int cg = C >> 3, Ho = H >> 1, Wo = W >> 1;
    const long long total = (long long)B * Ho * Wo * cg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cg);
        long long p = i / cg;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int n = (int)(p / Ho);
        const float4 s0 = __ldg((const float4*)(ss + c8 * 8)), s1 = __ldg((const float4*)(ss + c8 * 8 + 4));
        const float4 h0 = __ldg((const float4*)(ss + C + c8 * 8)), h1 = __ldg((const float4*)(ss + C + c8 * 8 + 4));
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const size_t base = ((size_t)(n * H + 2 * ho) * W + 2 * wo) * cg + c8;
        F8 m, xm;
        uint32_t am = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const F8 v = unpack8(__ldg(x + base + (size_t)(q >> 1) * W * cg + (size_t)(q & 1) * cg));
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float b = fmaf(v.v[e], sc[e], sh[e]);
                if (q == 0 || b > m.v[e]) {   // first maximum wins, like nn.MaxPool2d
                    m.v[e] = b;
                    xm.v[e] = v.v[e];
                    am = (am & ~(3u << (2 * e))) | ((uint32_t)q << (2 * e));
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) m.v[e] = act_fn(act, m.v[e]);
        y[i] = pack8(m);
        if (xhat_max != nullptr) {
            const float4 m0 = __ldg((const float4*)(ss + 2 * C + c8 * 8)), m1 = __ldg((const float4*)(ss + 2 * C + c8 * 8 + 4));
            const float4 i0 = __ldg((const float4*)(ss + 3 * C + c8 * 8)), i1 = __ldg((const float4*)(ss + 3 * C + c8 * 8 + 4));
            const float mean[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
            const float inv[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) xm.v[e] = (xm.v[e] - mean[e]) * inv[e];
            xhat_max[i] = pack8(xm);
            argmax[i] = (uint16_t)am;
        }
    }
}

// Backward of act(maxpool(bn(x))).
// PASS 0 (pooled tensors only): per-channel sum(g), sum(g*xhat) with g = dy * act'(y) at the arg-max position.
// PASS 1: dx = gamma*invstd*(g_full - sum(g)/n - xhat*sum(g*xhat)/n) at all 4 positions
//            = [q == argmax] * (gi*g) + (c0 + c1*x_q)   with per-channel constants, plus dgamma/dbeta.
template <int PASS>
__global__ void bn_pool_act_bwd_kernel(int B, int H, int W, int C, int act, const uint4* __restrict__ x,
                                       const uint4* __restrict__ yact, const uint4* __restrict__ dy,
                                       const uint4* __restrict__ xhat_max, const uint16_t* __restrict__ argmax,
                                       const float* __restrict__ ss, const float* __restrict__ gamma,
                                       double* sums, uint4* __restrict__ dx, float* dgamma, float* dbeta) {
    grid_dependency_sync();
    extern __shared__ float red[];  // PASS 0: [blockDim][16]
    const int cg = C >> 3, Ho = H >> 1, Wo = W >> 1;
    const int c8 = threadIdx.x % cg;
    const int lanes = blockDim.x / cg;  // pixel lanes per block
    const long long npix = (long long)B * Ho * Wo;
    const double count = (double)B * H * W;
    float gi[8], k0[8], k1[8];
    if (PASS == 1) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = c8 * 8 + e;
            const float mean = ss[2 * C + c], inv = ss[3 * C + c];
            const float s1 = (float)(sums[c] / count), s2 = (float)(sums[C + c] / count);
            gi[e] = gamma[c] * inv;
            k1[e] = -gi[e] * s2 * inv;                 // coefficient of x_q
            k0[e] = -gi[e] * s1 - k1[e] * mean;        // constant term
        }
    }
    float a1[8], a2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a1[e] = a2[e] = 0.f;

    for (long long p = (long long)blockIdx.x * lanes + threadIdx.x / cg; p < npix; p += (long long)gridDim.x * lanes) {
        const F8 yv = unpack8(__ldg(yact + p * cg + c8));
        const F8 dv = unpack8(__ldg(dy + p * cg + c8));
        float g[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
            g[e] = dv.v[e] * (act == 0 ? (yv.v[e] > 0.f ? 1.f : 0.f) : (1.f - yv.v[e] * yv.v[e]));
        if (PASS == 0) {
            const F8 xh = unpack8(__ldg(xhat_max + p * cg + c8));
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                a1[e] += g[e];
                a2[e] = fmaf(g[e], xh.v[e], a2[e]);
            }
        } else {
            const int wo = (int)(p % Wo);
            const int ho = (int)((p / Wo) % Ho);
            const int n = (int)(p / ((long long)Wo * Ho));
            const size_t base = ((size_t)(n * H + 2 * ho) * W + 2 * wo) * cg + c8;
            const uint32_t am = argmax[p * cg + c8];
            F8 xv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) xv[q] = unpack8(__ldg(x + base + (size_t)(q >> 1) * W * cg + (size_t)(q & 1) * cg));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                F8 o;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float r = fmaf(k1[e], xv[q].v[e], k0[e]);
                    o.v[e] = (((am >> (2 * e)) & 3u) == (uint32_t)q) ? fmaf(gi[e], g[e], r) : r;
                }
                dx[base + (size_t)(q >> 1) * W * cg + (size_t)(q & 1) * cg] = pack8(o);
            }
        }
    }
    if (PASS == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            red[threadIdx.x * 16 + e] = a1[e];
            red[threadIdx.x * 16 + 8 + e] = a2[e];
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 2 * C; t += blockDim.x) {
            const int which = t / C, c = t % C, g8 = c >> 3, e = c & 7;
            float s = 0.f;
            for (int l = 0; l < lanes; ++l) s += red[(l * cg + g8) * 16 + which * 8 + e];
            atomicAdd(sums + which * C + c, (double)s);
        }
    } else if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            dbeta[c] = (float)sums[c];
            dgamma[c] = (float)sums[C + c];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Latent kernel (vae_nets.py:48-51 reparametrize, :143 critic concat) and its backward
// ---------------------------------------------------------------------------------------------
// One block = 64 batch rows, 256 threads.  Loads are float4 (a warp covers four rows: 4 x 128 contiguous bytes of mu,
// of logvar, and 512 contiguous bytes of eps), the 33-float output rows (z | pred, 132 bytes: only 4-byte aligned on
// their own) are staged in shared memory and leave as one contiguous, 16-byte aligned run of float4 stores.  The KL
// term of the same mu / logvar (vae_nets.py:57-58: sum(1 + lv - mu^2 - exp(lv))) is reduced with warp shuffles in the
// same pass; one double partial per block, summed in block order by loss_finalize_kernel (deterministic).
static constexpr int kLatRows = 64;
__global__ void __launch_bounds__(256) latent_fwd_kernel(int B, int sample, const float* __restrict__ ml, const float* __restrict__ eps,
                                                         const float* __restrict__ pred, float* __restrict__ zc,
                                                         double* __restrict__ kld_partial) {
    grid_dependency_sync();
    __shared__ __align__(16) float stage[kLatRows * 33];
    __shared__ double red[8];
    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kLatRows;
    const int rows = (int)min((long long)kLatRows, (long long)B - row0);
    double kl = 0.0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i = tid + 256 * k, r = i >> 3, d4 = i & 7;
        if (r < rows) {
            const float4* mrow = reinterpret_cast<const float4*>(ml + (row0 + r) * 64);
            const float4 mu = __ldg(mrow + d4), lv = __ldg(mrow + 8 + d4);
            float4 z = mu;
            if (sample) {
                const float4 e = __ldg(reinterpret_cast<const float4*>(eps + (row0 + r) * 32) + d4);
                z.x = fmaf(e.x, expf(0.5f * lv.x), mu.x);
                z.y = fmaf(e.y, expf(0.5f * lv.y), mu.y);
                z.z = fmaf(e.z, expf(0.5f * lv.z), mu.z);
                z.w = fmaf(e.w, expf(0.5f * lv.w), mu.w);
            }
            float* s = stage + r * 33 + d4 * 4;     // bank (r + 4 d4 + j) % 32: the 32 lanes of a warp hit 32 banks
            s[0] = z.x; s[1] = z.y; s[2] = z.z; s[3] = z.w;
            if (kld_partial) {
                kl += (double)(1.f + lv.x - mu.x * mu.x - expf(lv.x));
                kl += (double)(1.f + lv.y - mu.y * mu.y - expf(lv.y));
                kl += (double)(1.f + lv.z - mu.z * mu.z - expf(lv.z));
                kl += (double)(1.f + lv.w - mu.w * mu.w - expf(lv.w));
            }
        }
    }
    if (tid < rows) stage[tid * 33 + 32] = __ldg(pred + row0 + tid);
    if (kld_partial) {
        kl = warp_sum(kl);
        if ((tid & 31) == 0) red[tid >> 5] = kl;
    }
    __syncthreads();
    float* dst = zc + row0 * 33;                    // 64 * 33 * 4 bytes per block: every block starts 16-byte aligned
    const int n = rows * 33, n4 = n >> 2;
    for (int i = tid; i < n4; i += 256) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(stage)[i];
    for (int i = (n4 << 2) + tid; i < n; i += 256) dst[i] = stage[i];
    if (kld_partial && tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        kld_partial[blockIdx.x] = t;
    }
}

// d_ml[b][0:32] = dz + dmu_ext + k mu ; d_ml[b][32:64] = dz * eps * 0.5 * exp(0.5 logvar) + dlogvar_ext + k 0.5 (exp(logvar) - 1)
// with k = kld_grad_scale = kld_weight / B (times the upstream gradient): the KL term's backward (vae_nets.py:57-58) folded in.
// Thread = one float4 of one row: all loads and stores are 16 bytes except the 33-stride d_z row (scalar, L1-resident).
__global__ void latent_bwd_kernel(int B, const float* __restrict__ ml, const float* __restrict__ eps,
                                  const float* __restrict__ dzc, const float* __restrict__ dmu_ext,
                                  const float* __restrict__ dlv_ext, float kld_grad_scale, float* __restrict__ dml) {
    grid_dependency_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * 8) return;
    const long long b = i >> 3;
    const int d4 = (int)(i & 7);
    const float4* mrow = reinterpret_cast<const float4*>(ml + b * 64);
    const float4 mu = __ldg(mrow + d4), lv = __ldg(mrow + 8 + d4);
    const float4 e = __ldg(reinterpret_cast<const float4*>(eps + b * 32) + d4);
    const float* dzr = dzc + b * 33 + d4 * 4;
    const float dz[4] = {__ldg(dzr), __ldg(dzr + 1), __ldg(dzr + 2), __ldg(dzr + 3)};
    float4 gm = make_float4(0.f, 0.f, 0.f, 0.f), gl = gm;
    if (dmu_ext) gm = __ldg(reinterpret_cast<const float4*>(dmu_ext + b * 32) + d4);
    if (dlv_ext) gl = __ldg(reinterpret_cast<const float4*>(dlv_ext + b * 32) + d4);
    const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, lvv[4] = {lv.x, lv.y, lv.z, lv.w}, ev[4] = {e.x, e.y, e.z, e.w};
    const float gmv[4] = {gm.x, gm.y, gm.z, gm.w}, glv[4] = {gl.x, gl.y, gl.z, gl.w};
    float om[4], ol[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float std_ = expf(0.5f * lvv[j]);
        om[j] = dz[j] + gmv[j];
        ol[j] = dz[j] * ev[j] * 0.5f * std_ + glv[j];
        if (kld_grad_scale != 0.f) {
            om[j] += kld_grad_scale * muv[j];
            ol[j] += kld_grad_scale * 0.5f * (expf(lvv[j]) - 1.f);
        }
    }
    float4* orow = reinterpret_cast<float4*>(dml + b * 64);
    orow[d4] = make_float4(om[0], om[1], om[2], om[3]);
    orow[8 + d4] = make_float4(ol[0], ol[1], ol[2], ol[3]);
}

// ---------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults, vae.py:36): flat fp32 buffers, step counter on the device so the
// launch is CUDA-graph replayable.  grad_scale folds the 1/world_size of the data-parallel mean.
// ---------------------------------------------------------------------------------------------
__global__ void adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, const long long* __restrict__ step, float lr, float b1, float b2,
                            float eps, float grad_scale) {
    grid_dependency_sync();
    // the two double-precision pow() of the bias corrections once per block, not once per thread (they made the kernel
    // compute-bound: 20 us at 25 % of the DRAM rate, profiles/r02_full_final.md)
    __shared__ float bias_corr[2];
    if (threadIdx.x == 0) {
        const double t = (double)(step[0] + 1);
        bias_corr[0] = (float)(1.0 - pow((double)b1, t));
        bias_corr[1] = (float)sqrt(1.0 - pow((double)b2, t));
    }
    __syncthreads();
    const float bc1 = bias_corr[0], sq_bc2 = bias_corr[1];
    const float step_size = lr / bc1;
    auto update = [&](float& pi, float gi, float& mi, float& vi) {
        const float gr = gi * grad_scale;
        mi = mi + (1.f - b1) * (gr - mi);                          // lerp form used by torch
        vi = vi * b2 + (1.f - b2) * gr * gr;
        const float denom = sqrtf(vi) / sq_bc2 + eps;
        pi = pi - step_size * (mi / denom);
    };
    // float4 body (the flat buffers are 16-byte aligned), scalar tail of n % 4 elements in the last thread's wake
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        update(pv.x, gv.x, mv.x, vv.x);
        update(pv.y, gv.y, mv.y, vv.y);
        update(pv.z, gv.z, mv.z, vv.z);
        update(pv.w, gv.w, mv.w, vv.w);
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        update(p[i], g[i], m[i], v[i]);
    }
}
__global__ void adam_tick_kernel(long long* step) {
    grid_dependency_sync(); step[0] += 1; }

// uint8 HWC frames -> fp32 NCHW in [0,1]: astype(float32) / 255 then HWC->CHW, exactly
// vae_utility.py:324-328,337-341 (adjust_values + transpose), so the bits match the reference's input.
__global__ void frames_u8_kernel(long long n_pix, const uint8_t* __restrict__ src, float* __restrict__ dst) {
    grid_dependency_sync();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += (long long)gridDim.x * blockDim.x) {
        const long long f = i >> 12;
        const int p = (int)(i & 4095);
        const uint8_t* s = src + i * 3;
        float* d = dst + f * 3 * 4096 + p;
        d[0] = __fdiv_rn((float)s[0], 255.f);
        d[4096] = __fdiv_rn((float)s[1], 255.f);
        d[8192] = __fdiv_rn((float)s[2], 255.f);
    }
}

}  // namespace cvae

using namespace cvae;

static int grid_for(long long items, int threads, int per_sm = 8) {
    long long b = (items + threads - 1) / threads;
    const long long cap = (long long)sm_count() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

extern "C" int cvae_bn_finalize(int channels, int64_t count, int training, const double* stats,
                                const float* gamma, const float* beta, const float* conv_bias,
                                float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                float momentum, float eps, float* scale_shift, void* stream) {
    CVAE_REQUIRE(channels > 0 && count > 0, CVAE_EINVAL, "bn_finalize: empty");
    CVAE_REQUIRE(gamma && beta && conv_bias && running_mean && running_var && scale_shift, CVAE_EINVAL, "bn_finalize: null tensor");
    CVAE_REQUIRE(!training || stats, CVAE_EINVAL, "bn_finalize: training needs stats");
    cvae::launch(bn_finalize_kernel, (channels + 127) / 128, 128, 0, (cudaStream_t)stream, 
        channels, (double)count, training, stats, gamma, beta, conv_bias, running_mean, running_var,
        (long long*)num_batches_tracked, momentum, eps, scale_shift);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_bn_pool_act_fwd(int batch, int height, int width, int channels, int act,
                                    const void* conv_out, const float* scale_shift, void* out, void* xhat_max,
                                    void* argmax, void* stream) {
    CVAE_REQUIRE(batch > 0 && height % 2 == 0 && width % 2 == 0 && channels % 8 == 0, CVAE_EINVAL, "bn_pool_act_fwd: shape");
    CVAE_REQUIRE(conv_out && scale_shift && out, CVAE_EINVAL, "bn_pool_act_fwd: null tensor");
    CVAE_REQUIRE((xhat_max == nullptr) == (argmax == nullptr), CVAE_EINVAL, "bn_pool_act_fwd: xhat_max and argmax go together");
    const long long items = (long long)batch * (height / 2) * (width / 2) * (channels / 8);
    cvae::launch(bn_pool_act_fwd_kernel, grid_for(items, 256), 256, 0, (cudaStream_t)stream, 
        batch, height, width, channels, act, (const uint4*)conv_out, scale_shift, (uint4*)out, (uint4*)xhat_max,
        (uint16_t*)argmax);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_bn_fwd(int batch, int height, int width, int channels, int act, int training, const void* conv_out,
                           const double* stats, const float* gamma, const float* beta, const float* conv_bias, float* running_mean,
                           float* running_var, int64_t* num_batches_tracked, float momentum, float eps, float* scale_shift, void* out,
                           void* xhat_max, void* argmax, void* stream) {
    CVAE_REQUIRE(batch > 0 && height % 2 == 0 && width % 2 == 0 && channels % 8 == 0 && channels <= 256, CVAE_EINVAL, "bn_fwd: shape");
    CVAE_REQUIRE(conv_out && gamma && beta && conv_bias && running_mean && running_var && scale_shift && out, CVAE_EINVAL, "bn_fwd: null tensor");
    CVAE_REQUIRE(!training || stats, CVAE_EINVAL, "bn_fwd: training needs stats");
    CVAE_REQUIRE((xhat_max == nullptr) == (argmax == nullptr), CVAE_EINVAL, "bn_fwd: xhat_max and argmax go together");
    const long long items = (long long)batch * (height / 2) * (width / 2) * (channels / 8);
    BnFusedArgs f{(double)batch * height * width, training, stats, gamma, beta, conv_bias, running_mean, running_var,
                  (long long*)num_batches_tracked, momentum, eps, scale_shift};
    static const int fwd_per_sm = getenv("CVAE_BN_FWD_BLOCKS_PER_SM") ? atoi(getenv("CVAE_BN_FWD_BLOCKS_PER_SM")) : 8;      // (experiments)
    cvae::launch(bn_fused_fwd_kernel, grid_for(items, 256, fwd_per_sm), 256, 0, (cudaStream_t)stream, batch, height, width, channels, act, (const uint4*)conv_out, f,
                                                                               (uint4*)out, (uint4*)xhat_max, (uint16_t*)argmax);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_bn_pool_act_bwd(int batch, int height, int width, int channels, int act,
                                    const void* conv_out, const void* act_out, const void* d_act,
                                    const void* xhat_max, const void* argmax,
                                    const float* scale_shift, const float* gamma, double* sums,
                                    void* d_conv, float* dgamma, float* dbeta, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(batch > 0 && height % 2 == 0 && width % 2 == 0 && channels % 8 == 0 && channels <= 256 &&
                     256 % (channels / 8) == 0, CVAE_EINVAL, "bn_pool_act_bwd: shape");
    CVAE_REQUIRE(conv_out && act_out && d_act && xhat_max && argmax && scale_shift && gamma && sums && d_conv && dgamma && dbeta,
                 CVAE_EINVAL, "bn_pool_act_bwd: null tensor");
    CVAE_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * channels, stream));
    const int threads = 256, lanes = threads / (channels / 8);
    const long long npix = (long long)batch * (height / 2) * (width / 2);
    long long blocks = (npix + lanes - 1) / lanes;
    const long long cap = (long long)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    // The reduction pass ends with 2 C double atomics per block on the same 2 C addresses, and same-address atomics
    // serialise in L2: with 4 blocks per SM the small layers spent more time there than loading (8x8x256: 29 -> 25 us for
    // the pair of passes at one block per SM).  One block per SM is also the best choice for the step as a whole on the
    // large layers (1.358 -> 1.347 ms, profiles/r02_bn_bwd_pass0_grid.log: the pass shares the machine with the weight
    // gradients of the side stream).  CVAE_BN_P0_BLOCKS_PER_SM overrides.
    static const int p0_per_sm = getenv("CVAE_BN_P0_BLOCKS_PER_SM") && atoi(getenv("CVAE_BN_P0_BLOCKS_PER_SM")) > 0
                                     ? atoi(getenv("CVAE_BN_P0_BLOCKS_PER_SM")) : 1;
    long long blocks0 = (npix + lanes - 1) / lanes;
    if (blocks0 > (long long)sm_count() * p0_per_sm) blocks0 = (long long)sm_count() * p0_per_sm;
    cvae::launch(bn_pool_act_bwd_kernel<0>, (int)blocks0, threads, threads * 16 * sizeof(float), stream, 
        batch, height, width, channels, act, (const uint4*)conv_out, (const uint4*)act_out, (const uint4*)d_act,
        (const uint4*)xhat_max, (const uint16_t*)argmax, scale_shift, gamma, sums, nullptr, nullptr, nullptr);
    CVAE_LAUNCH_CHECK();
    static const int p1_per_sm = getenv("CVAE_BN_P1_BLOCKS_PER_SM") ? atoi(getenv("CVAE_BN_P1_BLOCKS_PER_SM")) : 4;        // (experiments)
    if (blocks > (long long)sm_count() * p1_per_sm) blocks = (long long)sm_count() * p1_per_sm;
    cvae::launch(bn_pool_act_bwd_kernel<1>, (int)blocks, threads, 0, stream, 
        batch, height, width, channels, act, (const uint4*)conv_out, (const uint4*)act_out, (const uint4*)d_act,
        (const uint4*)xhat_max, (const uint16_t*)argmax, scale_shift, gamma, sums, (uint4*)d_conv, dgamma, dbeta);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_latent_kld_partials(int batch) { return batch > 0 ? (batch + kLatRows - 1) / kLatRows : 0; }

extern "C" int cvae_latent_fwd(int batch, int sample, const float* mu_logvar, const float* eps,
                               const float* pred, float* z_pred, double* kld_partial, void* stream) {
    CVAE_REQUIRE(batch > 0 && mu_logvar && pred && z_pred && (!sample || eps), CVAE_EINVAL, "latent_fwd: bad argument");
    CVAE_REQUIRE(((uintptr_t)mu_logvar | (uintptr_t)eps | (uintptr_t)z_pred) % 16 == 0, CVAE_EINVAL, "latent_fwd: tensors must be 16-byte aligned");
    cvae::launch(latent_fwd_kernel, cvae_latent_kld_partials(batch), 256, 0, (cudaStream_t)stream, batch, sample, mu_logvar, eps, pred, z_pred, kld_partial);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_latent_bwd(int batch, const float* mu_logvar, const float* eps, const float* d_z_pred,
                               const float* dmu_ext, const float* dlogvar_ext, float kld_grad_scale, float* d_mu_logvar, void* stream) {
    CVAE_REQUIRE(batch > 0 && mu_logvar && eps && d_z_pred && d_mu_logvar, CVAE_EINVAL, "latent_bwd: bad argument");
    CVAE_REQUIRE(((uintptr_t)mu_logvar | (uintptr_t)eps | (uintptr_t)dmu_ext | (uintptr_t)dlogvar_ext | (uintptr_t)d_mu_logvar) % 16 == 0, CVAE_EINVAL,
                 "latent_bwd: tensors must be 16-byte aligned");
    const long long items = (long long)batch * 8;
    cvae::launch(latent_bwd_kernel, (int)((items + 255) / 256), 256, 0, (cudaStream_t)stream, batch, mu_logvar, eps, d_z_pred, dmu_ext, dlogvar_ext,
                                                                                   kld_grad_scale, d_mu_logvar);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_frames_u8_to_f32(int frames, const uint8_t* hwc_u8, float* nchw_f32, void* stream) {
    CVAE_REQUIRE(frames >= 0 && (frames == 0 || (hwc_u8 && nchw_f32)), CVAE_EINVAL, "frames_u8_to_f32: bad argument");
    if (frames == 0) return CVAE_OK;
    const long long n = (long long)frames * 4096;
    cvae::launch(frames_u8_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, n, hwc_u8, nchw_f32);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_adam_update(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                int64_t* step, float lr, float beta1, float beta2, float eps, float grad_scale, int tick,
                                void* stream);

extern "C" int cvae_adam_step(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                              int64_t* step, float lr, float beta1, float beta2, float eps, float grad_scale,
                              void* stream) {
    return cvae_adam_update(n, params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, grad_scale, 1, stream);
}

// One Adam update of a RANGE of the flat buffers (all four pointers offset alike, 16-byte aligned) with the step count the
// device counter will have after this optimizer step; `tick` != 0 advances the counter afterwards.  A step may be applied
// in several ranges (the training step updates everything but the first conv block while that block's gradient is still
// being computed): every range but the last passes tick = 0.
extern "C" int cvae_adam_update(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                int64_t* step, float lr, float beta1, float beta2, float eps, float grad_scale, int tick,
                                void* stream) {
    CVAE_REQUIRE(n > 0 && params && grads && exp_avg && exp_avg_sq && step, CVAE_EINVAL, "adam_step: bad argument");
    CVAE_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, CVAE_EINVAL,
                 "adam_step: the flat buffers must be 16-byte aligned");
    cvae::launch(adam_kernel, grid_for((n + 3) / 4, 256, 16), 256, 0, (cudaStream_t)stream, n, params, grads, exp_avg, exp_avg_sq,
                                                                    (const long long*)step, lr, beta1, beta2, eps, grad_scale);
    CVAE_LAUNCH_CHECK();
    if (tick) {
        cvae::launch(adam_tick_kernel, 1, 1, 0, (cudaStream_t)stream, (long long*)step);
        CVAE_LAUNCH_CHECK();
    }
    return CVAE_OK;
}
